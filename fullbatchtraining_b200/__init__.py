"""fullbatchtraining_b200 -- B200 (sm_100a) implementation of the full-batch gradient-regularised step of
JonasGeiping/fullbatchtraining behind the reference's own Python surface (construct_model, GradRegularizer, train).

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + C ABI), ``lib`` (ctypes binding), ``ops`` (host-side
descriptors), ``engine`` (whole-microbatch executor), ``models`` / ``modules`` / ``training`` (the reference-facing
mirror).  There is no CPU fallback: importing is cheap and works anywhere, running needs the built library and a B200.
"""
from .models import construct_model, ResNet  # noqa: F401

__all__ = ["construct_model", "ResNet"]
