"""Reference-facing modules of the hot path: ``GradRegularizer`` and ``LabelSmoothCrossEntropyLoss``
(reference fullbatch/models/modules.py:86-101,136-348) backed by the B200 engine.

``GradRegularizer(model, optimizer, loss_fn, norm, block_strength, acc_strength, eps, implementation, mixed_precision)``
keeps the reference signature and call protocol: ``gradreg(grads, inputs, labels, pre_grads)`` mutates the tensors of
``grads`` in place, returns the list, leaves ``model.parameters()`` unchanged and reads
``optimizer.param_groups[0]["lr"]`` at call time (modules.py:214).  ``create_graph`` is False (modules.py:168-170).
The accelerated implementations are ``forward-differences`` and its README alias ``finite_diff`` (README.md:38-39 /
modules.py:139 default, which the reference's own dispatch rejects); the other reference implementations are not on the
B200 path and raise ``ValueError`` exactly like an unknown string does in the reference (modules.py:174-175).
"""
import torch

from .engine import FullBatchEngine

ACCELERATED = ("finite_diff", "forward-differences", "forward-differences-legacy", "central-differences")
REFERENCE_ONLY = ("autograd-pen", "autograd", "complex-step")


class LabelSmoothCrossEntropyLoss(torch.nn.Module):
    """modules.py:86-101.  Inside the engine the loss is evaluated by the fused head kernel (fb_head_fwd_bwd); this
    module carries ``smoothing`` to it and offers the same forward for evaluation code."""

    def __init__(self, smoothing=0.0, loss_modification=""):
        super().__init__()
        self.smoothing = smoothing

    def forward(self, input, target):
        log_prob = torch.nn.functional.log_softmax(input, dim=-1)
        weight = torch.full_like(input, self.smoothing / (input.shape[-1] - 1.0))
        weight.scatter_(-1, target.unsqueeze(-1), 1.0 - self.smoothing)
        return (-weight * log_prob).sum(dim=-1).mean()


class GradRegularizer:
    """Modify given iterable of gradients outside of autograd -- on the sm_100a kernels."""

    def __init__(self, model, optimizer, loss_fn, norm=2, block_strength=0.1, acc_strength=0.0, eps=1e-2,
                 implementation="finite_diff", mixed_precision=False, engine=None, microbatch=None, precision="split"):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.norm, self.block_strength, self.acc_strength, self.eps = norm, block_strength, acc_strength, eps
        self.mixed_precision = mixed_precision
        self.create_graph = False
        if self.block_strength == 0 and self.acc_strength == 0:
            self.forward = self._pass  # modules.py:151-153
        elif implementation in ACCELERATED:
            if norm != 2:
                raise ValueError("Only the 2-norm penalty is implemented by forward differences.")
            if mixed_precision:
                raise ValueError("mixed_precision (fp16 autocast) is not on the B200 path; use impl.precision instead.")
            self.forward = self._finite_differences
            self.implementation = implementation
        elif implementation in REFERENCE_ONLY:
            raise ValueError(f"Regularizer implementation {implementation} is not available on the B200 path "
                             f"(accelerated: {ACCELERATED}).")
        else:
            raise ValueError(f"Invalid spec. given for regularizer implementation: {implementation}")  # modules.py:175
        self._engine = engine
        self._microbatch = microbatch
        self._precision = precision

    @property
    def engine(self):
        if self._engine is None:
            if self._microbatch is None:
                raise RuntimeError("GradRegularizer needs an engine or the microbatch size to build one")
            smoothing = getattr(self.loss_fn, "smoothing", 0.0) or 0.0
            self._engine = FullBatchEngine(self.model, self._microbatch, precision=self._precision,
                                           label_smoothing=smoothing)
        return self._engine

    def _pass(self, grads, inputs, labels, pre_grads):
        return grads

    def _finite_differences(self, grads, inputs, labels, pre_grads):
        """modules.py:211-300 on the device: eps_n, theta' = theta +- eps_n*v with v = bs*g (+ acc*pre_grads), extra
        forward/backward pass(es) on the same microbatch, g += (lr/4) * vhp.  theta is never modified."""
        if self._microbatch is None:
            self._microbatch = inputs.shape[0]
        eng = self.engine
        if inputs.shape[0] != eng.mb:
            raise RuntimeError(f"GradRegularizer was built for microbatches of {eng.mb}, got {inputs.shape[0]}")
        lr = self.optimizer.param_groups[0]["lr"]  # modules.py:214, read at call time
        eng.load_grads(grads)
        acc = self.acc_strength if pre_grads is not None else 0.0  # modules.py:219-221
        if acc != 0:
            eng.load_pre(pre_grads)
        eng.regularize(inputs, labels, lr, self.block_strength, self.eps, self.implementation, acc)
        eng.store_grads(grads)
        return grads

    def __call__(self, grads, inputs, labels, pre_grads=None):
        return self.forward(grads, inputs, labels, pre_grads)
