"""ORACLE support (test infrastructure only) -- drives the UNMODIFIED reference implementation.

Works only where the reference tree is mounted (the build container: ``/root/reference``).  It cannot travel to the
GPU box; ``oracle/make_goldens.py`` uses it to generate the committed fixtures under ``tests/golden``.

Shims, none of which changes reference arithmetic:
  1. ``hydra`` / ``omegaconf`` are not installed -> stub modules before ``import fullbatch`` (fullbatch/utils.py:15-16),
     ``get_log`` replaced by a plain logger, ``cfg`` passed as an attribute dict.
  2. torch>=2 refuses in-place ``_foreach`` ops on leaf parameters outside ``no_grad`` (modules.py:226) -> wrap
     ``torch._foreach_{add_,sub_,div_,mul_}`` in ``no_grad`` (semantics preserving).
  3. ``setup`` dict and a ``TensorDataset`` loader built by hand (utils.py:79-80 refuses CPU; data_preparation.py:118
     would download CIFAR).
"""
import contextlib
import logging
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("FB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fullbatch"))


class AttrDict(dict):
    """Attribute dict with ``.items()`` / ``**`` support (training.py:71, optimizers.py:12)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    if "hydra" not in sys.modules:
        hydra = types.ModuleType("hydra")
        hydra.main = lambda *a, **k: (lambda f: f)
        hydra.core = types.ModuleType("hydra.core")
        hydra.core.hydra_config = types.ModuleType("hydra.core.hydra_config")
        hydra.core.hydra_config.HydraConfig = type("HydraConfig", (), {})
        hydra.utils = types.ModuleType("hydra.utils")
        hydra.utils.get_original_cwd = os.getcwd
        sys.modules.update({"hydra": hydra, "hydra.core": hydra.core, "hydra.utils": hydra.utils,
                            "hydra.core.hydra_config": hydra.core.hydra_config})
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        oc.OmegaConf = type("OmegaConf", (), {"to_yaml": staticmethod(lambda c: str(c))})
        oc.open_dict = lambda cfg: contextlib.nullcontext()
        oc.DictConfig = dict
        sys.modules["omegaconf"] = oc
    for name in ("_foreach_add_", "_foreach_sub_", "_foreach_div_", "_foreach_mul_"):
        fn = getattr(torch, name)
        if getattr(fn, "_fb_wrapped", False):
            continue

        def wrapped(*a, _fn=fn, **k):
            with torch.no_grad():
                return _fn(*a, **k)

        wrapped._fb_wrapped = True
        setattr(torch, name, wrapped)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fullbatch  # noqa: F401
    import fullbatch.training.training as T

    T.get_log = lambda cfg, name=None: logging.getLogger("fb_reference")
    _installed = True


def make_cfg(depth=18, batch_size=128, sub_batch=128, lr=0.8, block_strength=0.5, eps=1e-2, steps=1, warmup=0,
             implementation="forward-differences", grad_clip=0.25, label_smoothing=0.0, train_stochastic=False,
             size=50000, accumulation_dtype="float"):
    """Attribute-dict cfg with the keys listed in SURVEY.md Appendix B (hyp=gradreg defaults)."""
    return to_attr(dict(
        name="oracle", dryrun=False, seed=0, original_cwd=os.getcwd(),
        data=dict(batch_size=batch_size, channels=3, classes=10, pixels=32, size=size, name="CIFAR10"),
        model=dict(name=f"ResNet{depth}", depth=depth, width=64, stem="CIFAR", convolution="Standard", nonlin_fn="ReLU",
                   normalization="BatchNorm2d", downsample="C", initialization="skip-residual"),
        impl=dict(accumulation_dtype=accumulation_dtype, mixed_precision=False, non_blocking=True, validate_every_nth_step=100,
                  checkpoint=dict(name=None, save_every_nth_step=1000), setup=dict(dist=False, world_size=1),
                  dtype="float"),
        hyp=dict(steps=steps, sub_batch=sub_batch,
                 grad_reg=dict(norm=2, block_strength=block_strength, acc_strength=0.0, eps=eps,
                               implementation=implementation),
                 evaluate_ema=False, eval_ema_momentum=0.995,
                 optim=dict(name="Gradient Descent", lr=lr, momentum=0.9, weight_decay=5e-4, dampening=0.0,
                            nesterov=True, line_search="none"),
                 optim_modification=dict(name="none"), only_linear_layers_weight_decay=False,
                 scheduler="cosine-4000", warmup=warmup, batch_clip=None,
                 norm_bias=dict(strength=0.0, norm_type=1, bias=0), grad_clip=grad_clip, grad_clip_norm=2,
                 grad_noise=dict(additive=None, multiplicative=None), train_stochastic=train_stochastic,
                 train_switch_stochastic=None, train_semi_stochastic=False, stop_at_full_training_accuracy=0,
                 test_time_flips=False, label_smoothing=label_smoothing, loss_modification=None, shuffle=False),
        analysis=dict(type=None, check_every_nth_step=100, save_model_every_nth_step=None),
    ))


def make_loader(X, Y, batch_size):
    """data_preparation.py:56-72: SequentialSampler with a no-op set_epoch, drop_last=True."""
    ds = torch.utils.data.TensorDataset(X, Y)
    sampler = torch.utils.data.SequentialSampler(ds)
    sampler.set_epoch = lambda *a, **k: None
    return torch.utils.data.DataLoader(ds, batch_size=min(batch_size, len(ds)), sampler=sampler, drop_last=True,
                                       num_workers=0)


def construct_reference_model(cfg, seed=0, dtype=torch.float32):
    install_shims()
    from fullbatch.models import construct_model

    torch.manual_seed(seed)
    model = construct_model(cfg.model, cfg.data.channels, cfg.data.classes)
    return model.to(dtype)


def run_reference_train(model, X, Y, cfg, dtype=torch.float32, record=None):
    """Run ``fullbatch.training.train`` for cfg.hyp.steps steps on CPU.

    ``record`` (a list) receives, per GradRegularizer call, dict(raw=[...], reg=[...]) of cloned gradient lists.
    Returns (stats, accumulated gradient list); use cfg.hyp.grad_clip=None so the gradient is not clipped in place.
    """
    install_shims()
    import fullbatch.training.training as T
    from fullbatch.models import modules as M

    setup = dict(device=torch.device("cpu"), dtype=dtype, memory_format=torch.contiguous_format)
    loader = make_loader(X.to(dtype), Y, cfg.data.batch_size)
    valid = make_loader(X[: cfg.data.batch_size].to(dtype), Y[: cfg.data.batch_size], cfg.data.batch_size)

    orig_call = M.GradRegularizer.__call__

    def rec_call(self, grads, inputs, labels, pre_grads):
        raw = [g.detach().clone() for g in grads]
        out = orig_call(self, grads, inputs, labels, pre_grads)
        if record is not None:
            record.append(dict(raw=raw, reg=[g.detach().clone() for g in out]))
        return out

    M.GradRegularizer.__call__ = rec_call
    try:
        stats = T.train(model, loader, valid, setup, cfg)
    finally:
        M.GradRegularizer.__call__ = orig_call
    # with cfg.hyp.grad_clip=None, param.grad still holds the accumulated gradient (training.py:183; torch.optim.SGD
    # never writes to .grad)
    # (the stochastic branch zeroes the gradients after every optimizer step, training.py:280: None there)
    return stats, [p.grad.detach().clone() if p.grad is not None else None for p in model.parameters()]
