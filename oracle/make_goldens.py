"""ORACLE support (test infrastructure only): generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_goldens [case ...]

Each fixture holds fingerprints (per-tensor norms and sums + a strided sample of the flat vector, see
``fb_oracle.fingerprint``) of: the seed-0 initialisation, the accumulated regularised gradient after one full-batch
step of ``fullbatch.training.train`` (training.py:121-185), the raw and regularised gradients of the first two
microbatches (recorded around ``GradRegularizer.__call__``, modules.py:346-348), BN running statistics after the step,
and the scalar ``stats`` of ``_record_stats`` (training.py:85-119).  Inputs are ``fb_oracle.synthetic_cifar`` (seed 1234).
"""
import json
import os
import sys
import time

import numpy as np
import torch

from . import fb_oracle as O
from . import reference_harness as H

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (depth, microbatch, N, dtype)
CASES = {
    "r18_mb16_n32_f64": (18, 16, 32, "float64"),
    "r18_mb16_n32_f32": (18, 16, 32, "float32"),
    "r18_mb128_n256_f64": (18, 128, 256, "float64"),
    "r18_mb128_n256_f32": (18, 128, 256, "float32"),
    "r152_mb4_n8_f64": (152, 4, 8, "float64"),
    # remaining hyp.grad_reg variants (modules.py:243-300, training.py:128-142); 5th entry = overrides
    "r18_mb16_n48_f64_central": (18, 16, 48, "float64", dict(implementation="central-differences")),
    "r18_mb16_n48_f64_legacy": (18, 16, 48, "float64", dict(implementation="forward-differences-legacy")),
    "r18_mb16_n48_f64_acc": (18, 16, 48, "float64", dict(acc_strength=0.3)),
    "r18_mb16_n48_f64_central_acc": (18, 16, 48, "float64", dict(implementation="central-differences", acc_strength=0.2)),
    # BASELINE.json configs[0]: 2k images -> K = 15 microbatches of 128 (drop_last), fp64 truth and the fp32 reference
    "r18_mb128_n2000_f64": (18, 128, 2000, "float64"),
    "r18_mb128_n2000_f32": (18, 128, 2000, "float32"),
    # ResNet-152 at its real microbatch size (config 4, train.sh:10-12)
    "r152_mb32_n64_f64": (152, 32, 64, "float64"),
    # the stochastic sanity branch (training.py:241-286): two steps of 4 blocks with the regulariser per block
    "r18_sgd_mb16_n64_f64": (18, 16, 64, "float64", dict(train_stochastic=True, steps=2, lr=0.001)),
}
STRIDE = 4999
HYP = dict(lr=0.8, block_strength=0.5, eps=1e-2)


def pack(prefix, fp, out):
    for k, v in fp.items():
        out[f"{prefix}.{k}"] = np.asarray(v)


def make_case(name):
    depth, mb, n, dts = CASES[name][:4]
    extra = CASES[name][4] if len(CASES[name]) > 4 else {}
    dt = getattr(torch, dts)
    hyp = dict(HYP, lr=extra.get("lr", HYP["lr"]))
    cfg = H.make_cfg(depth=depth, batch_size=mb, sub_batch=mb, grad_clip=None, warmup=0,
                     accumulation_dtype="double" if dt == torch.float64 else "float",
                     implementation=extra.get("implementation", "forward-differences"),
                     train_stochastic=extra.get("train_stochastic", False), steps=extra.get("steps", 1), **hyp)
    cfg.hyp.grad_reg.acc_strength = extra.get("acc_strength", 0.0)
    model = H.construct_reference_model(cfg, seed=0, dtype=dt)
    X, Y = O.synthetic_cifar(n, dtype=dt)
    out = {}
    pack("init", O.fingerprint([p.detach() for p in model.parameters()], STRIDE), out)
    rec = []
    t0 = time.time()
    stats, avg = H.run_reference_train(model, X, Y, cfg, dtype=dt, record=rec)
    elapsed = time.time() - t0
    stochastic = extra.get("train_stochastic", False)
    if stochastic:  # the parameters after the optimizer steps are the result
        pack("theta", O.fingerprint([p.detach() for p in model.parameters()], STRIDE), out)
    else:
        pack("avg", O.fingerprint(avg, STRIDE), out)
    for i, r in enumerate([] if stochastic else rec[:2]):
        pack(f"mb{i}.raw", O.fingerprint(r["raw"], STRIDE), out)
        pack(f"mb{i}.reg", O.fingerprint(r["reg"], STRIDE), out)
    bufs = [b.detach() for k, b in model.named_buffers() if not k.endswith("num_batches_tracked")]
    pack("buffers", O.fingerprint(bufs, 97), out)
    scalars = {k: float(v[0]) for k, v in stats.items() if len(v) and k in
               ("train_loss", "train_acc", "param_norm", "grad_norm", "full_loss")}
    scalars["grad_norm_train"] = [float(stats[f"grad_norm_train_{i}"][0]) for i in range(n // mb)]
    if stochastic:
        scalars["train_loss_steps"] = [float(v) for v in stats["train_loss"]]
        scalars["train_acc_steps"] = [float(v) for v in stats["train_acc"]]
        scalars["grad_norm_steps"] = [float(v) for v in stats["grad_norm"]]
    meta = dict(case=name, depth=depth, mb=mb, n=n, dtype=dts, stride=STRIDE, hyp=hyp, extra=extra, scalars=scalars,
                torch=torch.__version__, threads=torch.get_num_threads(), reference_seconds=elapsed,
                num_params=int(sum(p.numel() for p in model.parameters())))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(f"[golden] {name}: {elapsed:.1f}s {scalars}", flush=True)


def make_scheduler_state():
    """State dict and learning-rate trace of the reference's scheduler stack (optimizers.py:81-91: CosineAnnealingLR(4000)
    behind GradualWarmupScheduler(total_epoch=3)) after 6 steps -> tests/golden/ref_scheduler_state.pt; the drop-in's
    schedulers.LinearWarmup must load it and continue with the same learning rates (training/utils.py:43-70)."""
    H.install_shims()
    from fullbatch.training.additional_optimizers.scheduler import GradualWarmupScheduler

    w = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.SGD([w], lr=0.8, momentum=0.9, nesterov=True, weight_decay=5e-4)
    after = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 4000, eta_min=0.0)
    sched = GradualWarmupScheduler(opt, multiplier=1.0, total_epoch=3, after_scheduler=after)
    lrs = []
    for _ in range(6):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    state = sched.state_dict()
    more = []
    for _ in range(4):
        more.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    torch.save(dict(state=state, lrs=lrs, lrs_after=more, base_lr=0.8, warmup=3, torch=torch.__version__),
               os.path.join(GOLDEN_DIR, "ref_scheduler_state.pt"))
    print("[golden] scheduler", lrs, more, sorted(state.keys()), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for case in (sys.argv[1:] or list(CASES)):
        if case == "scheduler":
            make_scheduler_state()
        else:
            make_case(case)
