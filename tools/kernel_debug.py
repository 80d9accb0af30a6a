"""Development aid: time single conv launches and print the in-kernel role wait counters (FB_KERNEL_DEBUG=1)."""
import ctypes
import os
import sys

os.environ["FB_KERNEL_DEBUG"] = "1"
sys.path.insert(0, ".")
import torch  # noqa: E402

from fullbatchtraining_b200 import lib, ops  # noqa: E402

DEV = "cuda"
NAMES = ["prod wait a_empty", "prod wait b_empty", "mma wait a_full", "mma wait b_full", "mma wait acc_empty", "mma total",
         "epi wait acc_full", "epi store", "epi total", "tiles", "prod total", "mma issue (incl. backpressure)"]


def counters(clear=True):
    buf = (ctypes.c_longlong * 32)()
    lib.check(lib.load().fb_debug_counters(buf, int(clear)), "fb_debug_counters")
    return list(buf)


def run(n, h, w, cin, cout, split, what):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, h, w, cin, device=DEV, generator=g)
    x_hi = x.to(torch.bfloat16)
    x_lo = (x - x_hi.float()).to(torch.bfloat16) if split else None
    bf = dict(device=DEV, dtype=torch.bfloat16)
    wf_hi, wd_hi = torch.randn(cout, 9 * cin, **bf), torch.randn(cin, 9 * cout, **bf)
    wf_lo = torch.randn(cout, 9 * cin, **bf) if split else None
    wd_lo = torch.randn(cin, 9 * cout, **bf) if split else None
    y = torch.empty(n, h, w, cout, device=DEV)
    dy = torch.randn(n, h, w, cout, **bf)
    dx = torch.zeros(n, h, w, cin, device=DEV)
    partial = torch.empty(ops.Conv2dPlan.partial_elems(n, h, w, cin, cout, 3, 1), device=DEV)
    plan = ops.Conv2dPlan(n, h, w, cin, cout, 3, 1, x_hi, x_lo, y, dy, dx, wf_hi, wf_lo, wd_hi, wd_lo, partial,
                          dx_accumulate=True, split=split)
    gw = torch.empty(cout, cin, 3, 3, device=DEV)
    fn = {"fwd": plan.forward, "dgrad": plan.dgrad, "wgrad": lambda: plan.wgrad(gw)}[what]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    c = counters()
    flops = 2.0 * n * h * w * cout * 9 * cin
    print(f"{what} n={n} {h}x{w} {cin}->{cout} split={split}: {us:.1f} us/launch, {flops / us / 1e6:.1f} TFLOP/s algorithmic")
    for name, v in zip(NAMES, c):
        print(f"    {name:22s} {v / reps:12.0f} cycles/launch")


if __name__ == "__main__" and os.environ.get("FB_HALO") == "1":
    run(128, 32, 32, 64, 64, True, "fwd")
    run(128, 32, 32, 64, 64, True, "dgrad")
    run(128, 16, 16, 128, 128, True, "fwd")
    run(128, 32, 32, 64, 64, True, "wgrad")
    run(128, 16, 16, 128, 128, True, "wgrad")
elif __name__ == "__main__":
    for shape in [(128, 32, 32, 64, 64), (128, 16, 16, 128, 128), (128, 8, 8, 256, 256), (128, 4, 4, 512, 512)]:
        run(*shape, True, "wgrad")
    run(128, 8, 8, 256, 256, True, "fwd")
    run(128, 4, 4, 512, 512, True, "fwd")
    run(128, 32, 32, 64, 64, True, "fwd")
    run(128, 32, 32, 64, 64, True, "dgrad")
    run(128, 16, 16, 128, 128, True, "fwd")
    run(128, 32, 32, 64, 64, False, "fwd")
