"""Development aid: where does the generic conv kernel's time go?  Times one launch (CUDA-graph timing) with parts of the
kernel switched off through FB_CONV_EXPERIMENT (1 = no global stores, 2 = no epilogue work, 4 = no operand re-fetch)."""
import os
import sys

os.environ["FB_KERNEL_DEBUG"] = "1"  # the experiment switch is only honoured in debug mode
sys.path.insert(0, ".")
import torch  # noqa: E402

from fullbatchtraining_b200 import ops  # noqa: E402
from tools.conv_geom_sweep import timeit  # noqa: E402

DEV = "cuda"


def run(n, h, w, cin, cout, planes_a, nt):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, h, w, cin, device=DEV, generator=g)
    x_hi = x.to(torch.bfloat16)
    x_lo = (x - x_hi.float()).to(torch.bfloat16)
    bf = dict(device=DEV, dtype=torch.bfloat16)
    w_hi, w_lo = torch.randn(cout, 9 * cin, **bf), torch.randn(cout, 9 * cin, **bf)
    y = torch.empty(n, h, w, cout, device=DEV)
    pa = [x_hi, x_lo][:planes_a]
    tile = ops.pixel_tile(h, w)
    xs = ops.MapSet(planes_a)
    for i, t in enumerate(pa):
        ops.encode_act(xs, i, t, n, h, w, cin, tile)
    taps = [(0, kh - 1, kw - 1, (kh * 3 + kw) * cin) for kh in range(3) for kw in range(3)]
    bs = ops.MapSet(2)
    for i, t in enumerate((w_hi, w_lo)):
        ops.encode_mat(bs, i, t, 9 * cin, cout, nt)
    conv = ops.ConvGemm(xs, bs, 1, planes_a, 2, taps, cin // 64, tile, h, n, cout, y, 0, (h * w * cout, w * cout, cout),
                        False, nt)
    out = []
    for exp in (0, 1, 2, 4, 5, 6):
        os.environ["FB_CONV_EXPERIMENT"] = str(exp)
        out.append(f"exp{exp}: {timeit(conv):6.1f}")
    os.environ["FB_CONV_EXPERIMENT"] = "0"
    print(f"n={n} {h}x{w} {cin}->{cout} PA={planes_a} nt={nt}:  " + "  ".join(out) + "  (us)")


if __name__ == "__main__":
    for pa in (2, 1):
        run(128, 32, 32, 64, 64, pa, 64)
        run(128, 16, 16, 128, 128, pa, 128)
        run(128, 8, 8, 256, 256, pa, 128)
        run(128, 4, 4, 512, 512, pa, 64)
