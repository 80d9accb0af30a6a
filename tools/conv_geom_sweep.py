"""Development aid: time fb_conv3x3 tile geometries against the generic per-tap kernel on the small-map conv shapes."""
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from fullbatchtraining_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=20, per_graph=10):
    """GPU time per call: `per_graph` calls captured in a CUDA graph (no host launch overhead inside the timed region)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * per_graph) * 1e3


def run(n, h, w, cin, cout, planes_a):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, h, w, cin, device=DEV, generator=g)
    x_hi = x.to(torch.bfloat16)
    x_lo = (x - x_hi.float()).to(torch.bfloat16)
    bf = dict(device=DEV, dtype=torch.bfloat16)
    w_hi, w_lo = torch.randn(cout, 9 * cin, **bf), torch.randn(cout, 9 * cin, **bf)
    y = torch.empty(n, h, w, cout, device=DEV)
    fk0 = [[(dhi * 3 + dwi) * cin for dhi in range(3)] for dwi in range(3)]
    pa = [x_hi, x_lo][:planes_a]
    flops = 2.0 * n * h * w * cout * 9 * cin
    res = []
    imgs = 1 if w >= 16 else 128 // (h * w)
    for halves in (1, 2):
        for nt in (64, 128):
            try:
                conv = ops.Conv3x3(pa, [w_hi, w_lo], n, h, w, cin, cout, fk0, y, False, (imgs, halves, nt))
                us = timeit(conv)
                res.append((f"halo imgs={imgs} halves={halves} nt={nt}", us))
            except RuntimeError as e:
                res.append((f"halo imgs={imgs} halves={halves} nt={nt}", float("nan")))
    # generic kernel
    tile = ops.pixel_tile(h, w)
    m_tiles = n * (h // tile[1]) if tile[2] == 1 else -(-n // tile[2])
    xs = ops.MapSet(planes_a)
    for i, t in enumerate(pa):
        ops.encode_act(xs, i, t, n, h, w, cin, tile)
    taps = [(0, kh - 1, kw - 1, (kh * 3 + kw) * cin) for kh in range(3) for kw in range(3)]
    for nt in (64, 128, 256):
        if cout % nt:
            continue
        bs = ops.MapSet(2)
        for i, t in enumerate((w_hi, w_lo)):
            ops.encode_mat(bs, i, t, 9 * cin, cout, nt)
        conv = ops.ConvGemm(xs, bs, 1, planes_a, 2, taps, cin // 64, tile, h, n, cout, y, 0, (h * w * cout, w * cout, cout),
                            False, nt)
        res.append((f"generic nt={nt}", timeit(conv)))
    chosen = ops.halo_geometry(n, h, w, 3, 1, cout, planes_a, 2)
    print(f"n={n} {h}x{w} {cin}->{cout} PA={planes_a}: model picks {chosen}, generic picks nt={ops.choose_n_tile(m_tiles, cout, planes_a, 2)}")
    for name, us in res:
        print(f"    {name:32s} {us:7.1f} us  {flops / us / 1e6:6.1f} TFLOP/s alg")


if __name__ == "__main__":
    for pa in (2, 1):
        run(128, 8, 8, 256, 256, pa)
        run(128, 4, 4, 512, 512, pa)
        run(128, 16, 16, 128, 128, pa)
        run(128, 32, 32, 64, 64, pa)
