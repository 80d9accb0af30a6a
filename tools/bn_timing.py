"""Development aid: CUDA-graph timing of the fused BatchNorm backward kernel (grid-barrier vs small-map cluster variant,
FB_BN_SLICED_MAX selects) with operands either L2-warm or evicted before every launch."""
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from fullbatchtraining_b200 import ops  # noqa: E402

DEV = "cuda"


def graph_time(fn, flush=None, reps=10, per_graph=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            if flush is not None:
                flush.zero_()
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


flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
t_flush = graph_time(lambda: None, flush)
print(f"flush alone: {t_flush:.1f} us")
for P, Cc in [(131072, 64), (32768, 128), (8192, 256), (2048, 512)]:
    y = torch.randn(P, Cc, device=DEV)
    gamma = torch.ones(Cc, device=DEV)
    mean, rstd = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    hi = torch.randn(P, Cc, device=DEV).to(torch.bfloat16)
    ws = torch.zeros(2 * Cc * 1024, device=DEV)
    dA = torch.randn(P, Cc, device=DEV)
    dy = torch.empty_like(hi)
    dz = torch.empty_like(dA)
    dg, db = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    fn = lambda: ops.bn_bwd_fused(dA, hi, y, mean, rstd, gamma, P, Cc, ws, dg, db, dy, dz_out=dz)  # noqa: E731
    warm = graph_time(fn)
    cold = graph_time(fn, flush) - t_flush
    print(f"P={P} C={Cc} bwd_fused (FB_BN_SLICED_MAX={os.environ.get('FB_BN_SLICED_MAX', 'default')}): "
          f"L2-warm {warm:.1f} us, evicted {cold:.1f} us", flush=True)
