import sys, time
sys.path.insert(0, ".")
import torch
from fullbatchtraining_b200 import ops
DEV = "cuda"
for P, Cc in [(131072, 64), (32768, 128), (8192, 256), (2048, 512), (128, 512), (512, 2048)]:
    y = torch.randn(P, Cc, device=DEV)
    gamma, beta = torch.ones(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    mean, rstd = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    hi = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16); lo = torch.empty_like(hi)
    ws = torch.zeros(2 * Cc * 1024, device=DEV)
    dA = torch.randn(P, Cc, device=DEV); dy = torch.empty_like(hi); dz = torch.empty_like(dA)
    dg, db = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    for name, fn in [("fwd_fused", lambda: ops.bn_fwd_fused(y, mean, rstd, gamma, beta, P, Cc, hi, lo, ws)),
                     ("bwd_fused", lambda: ops.bn_bwd_fused(dA, hi, y, mean, rstd, gamma, P, Cc, ws, dg, db, dy, dz_out=dz))]:
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"P={P} C={Cc} {name}: device {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/launch, host wall {(time.time() - t0) / 20 * 1e3:.2f} ms/launch", flush=True)
