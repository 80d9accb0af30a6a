"""Quick device timing of the group-launch graph (development aid; bench.py is the contract benchmark).

    python tools/quick_bench.py [depth] [mb] [K] [groups,comma,separated] [precision]

Prints images/s of a resident accumulation over K microbatches for every group count, and for the last one the
per-family time of ONE eager group launch (CUDA events around every kernel launch, single stream).
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from fullbatchtraining_b200 import construct_model, ops  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402


def families(eng, X, Y):
    ops.PROFILE = []
    eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, count=eng.G, use_graph=False)
    torch.cuda.synchronize()
    fam, detail = {}, {}
    for family, work, unit, a, b, label in ops.PROFILE:
        d = fam.setdefault(family, [0.0, 0.0, unit, 0])
        ms = a.elapsed_time(b)
        dd = detail.setdefault((family, label), [0.0, 0.0, unit, 0])
        dd[0] += ms
        dd[1] += work
        dd[3] += 1
        d[0] += ms
        d[1] += work
        d[3] += 1
    ops.PROFILE = None
    total = sum(d[0] for d in fam.values())
    imgs = eng.G * eng.mb
    print(f"  one eager launch of {eng.G} microbatches: {total:.2f} ms = {1e3 * total / imgs:.2f} us/image")
    for family, (ms, work, unit, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        rate = work / ms / 1e9 if unit == "flop" else work / ms / 1e6
        print(f"  {family:14s} {ms:8.3f} ms {100 * ms / total:5.1f}%  x{n:4d}  {1e3 * ms / imgs:6.2f} us/img  "
              f"{rate:9.1f} {'TFLOP/s' if unit == 'flop' else 'GB/s'}")
    print("  --- by launch shape")
    for (family, label), (ms, work, unit, n) in sorted(detail.items(), key=lambda kv: -kv[1][0])[:40]:
        rate = work / ms / 1e9 if unit == "flop" else work / ms / 1e6
        print(f"  {ms:8.3f} ms x{n:3d} avg {1e3 * ms / n:8.1f} us {rate:9.1f} {'TFLOP/s' if unit == 'flop' else 'GB/s'}  {label}")


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    groups = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    precision = sys.argv[5] if len(sys.argv) > 5 else "split"
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(K * mb, 3, 32, 32, device="cuda", generator=g)
    Y = torch.randint(0, 10, (K * mb,), device="cuda", generator=g)
    for G in groups:
        torch.manual_seed(0)
        model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
        eng = FullBatchEngine(model, mb, precision=precision, groups=G or None)
        eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2)
        torch.cuda.synchronize()
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            e0.record()
            eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"{precision} depth={depth} mb={mb} K={K} G={eng.G}: {ms / K:.3f} ms/microbatch, "
                  f"{K * mb / ms * 1e3:.0f} img/s (host {1e3 * (time.time() - t0) / K:.3f} ms/mb) "
                  f"loss={eng.results(K)['loss']:.4f} mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        if G == groups[-1]:
            families(eng, X, Y)
        del eng, model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
