"""Quick device timing of the microbatch graph (development aid; bench.py is the contract benchmark)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    for precision in ("split", "bf16"):
        torch.manual_seed(0)
        model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
        eng = FullBatchEngine(model, mb, precision=precision)
        g = torch.Generator(device="cuda").manual_seed(1)
        X = torch.randn(K * mb, 3, 32, 32, device="cuda", generator=g)
        Y = torch.randint(0, 10, (K * mb,), device="cuda", generator=g)
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
            print(f"{precision} depth={depth} mb={mb} K={K}: {ms / K:.3f} ms/microbatch, {K * mb / ms * 1e3:.0f} img/s "
                  f"(host {1e3 * (time.time() - t0) / K:.3f} ms/mb) loss={eng.results(K)['loss']:.4f}", flush=True)
        del eng, model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
