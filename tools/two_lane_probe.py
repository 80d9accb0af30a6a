"""Probe: aggregate throughput of TWO independent engines replaying their microbatch graphs concurrently on two
streams of one GPU (upper bound of what pipelining two microbatches inside one engine could give)."""
import sys
import threading

import torch

sys.path.insert(0, ".")
from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 32
engines, data, streams = [], [], []
for i in range(2):
    torch.manual_seed(i)
    eng = FullBatchEngine(construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10), mb)
    g = torch.Generator(device="cuda").manual_seed(i)
    X = torch.randn(K * mb, 3, 32, 32, device="cuda", generator=g)
    Y = torch.randint(0, 10, (K * mb,), device="cuda", generator=g)
    eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2)  # capture + warm
    engines.append(eng); data.append((X, Y)); streams.append(torch.cuda.Stream())
torch.cuda.synchronize()


def timed(active):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for i in active:
        with torch.cuda.stream(streams[i]):
            engines[i].accumulate_resident(*data[i], 0.8, 0.5, 1e-2)
    for i in active:
        torch.cuda.current_stream().wait_stream(streams[i])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for _ in range(2):
    t1 = timed([0])
    t2 = timed([0, 1])
    print(f"depth={depth} mb={mb}: one lane {K * mb / t1 * 1e3:.0f} img/s ({t1 / K:.3f} ms/mb); two concurrent lanes "
          f"{2 * K * mb / t2 * 1e3:.0f} img/s aggregate ({t2 / K:.3f} ms per pair) -> x{2 * t1 / t2:.2f}", flush=True)
