"""Profile target: one warm microbatch, then one microbatch inside cudaProfilerStart/Stop (use with
`ncu --profile-from-start off ...`).  argv: depth mb precision [graph]"""
import sys

import torch

sys.path.insert(0, ".")
from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
precision = sys.argv[3] if len(sys.argv) > 3 else "split"
torch.manual_seed(0)
model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
eng = FullBatchEngine(model, mb, precision=precision)
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.randn(2 * mb, 3, 32, 32, device="cuda", generator=g)
Y = torch.randint(0, 10, (2 * mb,), device="cuda", generator=g)
eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, count=1, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, count=1, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one microbatch", eng.results(1)["loss"])
