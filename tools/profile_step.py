"""Profile target: one warm group launch, then one group launch inside cudaProfilerStart/Stop (use with
`ncu --profile-from-start off ...`).  argv: depth mb precision [groups]"""
import sys

import torch

sys.path.insert(0, ".")
from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
precision = sys.argv[3] if len(sys.argv) > 3 else "split"
groups = int(sys.argv[4]) if len(sys.argv) > 4 else None
torch.manual_seed(0)
model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
eng = FullBatchEngine(model, mb, precision=precision, groups=groups)
g = torch.Generator(device="cuda").manual_seed(1)
n = eng.G * mb
X = torch.randn(n, 3, 32, 32, device="cuda", generator=g)
Y = torch.randint(0, 10, (n,), device="cuda", generator=g)
eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled one group launch of {eng.G} microbatches", eng.results(eng.G)["loss"])
