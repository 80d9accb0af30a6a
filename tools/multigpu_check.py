"""Multi-GPU parity check (run with torchrun, one rank per GPU): the sharded + all-reduced full-batch gradient must
equal the single-GPU result up to fp32 summation order.  Writes gpurun_out/multigpu_check.json on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.data import synthetic_cifar  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402
from fullbatchtraining_b200.training import shard_range  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mb, n = 128, 128 * 13
K = n // mb
torch.manual_seed(0)
model = construct_model(dict(name="ResNet18", depth=18), 3, 10)
eng = FullBatchEngine(model, mb, device=dev)
X, Y = synthetic_cifar(n, device=dev)
k0, k1 = shard_range(rank, world, K)
cnt = eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, first=k0 * mb, count=k1 - k0, num_norms=K, norm_offset=k0)
eng.all_reduce_mean(cnt, K)
sharded = eng.avg.clone()
res = eng.results(K)
# single-GPU reference on every rank (BN running statistics differ afterwards; irrelevant for the gradient)
for name, b in model.named_buffers():
    if name.endswith("running_mean"):
        b.zero_()
    elif name.endswith("running_var"):
        b.fill_(1)
eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2)
single = eng.avg
err = float((sharded - single).norm() / single.norm())
res1 = eng.results(K)
if rank == 0:
    out = dict(world=world, microbatches=K, rel_err_vs_single_gpu=err, loss_sharded=res["loss"], loss_single=res1["loss"],
               grad_norms_equal=bool(torch.equal(res["grad_norms"], res1["grad_norms"])))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/multigpu_check.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    assert err < 1e-5, err
dist.destroy_process_group()
