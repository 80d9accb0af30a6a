"""Host-side logic of the multi-GPU path, run with 2 gloo processes on CPU (no CUDA): contiguous microbatch sharding and
the weighted all-reduce that turns per-rank running means into the exact global mean (engine.all_reduce_mean;
reference fullbatch/training/utils.py:31-41 + SURVEY.md 8e).  The CUDA scale kernel is replaced by a torch stub."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fullbatchtraining_b200 import engine as E
from fullbatchtraining_b200.training import shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _StubEngine:
    """Only the state all_reduce_mean touches."""
    all_reduce_mean = E.FullBatchEngine.all_reduce_mean
    all_reduce_flat = E.FullBatchEngine.all_reduce_flat

    def __init__(self, numel, K):
        self.numel = numel
        self.avg = torch.zeros(numel)
        self.scal = torch.zeros(E.SCAL_SLOTS)
        self.grad_norms = torch.zeros(max(K, 16))


def _worker(rank, world, port, K, numel, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E.ops.flat_scale = lambda x, n, alpha: x.mul_(alpha)  # stand-in for the CUDA kernel fb_flat_scale
    g = torch.Generator().manual_seed(7)
    grads = torch.randn(K, numel, generator=g)          # per-microbatch regularised gradients (same on all ranks)
    losses = torch.rand(K, generator=g)
    k0, k1 = shard_range(rank, world, K)
    eng = _StubEngine(numel, K)
    for j, k in enumerate(range(k0, k1)):                 # local running mean, training.py:45-47
        eng.avg += (grads[k] - eng.avg) / (j + 1)
        eng.scal[E.S_LOSS] += losses[k]
        eng.scal[E.S_CORRECT] += 1.0
        eng.grad_norms[k] = grads[k].pow(2).sum()
    eng.all_reduce_mean(k1 - k0, K)
    if rank == 0:
        torch.save(dict(avg=eng.avg, loss=eng.scal[E.S_LOSS], correct=eng.scal[E.S_CORRECT],
                        norms=eng.grad_norms[:K], ref=grads.mean(0), ref_loss=losses.sum(),
                        ref_norms=grads.pow(2).sum(1)), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("K", [7, 390])
def test_weighted_allreduce_gives_global_mean(tmp_path, K):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), K, 1000, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["avg"], r["ref"], rtol=1e-5, atol=1e-6)
    assert float(r["loss"]) == pytest.approx(float(r["ref_loss"]), rel=1e-5)
    assert float(r["correct"]) == K
    assert torch.allclose(r["norms"], r["ref_norms"], rtol=1e-6)


def test_shard_ranges_partition_the_microbatch_list():
    for K in (1, 7, 390, 3906, 1562):
        for world in (1, 2, 4, 8):
            ranges = [shard_range(r, world, K) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1   # 390 = 6*49 + 2*48 at 8 GPUs
