"""Reference values for bench.py's `check` (run on the B200 box: the synthetic data of bench.py comes from the CUDA
generator):  python tools/make_bench_check.py [workload ...]  ->  tests/golden/bench_check.json

For each workload the ORACLE (oracle/fb_oracle.py, torch ops, fp32 with TF32 disabled = the reference's own arithmetic,
and fp64 where it is affordable) evaluates pass 1 of the full-batch step on bench.py's data and initialisation: the mean
loss over the microbatches (`stats["train_loss"]`, training.py:185) and the mean squared raw gradient norm
(training.py:162,93).  Both are independent of the regulariser, so one pass suffices.  bench.py compares its own line
against these numbers (`check.ok`) and tests/test_engine_gpu.py asserts it for the 50k configuration."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from fullbatchtraining_b200.data import synthetic_cifar  # noqa: E402
from oracle import fb_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_check.json")


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    names = sys.argv[1:] or ["r18_2k", "r18_50k"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        w = WORKLOADS[name]
        K = w["images"] // w["mb"]
        X, Y = synthetic_cifar(K * w["mb"], device=dev)
        entry = dict(images=K * w["mb"], microbatches=K, torch=torch.__version__, gpu=torch.cuda.get_device_name(0))
        for dt, key in ((torch.float32, ""), (torch.float64, "_fp64")):
            if dt == torch.float64 and K > 400:
                continue
            torch.manual_seed(0)
            p, b = O.build_resnet_state(w["depth"], dtype=dt)
            p = {k: v.to(dev) for k, v in p.items()}
            b = {k: v.to(dev) for k, v in b.items()}
            t0 = time.time()
            out = O.full_batch_step(w["depth"], p, b, X.to(dt), Y, w["mb"], lr=0.8, block_strength=0.0)
            entry["loss" + key] = float(out["loss"])
            entry["mean_grad_norm_sq" + key] = float(out["grad_norms"].double().mean())
            entry["seconds" + key] = round(time.time() - t0, 1)
        res[name] = entry
        print(name, entry, flush=True)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)  # travels back from the GPU box
    with open(os.path.join(ROOT, "gpurun_out", "bench_check.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
