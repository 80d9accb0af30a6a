#!/usr/bin/env python
"""Benchmark of the hot path: full-batch step images/s (incl. the grad-reg pass), ResNet-18 on CIFAR-shaped data.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--precision split|bf16] [--images 50000]

A "step" is one full-batch gradient evaluation: every microbatch of the dataset goes through pass 1, the
finite-difference pass 2 and the running-mean accumulation, followed (N > 1) by the single all-reduce of the flat
buffer.  Each image is counted once although it is processed by two forward/backward passes.

  value : images/s with the dataset resident in HBM (CUDA events, max over ranks)
  e2e   : images/s through the reference-facing API (Trainer.step of fullbatchtraining_b200.training: closure + clip +
          SGD step) fed from pinned HOST memory block by block, with the per-step device->host read of the statistics
  roofline     : dominant kernel family, algorithmic FLOPs / CUDA-event time of its launches in an instrumented microbatch
  cpu_baseline : the oracle restatement of the reference (torch CPU ops, all host threads) on a bounded sample

`--impl reference` times the reference algorithm on the host cores (oracle port; /root/reference does not exist on the
GPU box), one bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GFLOP_PER_IMAGE = {18: 6.6580, 152: 44.6586}  # BASELINE.md 2 (2 passes, conv + fc, stem dgrad excluded)
HYP = dict(lr=0.8, block_strength=0.5, eps=1e-2)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p["bf16_tflops"], tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.file = index, None, None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm), reasons=sorted(reasons))


def oracle_cpu_step(images, depth=18, mb=128, threads=None):
    """One bounded sample of the reference algorithm on the host: `images` synthetic images through the oracle port."""
    from oracle import fb_oracle as O

    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    p, b = O.build_resnet_state(depth)
    X, Y = O.synthetic_cifar(images)
    t0 = time.time()
    O.full_batch_step(depth, p, b, X, Y, mb, **HYP)
    return time.time() - t0


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = args.mb  # one microbatch (pass 1 + FD pass 2 + accumulation) per step
    for _ in range(args.warmup):
        oracle_cpu_step(sample, args.depth, args.mb)
    t0 = time.time()
    for _ in range(args.steps):
        oracle_cpu_step(sample, args.depth, args.mb)
    dt = time.time() - t0
    value = sample * args.steps / dt
    line = dict(metric="full-batch step images/s (incl. grad-reg pass)", value=value, unit="images/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(args),
                cpu_baseline=dict(value=value, unit="images/s", cores=threads, kind="port",
                                  sample=f"{sample} images (1 microbatch: 2 fwd+bwd passes + accumulation) per step"),
                e2e=dict(value=value, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(args):
    return dict(workload=f"ResNet-{args.depth} full-batch GD, {args.images} synthetic CIFAR-10-shaped images, "
                         f"microbatch {args.mb}, forward-differences grad-reg (block_strength 0.5, eps 1e-2, lr 0.8)",
                images_per_step=(args.images // args.mb) * args.mb, microbatches_per_step=args.images // args.mb,
                precision=args.precision, l2="working set of a launch (several GB of activations) exceeds the 126 MB L2")


def run_ours(args):
    from fullbatchtraining_b200 import construct_model, ops
    from fullbatchtraining_b200.config import default_cfg
    from fullbatchtraining_b200.data import HostBlockLoader, synthetic_cifar
    from fullbatchtraining_b200.engine import FullBatchEngine
    from fullbatchtraining_b200.training import Trainer

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = world > 1
    if dist:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    mb, depth = args.mb, args.depth
    K = args.images // mb
    k0, k1 = (rank * K) // world, ((rank + 1) * K) // world

    # ------------------------------------------------------------------ device-resident arm (value)
    torch.manual_seed(0)
    model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
    eng = FullBatchEngine(model, mb, precision=args.precision, device=dev, groups=args.groups or None)
    X, Y = synthetic_cifar(K * mb, device=dev)

    def step():
        n = eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"], first=k0 * mb, count=k1 - k0,
                                    num_norms=K, norm_offset=k0)
        if dist:
            eng.all_reduce_mean(n, K)

    def barrier():
        if dist:
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([ms], device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms = float(t)
    ms_per_step = ms / args.steps
    value = K * mb / (ms_per_step * 1e-3)
    res = eng.results(K)

    # ------------------------------------------------------------------ launches + per-kernel roofline (instrumented)
    # One group launch (G microbatches) replayed eagerly on a single stream with a CUDA event pair around every kernel
    # launch.  The launches are hundreds of microseconds long, so event time = kernel time (no rescaling); the sum of
    # the families is reported next to the graph-replay time of the same launch.
    G = eng.G
    ng_prof = min(G, k1 - k0)
    ops.LAUNCHES["count"] = 0
    ops.PROFILE = []
    eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"], first=k0 * mb, count=ng_prof,
                            use_graph=False)
    torch.cuda.synchronize()
    launches_per_group = ops.LAUNCHES["count"]
    fam = {}
    for family, work, unit, a, b, label in ops.PROFILE:
        d = fam.setdefault(family, dict(ms=0.0, work=0.0, unit=unit, n=0))
        d["ms"] += a.elapsed_time(b)
        d["work"] += work
        d["n"] += 1
    ops.PROFILE = None
    total_ms = sum(d["ms"] for d in fam.values())
    group_launches = -(-(k1 - k0) // G)
    graph_ms_per_launch = ms_per_step / max((k1 - k0) / ng_prof, 1e-9)
    kernels = {}
    for family, d in fam.items():
        if d["work"] <= 0 or d["ms"] <= 0:
            continue
        if d["unit"] == "flop":
            ach, peak, u, bound = d["work"] / d["ms"] / 1e9, peaks["tflops_sustained"], "TFLOP/s", "tensor"
        else:
            ach, peak, u, bound = d["work"] / d["ms"] / 1e6, peaks["hbm"], "GB/s", "hbm"
        kernels[family] = dict(bound=bound, achieved=round(ach, 2), peak=peak, unit=u, frac=round(ach / peak, 4),
                               launches=d["n"], avg_launch_us=round(1e3 * d["ms"] / d["n"], 2),
                               share_of_step=round(d["ms"] / total_ms, 4))
    dominant = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    roofline = None
    if dominant:
        kd = kernels[dominant]
        roofline = dict(kernel=dominant, bound=kd["bound"], achieved=kd["achieved"], peak=kd["peak"], unit=kd["unit"],
                        frac=kd["frac"], traffic=None, peak_source=peaks["source"] + " (sustained: timed inside a step)",
                        share_of_step=kd["share_of_step"], launches_timed=kd["launches"],
                        sum_of_families_ms=round(total_ms, 3), graph_replay_ms=round(graph_ms_per_launch, 3),
                        microbatches_per_launch=ng_prof)
    step_tflops = value * GFLOP_PER_IMAGE[depth] / 1e3
    del eng, model
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ end-to-end arm (host buffers, public API)
    torch.manual_seed(0)
    model2 = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
    Xh, Yh = X[k0 * mb:k1 * mb].cpu(), Y[k0 * mb:k1 * mb].cpu()
    loader = HostBlockLoader(Xh, Yh, mb)
    cfg = default_cfg({"data.batch_size": mb, "hyp.sub_batch": mb, "hyp.warmup": 0, "hyp.steps": 10 ** 9,
                       "impl.precision": args.precision, "impl.resident_dataset": False, "impl.groups": args.groups or None,
                       "impl.setup.sharded_loader": True})
    trainer = Trainer(model2, loader, None, dict(device=dev, dtype=torch.float32), cfg)
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(max(1, min(args.warmup, 2))):
        trainer.step(validate=False)
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        trainer.step(validate=False)
    barrier()
    e2e_s = time.time() - t0
    if dist:
        t = torch.tensor([e2e_s], device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t)
    e2e_value = K * mb * e2e_steps / e2e_s
    h2d = K * mb * (3 * 32 * 32 * 4 + 8)
    d2h = 16 * 4 + max(K, 16) * 4  # step scalars + grad_norms

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = 2 * mb
        oracle_cpu_step(mb, depth, mb, threads)  # warm-up (thread pools, oneDNN primitive cache)
        dt = oracle_cpu_step(sample, depth, mb, threads)
        cpu = dict(value=sample / dt, unit="images/s", cores=threads, kind="port",
                   sample=f"{sample} images = 2 microbatches of the same workload through oracle/fb_oracle.py "
                          f"(torch {torch.__version__} CPU ops)")

    if rank == 0:
        line = dict(metric="full-batch step images/s (incl. grad-reg pass)", value=value, unit="images/s",
                    n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "bf16x2 (hi+lo)",
                    data="synthetic", config=workload_config(args), clocks=clocks,
                    e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             steps=e2e_steps, api="fullbatchtraining_b200.training.Trainer.step"),
                    gpu_launches=launches_per_group * group_launches * args.steps, roofline=roofline,
                    step_roofline=dict(bound="tensor", achieved=round(step_tflops, 2), peak=peaks["tflops"],
                                       unit="TFLOP/s", frac=round(step_tflops / peaks["tflops"], 4),
                                       note="whole step: images/s x 6.658 algorithmic GFLOP per image"),
                    kernels=kernels, cpu_baseline=cpu,
                    check=dict(loss=res["loss"], mean_grad_norm_sq=float(res["grad_norms"].mean())))
        print(json.dumps(line), flush=True)
    if dist:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="split", choices=["split", "bf16"])
    ap.add_argument("--images", type=int, default=50000)
    ap.add_argument("--mb", type=int, default=128)
    ap.add_argument("--depth", type=int, default=18)
    ap.add_argument("--groups", type=int, default=0, help="microbatches per launch (0: engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
