#!/usr/bin/env python
"""Benchmark of the hot path: full-batch step images/s (incl. the grad-reg pass) on CIFAR-shaped synthetic data.

    python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference|torch-gpu]
                    [--workload r18_50k|r18_500k|r18_2k|r152_highreg|r18_sgd] [--precision split|bf16] [--groups G]

Workloads = BASELINE.json `configs`: r18_2k (configs[0]), r18_50k (configs[1], the metric's configuration, default),
r18_500k (configs[2]), r152_highreg (configs[3]: ResNet-152, microbatch 32, per-step shuffling), r18_sgd (configs[4]: the
stochastic sanity branch, one optimizer step per block of 128; a "step" is one pass over the dataset).

A "step" of the full-batch workloads is one full-batch gradient evaluation: every microbatch of the dataset goes through
pass 1, the finite-difference pass 2 and the running-mean accumulation, followed (N > 1) by the single all-reduce of the
flat buffer.  Each image is counted once although it is processed by two forward/backward passes.

  value : images/s with the dataset resident in HBM (CUDA events, max over ranks)
  e2e   : images/s through the reference-facing API (Trainer.step of fullbatchtraining_b200.training: closure + clip +
          SGD step) fed from pinned HOST memory block by block, with the per-step device->host read of the statistics
  roofline / kernels : one group launch replayed eagerly with a CUDA-event pair around every kernel launch (the launches
          are 0.1-0.6 ms long, so no rescaling): algorithmic FLOPs or bytes / event time per kernel family;
          `traffic` = DRAM bytes per launch of the dominant family from the committed ncu capture (profiles/)
  check : loss and mean squared gradient norm of the step against committed values of the fp32 oracle
          (tests/golden/bench_check.json, written by tools/make_bench_check.py)
  cpu_baseline : the oracle restatement of the reference (torch CPU ops, all host threads) on a bounded sample

`--impl reference` times the reference algorithm on the host cores (oracle port; /root/reference does not exist on the
GPU box), one bounded sample per step.  `--impl torch-gpu` times the same restatement with stock PyTorch CUDA ops
(cuDNN / cuBLAS + autograd: what the reference executes on a GPU) in fp32 and with TF32 convolutions on a bounded sample:
the library path this framework replaces.
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
WORKLOADS = {
    "r18_2k": dict(depth=18, images=2000, mb=128, shuffle=False, mode="fullbatch"),
    "r18_50k": dict(depth=18, images=50000, mb=128, shuffle=False, mode="fullbatch"),
    "r18_500k": dict(depth=18, images=500000, mb=128, shuffle=False, mode="fullbatch"),
    "r152_highreg": dict(depth=152, images=50000, mb=32, shuffle=True, mode="fullbatch"),
    "r18_sgd": dict(depth=18, images=50000, mb=128, shuffle=True, mode="sgd"),
}
METRIC = "full-batch step images/s (incl. grad-reg pass)"


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


def workload_config(args):
    w = args.w
    passes = "one pass per block + SGD step (stochastic sanity branch)" if w["mode"] == "sgd" else \
        "forward-differences grad-reg (block_strength 0.5, eps 1e-2, lr 0.8)"
    return dict(workload=f"{args.workload}: ResNet-{w['depth']} {'SGD' if w['mode'] == 'sgd' else 'full-batch GD'}, "
                         f"{w['images']} synthetic CIFAR-10-shaped images, microbatch {w['mb']}, {passes}"
                         f"{', per-step shuffling' if w['shuffle'] else ''}",
                images_per_step=(w["images"] // w["mb"]) * w["mb"], microbatches_per_step=w["images"] // w["mb"],
                precision=args.precision, l2="working set of a launch (several GB of activations) exceeds the 126 MB L2")


# ---------------------------------------------------------------------------------------------------------------------
# baseline arms (the only places that execute oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def oracle_step(images, depth, mb, device="cpu", threads=None, block_strength=None):
    """One bounded sample of the reference algorithm: `images` synthetic images through the oracle port."""
    from oracle import fb_oracle as O

    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    p, b = O.build_resnet_state(depth)
    X, Y = O.synthetic_cifar(images)
    if device != "cpu":
        p = {k: v.to(device) for k, v in p.items()}
        b = {k: v.to(device) for k, v in b.items()}
        X, Y = X.to(device), Y.to(device)
        torch.cuda.synchronize()
    hyp = dict(HYP)
    if block_strength is not None:
        hyp["block_strength"] = block_strength
    t0 = time.time()
    O.full_batch_step(depth, p, b, X, Y, mb, **hyp)
    if device != "cpu":
        torch.cuda.synchronize()
    return time.time() - t0


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port), all host threads.  Each step is a
    bounded sample of the workload: ONE microbatch (both passes + accumulation; one pass for the SGD workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = args.w
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = w["mb"]
    bs = 0.0 if w["mode"] == "sgd" else None
    for _ in range(args.warmup):
        oracle_step(sample, w["depth"], w["mb"], block_strength=bs)
    t0 = time.time()
    for _ in range(args.steps):
        oracle_step(sample, w["depth"], w["mb"], block_strength=bs)
    dt = time.time() - t0
    value = sample * args.steps / dt
    cfg = workload_config(args)
    cfg["sample_per_step"] = f"{sample} images = 1 microbatch of the workload (the CPU needs minutes for the whole set)"
    line = dict(metric=METRIC, value=value, unit="images/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic", config=cfg,
                cpu_baseline=dict(value=value, unit="images/s", cores=threads, kind="port",
                                  sample=f"{sample} images (1 microbatch: fwd+bwd passes + accumulation) per step"),
                e2e=dict(value=value, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    """Stock PyTorch on the same B200: the torch-op restatement of the path with cuDNN / cuBLAS convolutions and
    autograd (what the reference itself executes on a GPU), fp32 and TF32, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("--impl torch-gpu needs a GPU")
    w = args.w
    sample = min(w["images"] // w["mb"], 24) * w["mb"]
    bs = 0.0 if w["mode"] == "sgd" else None
    out = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        for _ in range(max(args.warmup, 1)):
            oracle_step(2 * w["mb"], w["depth"], w["mb"], device="cuda", block_strength=bs)
        t = [oracle_step(sample, w["depth"], w["mb"], device="cuda", block_strength=bs) for _ in range(args.steps)]
        out[name] = sample / (sum(t) / len(t))
    cfg = workload_config(args)
    cfg["sample_per_step"] = f"{sample} images ({sample // w['mb']} microbatches) of the workload"
    line = dict(metric=METRIC, value=out["fp32"], unit="images/s", impl="torch-gpu", n_gpus=1, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * sample / out["fp32"], higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f32", data="synthetic", config=cfg,
                torch_gpu=dict(fp32=out["fp32"], tf32_convolutions=out["tf32"], unit="images/s", torch=torch.__version__,
                               cudnn=torch.backends.cudnn.version()))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def load_traffic(args, family):
    """DRAM bytes per launch of a kernel family from the committed ncu capture of this workload (None if absent)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    # r18_2k / r18_500k replay the very same group launch (8 microbatches of 128) as r18_50k
    workload = "r18_50k" if args.workload in ("r18_2k", "r18_500k") else args.workload
    entry = t.get(f"{workload}:{args.precision}", {}).get(family)
    return (entry["dram_bytes_per_launch"], t.get("source")) if entry else (None, None)


def expected_check(args):
    path = os.path.join(ROOT, "tests", "golden", "bench_check.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(args.workload)


class DeviceBlockLoader:
    """The DataLoader protocol of the stochastic branch over a device-resident dataset (value arm of r18_sgd)."""

    def __init__(self, X, Y, mb, shuffle, seed=0):
        self.X, self.Y, self.mb, self.shuffle = X, Y, mb, shuffle
        self.gen = torch.Generator(device=X.device).manual_seed(seed)
        self.sampler = self
        self.batch_size = mb

    def set_epoch(self, *a, **k):
        pass

    def __len__(self):
        return self.X.shape[0] // self.mb

    def __iter__(self):
        n = len(self) * self.mb
        idx = torch.randperm(self.X.shape[0], device=self.X.device, generator=self.gen)[:n] if self.shuffle else None
        for i in range(len(self)):
            sl = slice(i * self.mb, (i + 1) * self.mb)
            yield (self.X[sl], self.Y[sl]) if idx is None else (self.X[idx[sl]], self.Y[idx[sl]])


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
    w = args.w
    mb, depth, sgd = w["mb"], w["depth"], w["mode"] == "sgd"
    K = w["images"] // mb
    k0, k1 = (rank * K) // world, ((rank + 1) * K) // world
    setup = dict(device=dev, dtype=torch.float32)

    def barrier():
        if dist:
            td.barrier()
        torch.cuda.synchronize()

    def cfg_for(**extra):
        o = {"data.batch_size": mb, "hyp.sub_batch": mb, "hyp.warmup": 0, "hyp.steps": 10 ** 9,
             "impl.precision": args.precision, "impl.groups": args.groups or None, "hyp.shuffle": w["shuffle"],
             "impl.setup.sharded_loader": True}
        if sgd:  # hyp=base_sgd: stochastic branch, no regulariser, lr small enough for synthetic noise data
            o.update({"hyp.train_stochastic": True, "hyp.grad_reg.block_strength": 0.0, "hyp.optim.lr": 0.05})
        o.update(extra)
        return default_cfg(o)

    # ------------------------------------------------------------------ device-resident arm (value)
    torch.manual_seed(0)
    model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
    X, Y = synthetic_cifar(K * mb, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    if sgd:
        trainer0 = Trainer(model, DeviceBlockLoader(X[k0 * mb:k1 * mb], Y[k0 * mb:k1 * mb], mb, w["shuffle"]), None, setup,
                           cfg_for())
        eng = trainer0.engine

        def step():
            trainer0.step(validate=False)
    else:
        eng = FullBatchEngine(model, mb, precision=args.precision, device=dev, groups=args.groups or None)
        perm = torch.zeros(K * mb, device=dev, dtype=torch.int64) if w["shuffle"] else None

        def step():
            if perm is not None:  # hyp.shuffle: a fresh device permutation per step (data_preparation.py:53-54)
                perm.copy_(torch.randperm(K * mb, device=dev, generator=gen))
            n = eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"], first=k0 * mb, count=k1 - k0,
                                        num_norms=K, norm_offset=k0, perm=perm)
            if dist:
                eng.all_reduce_mean(n, K)

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
    if sgd:
        check = dict(loss=trainer0.stats["train_loss"][-1], first_loss=trainer0.stats["train_loss"][0])
    else:
        res = eng.results(K)
        check = dict(loss=res["loss"], mean_grad_norm_sq=float(res["grad_norms"].mean()))
        exp = expected_check(args) if not w["shuffle"] else None
        if exp:  # the fp32 oracle's loss / mean squared gradient norm of the same step (committed fixture)
            check["expected"] = exp
            check["ok"] = bool(abs(check["loss"] - exp["loss"]) <= 1e-4 * abs(exp["loss"]) and
                               abs(check["mean_grad_norm_sq"] - exp["mean_grad_norm_sq"]) <= 5e-3 * exp["mean_grad_norm_sq"])

    # ------------------------------------------------------------------ launches + per-kernel roofline (instrumented)
    # One group launch (G microbatches; the SGD workload: one block) replayed eagerly on a single stream with a CUDA
    # event pair around every kernel launch.  The launches are hundreds of microseconds long, so event time = kernel
    # time (no rescaling); the sum of the families is reported next to the graph-replay time of the same launch.
    G = eng.G
    ng_prof = min(G, k1 - k0)
    ops.LAUNCHES["count"] = 0
    ops.PROFILE = []
    eng.accumulate_resident(X, Y, HYP["lr"], 0.0 if sgd else HYP["block_strength"], HYP["eps"], first=k0 * mb,
                            count=ng_prof, use_graph=False)
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
        tr, _ = load_traffic(args, family)
        if tr is not None:
            kernels[family]["traffic"] = tr
    dominant = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    roofline = None
    if dominant:
        kd = kernels[dominant]
        traffic, tsrc = load_traffic(args, dominant)
        roofline = dict(kernel=dominant, bound=kd["bound"], achieved=kd["achieved"], peak=kd["peak"], unit=kd["unit"],
                        frac=kd["frac"], traffic=traffic, traffic_source=tsrc,
                        peak_source=peaks["source"] + " (sustained: timed inside a step)",
                        share_of_step=kd["share_of_step"], launches_timed=kd["launches"],
                        sum_of_families_ms=round(total_ms, 3), graph_replay_ms=round(graph_ms_per_launch, 3),
                        microbatches_per_launch=ng_prof)
    gflop = GFLOP_PER_IMAGE[depth] / (2.0 if sgd else 1.0)
    step_tflops = value * gflop / 1e3
    if sgd:
        del trainer0
    engine_info = dict(microbatches_per_launch=eng.G, lanes=len(eng.lanes),
                       device_memory_gib=round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1))
    del eng, model
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ end-to-end arm (host buffers, public API)
    torch.manual_seed(0)
    model2 = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
    Xh, Yh = X[k0 * mb:k1 * mb].cpu(), Y[k0 * mb:k1 * mb].cpu()
    del X, Y
    torch.cuda.empty_cache()
    loader = HostBlockLoader(Xh, Yh, mb)
    # a host-side loader is this rank's shard; shuffling a streamed loader is the loader's business (sequential here)
    trainer = Trainer(model2, loader, None, setup, cfg_for(**{"impl.resident_dataset": False,
                                                              "impl.setup.sharded_loader": True, "hyp.shuffle": False}))
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
        sample = (2 if depth < 100 else 1) * mb
        bs = 0.0 if sgd else None
        oracle_step(mb, depth, mb, threads=threads, block_strength=bs)  # warm-up (thread pools, primitive cache)
        dt = oracle_step(sample, depth, mb, threads=threads, block_strength=bs)
        cpu = dict(value=sample / dt, unit="images/s", cores=threads, kind="port",
                   sample=f"{sample} images = {sample // mb} microbatch(es) of the same workload through "
                          f"oracle/fb_oracle.py (torch {torch.__version__} CPU ops)")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="images/s",
                    n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "bf16x2 (hi+lo)",
                    data="synthetic", config=dict(workload_config(args), **engine_info),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             steps=e2e_steps, api="fullbatchtraining_b200.training.Trainer.step"),
                    gpu_launches=launches_per_group * group_launches * args.steps, roofline=roofline,
                    step_roofline=dict(bound="tensor", achieved=round(step_tflops, 2), peak=peaks["tflops"],
                                       unit="TFLOP/s", frac=round(step_tflops / peaks["tflops"], 4),
                                       frac_of_sustained=round(step_tflops / peaks["tflops_sustained"], 4),
                                       note=f"whole step: images/s x {gflop:.3f} algorithmic GFLOP per image"),
                    kernels=kernels, cpu_baseline=cpu, check=check)
        print(json.dumps(line), flush=True)
    if dist:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--workload", default="r18_50k", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="split", choices=["split", "bf16"])
    ap.add_argument("--images", type=int, default=None, help="override the workload's dataset size")
    ap.add_argument("--mb", type=int, default=None, help="override the workload's microbatch size")
    ap.add_argument("--depth", type=int, default=None, help="override the workload's ResNet depth")
    ap.add_argument("--groups", type=int, default=0, help="microbatches per launch (0: engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.w = dict(WORKLOADS[args.workload])
    for key in ("images", "mb", "depth"):
        if getattr(args, key) is not None:
            args.w[key] = getattr(args, key)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
