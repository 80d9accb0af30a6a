"""Step-level parity on the B200: the CUDA engine against the oracle restatement (oracle/fb_oracle.py) on identical
seeded inputs and initialisation.

Tolerance design: the fp32 reference itself is only ~2e-3 (raw gradient) / ~5e-2 (regularised
gradient) accurate per microbatch against an fp64 run at initialisation, because the finite difference divides rounding
noise by eps_n.  Ground truth is therefore the oracle in fp64; the oracle in fp32 (TF32 disabled) gives the noise floor
e32, and the engine's error e_new must satisfy  e_new <= RATIO * max(e32, FLOOR).  The measured values are written to
gpurun_out/parity_*.json so DESIGN.md can quote them.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.engine import FullBatchEngine  # noqa: E402
from oracle import fb_oracle as O  # noqa: E402

DEV = torch.device("cuda")
HYP = dict(lr=0.8, block_strength=0.5, eps=1e-2)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")

# e_new <= RATIO * max(e32, FLOOR): split mode carries ~16 mantissa bits per tensor-core operand (8 for the output
# gradient) vs 24 in fp32.  Measured on B200 (profiles/r2_parity_*.json): raw 2.2-4.2x the fp32 reference's own error,
# regularised 1.1-2.4x, accumulated 1.1-2.2x.  The bounds are the ones VERDICT.md (round 1) asked for, and the absolute
# caps state the accuracy actually delivered for ResNet-18 at initialisation.  The regularised gradient of a SINGLE
# microbatch gets more room: its fp32 reference error is itself noisy from box to box (cuDNN picks: 0.039 / 0.047 on two
# B200s for the 16-image case, i.e. ratios 2.9 / 2.4 for the same engine bits).
RATIO_RAW, RATIO_REG, RATIO_REG_MB = 6.0, 2.5, 3.5
FLOOR_RAW, FLOOR_REG = 1e-3, 2e-2
ABS_RAW_R18, ABS_REG_R18 = 2e-2, 0.2


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def cos(a, b):
    a, b = a.double(), b.double()
    return float((a * b).sum() / (a.norm() * b.norm()))


def oracle_run(depth, params, buffers, X, Y, mb, dtype, keep=2, **kw):
    p = {k: v.to(DEV, dtype).clone() for k, v in params.items()}
    b = {k: (v.to(DEV).clone() if v.dtype == torch.long else v.to(DEV, dtype).clone()) for k, v in buffers.items()}
    out = O.full_batch_step(depth, p, b, X.to(DEV, dtype), Y.to(DEV), mb, keep_microbatches=keep, **dict(HYP, **kw))
    return out, b


def setup_case(depth, mb, n, data="randn"):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)
    params = {k: v.detach().clone() for k, v in model.named_parameters()}
    buffers = {k: v.detach().clone() for k, v in model.named_buffers()}
    X, Y = O.synthetic_cifar(n) if data == "randn" else O.structured_cifar(n)
    return model, params, buffers, X.to(DEV), Y.to(DEV)


def dump(name, d):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"parity_{name}.json"), "w") as f:
        json.dump(d, f, indent=1)
    print(name, json.dumps(d))


def compare_step(name, depth, mb, model, params, buffers, X, Y, groups=None, per_microbatch=True, precision="split"):
    """One full-batch step of the engine against the oracle in fp64 (truth) and fp32 (the reference's own arithmetic);
    returns the report that is also written to gpurun_out/parity_<name>.json."""
    n = X.shape[0]
    ref64, buf64 = oracle_run(depth, params, buffers, X, Y, mb, torch.float64)
    ref32, _ = oracle_run(depth, params, buffers, X, Y, mb, torch.float32)
    eng = FullBatchEngine(model, mb, precision=precision, groups=groups)
    theta0 = eng.theta.clone()
    K = eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"])
    res = eng.results(K)
    assert K == n // mb
    assert torch.equal(eng.theta, theta0), "parameters must be unchanged by the regulariser"
    avg64, avg32 = O.flat(ref64["avg"]), O.flat(ref32["avg"])
    rep = dict(K=K, groups=eng.G, e_new_avg=rel(eng.avg, avg64), e32_avg=rel(avg32, avg64),
               e_new_avg_vs_fp32=rel(eng.avg, avg32), cos_avg=cos(eng.avg, avg64), cos_avg_vs_fp32=cos(eng.avg, avg32),
               loss=res["loss"], loss64=float(ref64["loss"]), loss32=float(ref32["loss"]),
               grad_norms=res["grad_norms"].tolist(), grad_norms64=ref64["grad_norms"].tolist())
    if per_microbatch:
        # raw / regularised gradient of microbatch 0 through the GradRegularizer protocol of the engine
        eng.microbatch_gradient(X[:mb], Y[:mb])
        raw = eng.g.clone()
        eng.regularize(X[:mb], Y[:mb], HYP["lr"], HYP["block_strength"], HYP["eps"])
        reg = eng.g.clone()
        k64, k32 = ref64["kept"][0], ref32["kept"][0]
        rep.update(e_new_raw=rel(raw, O.flat(k64["raw"])), e32_raw=rel(O.flat(k32["raw"]), O.flat(k64["raw"])),
                   e_new_raw_vs_fp32=rel(raw, O.flat(k32["raw"])),
                   e_new_reg=rel(reg, O.flat(k64["reg"])), e32_reg=rel(O.flat(k32["reg"]), O.flat(k64["reg"])),
                   e_new_reg_vs_fp32=rel(reg, O.flat(k32["reg"])),
                   cos_raw=cos(raw, O.flat(k64["raw"])), cos_reg=cos(reg, O.flat(k64["reg"])))
    l64, l32 = float(ref64["loss"]), float(ref32["loss"])
    gn64, gn32 = ref64["grad_norms"].double().cpu(), ref32["grad_norms"].double().cpu()
    rep.update(e32_loss=abs(l32 - l64) / abs(l64), e32_grad_norms=float(((gn32 - gn64).abs() / gn64).max()),
               e_new_loss=abs(res["loss"] - l64) / abs(l64),
               e_new_grad_norms=float(((res["grad_norms"].double() - gn64).abs() / gn64).max()),
               correct=res["correct"], correct64=float(ref64["correct"]))
    dump(name, rep)
    return rep, eng, ref64, buf64


def assert_step(rep, depth):
    # loss: forward pass with ~16-bit operands; 1e-4 for the 20 convs of ResNet-18, 1e-3 for the 155 of ResNet-152
    assert rep["e_new_loss"] <= max(1e-4 if depth < 100 else 1e-3, RATIO_RAW * rep["e32_loss"])
    assert rep["e_new_grad_norms"] <= max(2e-2 if depth > 100 else 2e-3, RATIO_RAW * rep["e32_grad_norms"])
    assert rep["correct"] == rep["correct64"]
    if "e_new_raw" in rep:
        assert rep["e_new_raw"] <= RATIO_RAW * max(rep["e32_raw"], FLOOR_RAW)
        assert rep["e_new_reg"] <= RATIO_REG_MB * max(rep["e32_reg"], FLOOR_REG)
        if depth < 100:
            assert rep["e_new_raw"] <= ABS_RAW_R18 and rep["e_new_reg"] <= ABS_REG_R18
            assert rep["cos_raw"] > 0.9995 and rep["cos_reg"] > 0.98
    assert rep["e_new_avg"] <= RATIO_REG * max(rep["e32_avg"], FLOOR_REG)
    if depth < 100:
        assert rep["e_new_avg"] <= ABS_REG_R18
    if rep["e32_avg"] < 0.2:  # 20 convolutions (ResNet-18) / 155 (ResNet-152) deep
        assert rep["cos_avg"] > (0.98 if depth < 100 else 0.97)


@pytest.mark.parametrize("depth,mb,n,groups", [(18, 16, 32, None), (18, 128, 256, None), (152, 32, 64, 2)],
                         ids=["r18_mb16", "r18_mb128", "r152_mb32"])
def test_full_batch_step_matches_oracle(depth, mb, n, groups):
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    rep, _, _, _ = compare_step(f"r{depth}_mb{mb}_n{n}", depth, mb, model, params, buffers, X, Y, groups=groups)
    assert_step(rep, depth)


def test_numerics_ablation_plain_bf16_against_split_operands():
    """The numerics decision of DESIGN.md section 4 as a measurement: the same step with single bf16 operands (one product
    per MAC, `precision="bf16"`) and with hi/lo split operands (3 / 2 / 2 products), each against the fp64 oracle, next
    to the fp32 oracle's own error.  Plain bf16 fails the parity bars of this file by an order of magnitude on the
    finite-difference term; the split operands pass them.  Written to gpurun_out/parity_ablation_r18_mb128_n256.json."""
    out = {}
    for precision in ("bf16", "split_w", "split"):  # split_w: forward without the x_lo * w_hi product (2 per MAC)
        model, params, buffers, X, Y = setup_case(18, 128, 256)
        rep, eng, _, _ = compare_step(f"ablation_{precision}", 18, 128, model, params, buffers, X, Y, precision=precision)
        out[precision] = {k: rep[k] for k in ("e_new_raw", "e_new_reg", "e_new_avg", "cos_raw", "cos_reg", "cos_avg",
                                              "e_new_loss", "e_new_grad_norms")}
        out["fp32_oracle"] = {k: rep[k] for k in ("e32_raw", "e32_reg", "e32_avg", "e32_loss", "e32_grad_norms")}
        del eng
    dump("ablation_r18_mb128_n256", out)
    s, b = out["split"], out["bf16"]
    assert s["e_new_raw"] * 20 < b["e_new_raw"] and s["e_new_reg"] * 5 < b["e_new_reg"]
    # plain bf16 would not pass assert_step
    assert b["e_new_raw"] > RATIO_RAW * max(out["fp32_oracle"]["e32_raw"], FLOOR_RAW)


def test_resnet152_real_microbatch_in_a_conditioned_state():
    """ResNet-152 at microbatch 32 (config 4).  At the reference's initialisation the 50 residual branches add up
    unattenuated (zero_init_residual is off on the config path, models.py:22): squared gradient norms are ~3e6 and the
    fp32 reference itself is 12 % / 91 % away from fp64 (raw / regularised), so nothing can be asserted there beyond the
    ratios (test_full_batch_step_matches_oracle[r152_mb32]).  With the last BatchNorm scale of every block at 0.25 -- the
    regime a trained network is in -- the problem is conditioned, the fp32 floor drops below 0.2 and the cosine
    assertions are live."""
    depth, mb, n = 152, 32, 64
    model, _, _, X, Y = setup_case(depth, mb, n)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("bn3.weight"):
                p.fill_(0.25)
    params = {k: v.detach().clone() for k, v in model.named_parameters()}
    buffers = {k: v.detach().clone() for k, v in model.named_buffers()}
    rep, _, _, _ = compare_step("r152_mb32_n64_damped", depth, mb, model, params, buffers, X, Y, groups=2)
    assert_step(rep, depth)
    assert rep["e32_reg"] < 0.2 and rep["e32_avg"] < 0.2, "the state is meant to be conditioned"
    assert rep["cos_raw"] > 0.998 and rep["cos_reg"] > 0.97 and rep["cos_avg"] > 0.97  # measured 0.9989 / 0.9765 / 0.9765


def golden(name):
    import numpy as np

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def check_fingerprint(z, prefix, tensors, stride, rtol):
    import numpy as np

    fp = O.fingerprint(tensors, stride)
    assert abs(fp["total_norm"] - z[f"{prefix}.total_norm"]) <= rtol * z[f"{prefix}.total_norm"], prefix
    assert np.linalg.norm(fp["norms"] - z[f"{prefix}.norms"]) <= rtol * np.linalg.norm(z[f"{prefix}.norms"]), prefix
    assert np.linalg.norm(fp["sample"] - z[f"{prefix}.sample"]) <= rtol * np.linalg.norm(z[f"{prefix}.sample"]), prefix


@pytest.mark.parametrize("name,groups", [("r18_mb128_n2000_f64", None), ("r152_mb32_n64_f64", 2)])
def test_baseline_config_goldens_of_the_unmodified_reference(name, groups):
    """BASELINE.json configs[0] (ResNet-18, 2,000 images -> K = 15 microbatches of 128: one launch of 8 and one of 7
    groups) and config 4's shape (ResNet-152, microbatch 32): fixtures produced by the UNMODIFIED reference in fp64
    (oracle/make_goldens.py).  The oracle (fp64, on this GPU) must reproduce them, and the engine is compared with
    both: accumulated gradient, loss and the 15 per-microbatch gradient norms of `stats`."""
    z, meta = golden(name)
    depth, mb, n = meta["depth"], meta["mb"], meta["n"]
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    rep, eng, ref64, buf64 = compare_step(name, depth, mb, model, params, buffers, X, Y, groups=groups,
                                          per_microbatch=False)
    # 1) the oracle reproduces the reference (fp64 on the GPU vs fp64 on the CPU that made the fixture)
    check_fingerprint(z, "avg", ref64["avg"], meta["stride"], 1e-6)
    sc = meta["scalars"]
    assert float(ref64["loss"]) == pytest.approx(sc["train_loss"], rel=1e-9)
    assert ref64["grad_norms"].sqrt().tolist() == pytest.approx(sc["grad_norm_train"], rel=1e-7)
    # 2) the engine against the reference's own numbers
    assert rep["K"] == n // mb == len(sc["grad_norm_train"])
    assert rep["loss"] == pytest.approx(sc["train_loss"], rel=1e-4 if depth < 100 else 1e-3)
    assert [v ** 0.5 for v in rep["grad_norms"]] == pytest.approx(sc["grad_norm_train"], rel=1e-3 if depth < 100 else 2e-2)
    assert rep["correct"] / (rep["K"] * mb) == pytest.approx(sc["train_acc"])
    assert_step(rep, depth)
    # fingerprint of the engine's accumulated gradient against the fixture: per-tensor norms within the same bound
    import numpy as np

    fp = O.fingerprint(eng.grads_list(eng.avg), meta["stride"])
    tol = RATIO_REG * max(rep["e32_avg"], FLOOR_REG)
    assert abs(fp["total_norm"] - z["avg.total_norm"]) <= tol * z["avg.total_norm"]
    assert np.linalg.norm(fp["norms"] - z["avg.norms"]) <= tol * np.linalg.norm(z["avg.norms"])  # per-tensor norms
    # a strided sample of 2,236 of the 11.2 M elements: dominated by the small gradients of the wide late layers, where
    # the (absolute) finite-difference noise weighs more than in the whole-vector norm -> twice the bound
    assert np.linalg.norm(fp["sample"] - z["avg.sample"]) <= 2 * tol * np.linalg.norm(z["avg.sample"])


def test_parity_away_from_initialisation():
    """SURVEY.md 7 hard part 2: after three gradient-descent steps of the oracle (fp64: weights, BatchNorm parameters and
    running statistics have moved) the fourth step's gradients are compared again."""
    depth, mb, n = 18, 128, 256
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    p = {k: v.to(DEV, torch.float64).clone() for k, v in params.items()}
    b = {k: (v.to(DEV).clone() if v.dtype == torch.long else v.to(DEV, torch.float64).clone()) for k, v in buffers.items()}
    for _ in range(3):
        out = O.full_batch_step(depth, p, b, X.double(), Y, mb, **HYP)
        for v, g in zip(p.values(), out["avg"]):
            v.sub_(0.01 * g)
    moved = rel(O.flat(list(p.values())), O.flat([v.to(DEV) for v in params.values()]))
    assert moved > 1e-3
    with torch.no_grad():
        model.load_state_dict({**{k: v.float().cpu() for k, v in p.items()},
                               **{k: (v.cpu() if v.dtype == torch.long else v.float().cpu()) for k, v in b.items()}})
    params = {k: v.detach().clone() for k, v in model.named_parameters()}  # fp32-rounded: every path starts from these
    buffers = {k: v.detach().clone() for k, v in model.named_buffers()}
    rep, _, _, _ = compare_step("r18_mb128_after3steps", depth, mb, model, params, buffers, X, Y)
    assert_step(rep, depth)


def test_parity_on_structured_images():
    """spatially correlated, class-structured inputs (oracle.structured_cifar) instead of white noise"""
    depth, mb, n = 18, 128, 256
    model, params, buffers, X, Y = setup_case(depth, mb, n, data="structured")
    rep, _, _, _ = compare_step("r18_mb128_structured", depth, mb, model, params, buffers, X, Y)
    assert_step(rep, depth)


def test_config2_50k_step_against_committed_oracle_scalars():
    """BASELINE.json configs[1] at full size (ResNet-18, 49,920 images = 390 microbatches = 48 launches of 8 groups + one
    of 6), on bench.py's data: mean loss and mean squared gradient norm against the fp32 / fp64 ORACLE values committed in
    tests/golden/bench_check.json (tools/make_bench_check.py), i.e. the running sums survive 390 microbatches."""
    from fullbatchtraining_b200.data import synthetic_cifar

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_check.json")) as f:
        exp = json.load(f)["r18_50k"]
    mb, K = 128, exp["microbatches"]
    torch.manual_seed(0)
    model = construct_model(dict(name="ResNet18", depth=18), 3, 10)
    X, Y = synthetic_cifar(K * mb, device=DEV)
    eng = FullBatchEngine(model, mb, precision="split")
    assert eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"]) == K
    res = eng.results(K)
    assert res["loss"] == pytest.approx(exp["loss_fp64"], rel=2e-5)
    assert abs(res["loss"] - exp["loss_fp64"]) <= 6.0 * max(abs(exp["loss"] - exp["loss_fp64"]), 1e-6)
    assert float(res["grad_norms"].mean()) == pytest.approx(exp["mean_grad_norm_sq_fp64"], rel=1e-3)
    assert torch.isfinite(eng.avg).all() and 0.5 < float(eng.avg.norm()) < 50.0
    dump("r18_50k_scalars", dict(loss=res["loss"], loss_fp32_oracle=exp["loss"], loss_fp64_oracle=exp["loss_fp64"],
                                 mean_grad_norm_sq=float(res["grad_norms"].mean()),
                                 mean_grad_norm_sq_fp64_oracle=exp["mean_grad_norm_sq_fp64"], avg_norm=float(eng.avg.norm())))


def test_running_stats_and_determinism():
    depth, mb, n = 18, 16, 32
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    _, buf64 = oracle_run(depth, params, buffers, X, Y, mb, torch.float64, keep=0)
    eng = FullBatchEngine(model, mb, precision="split")
    K = eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"])
    eng.sync_bn_counters()
    first = eng.avg.clone()
    for name, b in model.named_buffers():
        if name.endswith("num_batches_tracked"):
            assert int(b) == 2 * K == int(buf64[name])
        else:
            assert rel(b, buf64[name]) < 1e-3, name
    # measure_floating_point_accuracy-style drift (training.py:573-598): rerun from the same state -> bit identical
    for name, b in model.named_buffers():
        b.copy_(buffers[name].to(DEV))
    eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"])
    assert torch.equal(first, eng.avg)


def test_result_does_not_depend_on_the_group_count():
    """G microbatches per launch is a pure performance knob: 1, 2 (with a remainder launch) and 5 groups give the same
    bits for the accumulated gradient, the gradient norms, the loss and the BatchNorm running statistics."""
    depth, mb, n = 18, 16, 80
    outs = []
    for G in (1, 2, 5):
        model, params, buffers, X, Y = setup_case(depth, mb, n)
        eng = FullBatchEngine(model, mb, precision="split", groups=G)
        K = eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2)
        res = eng.results(K)
        outs.append((eng.avg.clone(), res["grad_norms"].clone(), torch.tensor(res["loss_sum"]),
                     torch.cat([b.reshape(-1).float() for b in model.buffers()])))
        assert K == 5
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)


def test_result_does_not_depend_on_the_lane_count():
    """Two lanes (alternating group launches on two streams, commits in loader order) == one lane, bit for bit; also for
    the acc_strength pre-pass and through the host-streamed path."""
    from fullbatchtraining_b200.data import HostBlockLoader

    depth, mb, n = 18, 16, 112
    for extra in (dict(), dict(acc_strength=0.3, batch_clip=5.0)):
        outs = []
        for lanes in (1, 2):
            model, params, buffers, X, Y = setup_case(depth, mb, n)
            eng = FullBatchEngine(model, mb, precision="split", groups=2, lanes=lanes)
            assert len(eng.lanes) == lanes
            K = eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, **extra)
            res = eng.results(K)
            outs.append((eng.avg.clone(), res["grad_norms"].clone(), torch.tensor([res["loss_sum"], res["correct"],
                                                                                   float(res["clipped_batches"])]),
                         torch.cat([b.reshape(-1).float() for b in model.buffers()])))
            assert K == 7
        for a, b in zip(*outs):  # gradient, gradient norms, loss / accuracy / clip counters, running statistics
            assert torch.equal(a, b)
    # streamed from the host: three staging buffers, two lanes
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    eng = FullBatchEngine(model, mb, precision="split", groups=2, lanes=2)
    K = eng.accumulate_stream(HostBlockLoader(X.cpu(), Y.cpu(), mb), 0.8, 0.5, 1e-2, n // mb)
    model1, _, _, _, _ = setup_case(depth, mb, n)
    eng1 = FullBatchEngine(model1, mb, precision="split", groups=2, lanes=1)
    eng1.accumulate_resident(X, Y, 0.8, 0.5, 1e-2)
    assert K == 7 and torch.equal(eng.avg, eng1.avg)
    assert eng.results(K)["loss_sum"] == eng1.results(K)["loss_sum"]


def test_graph_replay_equals_eager():
    depth, mb, n = 18, 16, 48
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    eng = FullBatchEngine(model, mb, precision="split")
    eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, use_graph=False)
    eager = eng.avg.clone()
    norms = eng.grad_norms[:3].clone()
    for name, b in model.named_buffers():
        b.copy_(buffers[name].to(DEV))
    eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, use_graph=True)
    assert torch.equal(eager, eng.avg)
    assert torch.equal(norms, eng.grad_norms[:3])


def test_no_regulariser_pass_is_plain_mean():
    depth, mb, n = 18, 16, 32
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    eng = FullBatchEngine(model, mb, precision="split")
    K = eng.accumulate_resident(X, Y, 0.8, 0.0, 1e-2)
    p = {k: v.to(DEV, torch.float64) for k, v in params.items()}
    b = {k: (v.to(DEV) if v.dtype == torch.long else v.to(DEV, torch.float64)) for k, v in buffers.items()}
    ref = O.full_batch_step(depth, p, b, X.double(), Y, mb, lr=0.8, block_strength=0.0)
    assert rel(eng.avg, O.flat(ref["avg"])) < 2e-2  # mb=16: fp32 floor ~2.5e-3, split ~1e-2
    assert K == 2


def test_shuffled_order_through_permutation():
    """hyp.shuffle (data_preparation.py:53-54): microbatches are gathered through an index tensor on the device."""
    depth, mb, n = 18, 16, 48
    model, params, buffers, X, Y = setup_case(depth, mb, n)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(3)).to(DEV)
    eng = FullBatchEngine(model, mb, precision="split")
    K = eng.accumulate_resident(X, Y, 0.8, 0.5, 1e-2, perm=perm)
    p = {k: v.to(DEV, torch.float64) for k, v in params.items()}
    b = {k: (v.to(DEV) if v.dtype == torch.long else v.to(DEV, torch.float64)) for k, v in buffers.items()}
    ref = O.full_batch_step(depth, p, b, X.double(), Y, mb, order=perm, **HYP)
    assert K == 3
    assert abs(eng.results(K)["loss"] - float(ref["loss"])) < 1e-4 * float(ref["loss"])
    assert rel(eng.avg, O.flat(ref["avg"])) < 0.2  # ABS_REG_R18


VARIANTS = [
    ("central", dict(implementation="central-differences")),
    ("legacy", dict(implementation="forward-differences-legacy")),
    ("acc", dict(acc_strength=0.3)),
    ("central_acc", dict(implementation="central-differences", acc_strength=0.2)),
    ("batch_clip", dict(batch_clip=5.0)),
    ("acc_batch_clip", dict(acc_strength=0.3, batch_clip=5.0)),
]


@pytest.mark.parametrize("name,extra", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_grad_reg_variants_match_oracle(name, extra):
    """The rest of the hyp.grad_reg / hyp.batch_clip surface (modules.py:243-300, training.py:128-142,166-168) on the
    same kernels, against the oracle in fp64 with the fp32 oracle as noise floor."""
    depth, mb, n = 18, 16, 48
    model, params, buffers, X, Y = setup_case(depth, mb, n)

    def oracle(dtype):
        p = {k: v.to(DEV, dtype).clone() for k, v in params.items()}
        b = {k: (v.to(DEV).clone() if v.dtype == torch.long else v.to(DEV, dtype).clone()) for k, v in buffers.items()}
        return O.full_batch_step(depth, p, b, X.to(dtype), Y, mb, **HYP, **extra), b

    ref64, buf64 = oracle(torch.float64)
    ref32, _ = oracle(torch.float32)
    eng = FullBatchEngine(model, mb, precision="split")
    K = eng.accumulate_resident(X, Y, HYP["lr"], HYP["block_strength"], HYP["eps"], **extra)
    eng.sync_bn_counters()
    res = eng.results(K)
    avg64 = O.flat(ref64["avg"])
    e_new, e32 = rel(eng.avg, avg64), rel(O.flat(ref32["avg"]), avg64)
    dump(f"variant_{name}", dict(e_new_avg=e_new, e32_avg=e32, cos=cos(eng.avg, avg64),
                                 clipped=res["clipped_batches"], clipped64=ref64["clipped_batches"]))
    assert e_new <= RATIO_REG_MB * max(e32, FLOOR_REG)  # 3 microbatches of 16: the fp32 reference error is noisy
    assert e_new <= ABS_REG_R18 and cos(eng.avg, avg64) > 0.98
    assert abs(res["loss"] - float(ref64["loss"])) < 1e-4 * float(ref64["loss"])
    assert res["clipped_batches"] == ref64["clipped_batches"]
    if "acc_strength" in extra:
        assert rel(eng.pre, O.flat(ref64["pre_grads"])) < ABS_RAW_R18  # mean RAW gradient: 1.1-1.4e-2 at microbatch 16
    nbt = [b for k, b in model.named_buffers() if k.endswith("num_batches_tracked")][0]
    assert int(nbt) == int(buf64["stem.1.num_batches_tracked"])  # 2 or 3 passes per microbatch (+1 for the pre-pass)


def test_stock_pytorch_gpu_baseline_is_recorded():
    """SURVEY.md 8d "the real kernel to beat": the torch-op restatement of the path (cuDNN / cuBLAS convolutions and
    autograd, exactly what the reference executes on a GPU) timed on the same B200 in fp32 and with TF32 convolutions,
    next to the engine on the same microbatches.  Written to gpurun_out/stock_pytorch.json; the assertion only guards
    against the engine being slower than the library path it replaces."""
    depth, mb, n = 18, 128, 128 * 6
    model, params, buffers, X, Y = setup_case(depth, mb, n)

    def stock(tf32):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        p = {k: v.to(DEV).clone() for k, v in params.items()}
        b = {k: v.to(DEV).clone() for k, v in buffers.items()}
        O.full_batch_step(depth, p, b, X[:2 * mb], Y[:2 * mb], mb, **HYP)  # warm-up, cuDNN autotune
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        O.full_batch_step(depth, p, b, X, Y, mb, **HYP)
        e1.record()
        torch.cuda.synchronize()
        return n / (e0.elapsed_time(e1) * 1e-3)

    try:
        fp32, tf32 = stock(False), stock(True)
    finally:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = False
    eng = FullBatchEngine(model, mb, precision="split", device=DEV)
    Xf = X.float().contiguous()
    for _ in range(2):
        eng.begin_step(n // mb)
        eng.accumulate_resident(Xf, Y, HYP["lr"], HYP["block_strength"], HYP["eps"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.begin_step(n // mb)
    eng.accumulate_resident(Xf, Y, HYP["lr"], HYP["block_strength"], HYP["eps"])
    e1.record()
    torch.cuda.synchronize()
    ours = n / (e0.elapsed_time(e1) * 1e-3)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "stock_pytorch.json"), "w") as f:
        json.dump(dict(workload=f"ResNet-{depth} mb {mb}, {n} images, one full-batch grad-reg step", unit="images/s",
                       stock_pytorch_fp32=fp32, stock_pytorch_tf32=tf32, engine_split=ours,
                       torch=torch.__version__), f, indent=1)
    assert ours > fp32, (ours, fp32)
