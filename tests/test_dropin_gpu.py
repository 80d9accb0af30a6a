"""Drop-in surface on the B200: GradRegularizer call protocol (reference fullbatch/models/modules.py:346-348,211-241),
Trainer/train (fullbatch/training/training.py:50-340) with resident and host-streamed data, SGD sanity branch."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fullbatchtraining_b200 import construct_model  # noqa: E402
from fullbatchtraining_b200.config import default_cfg  # noqa: E402
from fullbatchtraining_b200.data import HostBlockLoader  # noqa: E402
from fullbatchtraining_b200.modules import GradRegularizer, LabelSmoothCrossEntropyLoss  # noqa: E402
from fullbatchtraining_b200.training import Trainer, train  # noqa: E402
from oracle import fb_oracle as O  # noqa: E402

DEV = torch.device("cuda")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def fresh(depth=18):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    return construct_model(dict(name=f"ResNet{depth}", depth=depth), 3, 10)


def test_gradregularizer_call_protocol():
    model = fresh().to(DEV)
    mb = 16
    X, Y = O.synthetic_cifar(mb)
    X, Y = X.to(DEV), Y.to(DEV)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.8)
    loss_fn = LabelSmoothCrossEntropyLoss(0.0)
    # raw gradient by plain autograd, as in the reference README (README.md:36-44)
    model.train()
    grads = torch.autograd.grad(loss_fn(model(X), Y), list(model.parameters()))
    grads = [g.clone() for g in grads]
    raw = [g.clone() for g in grads]
    theta_before = [p.detach().clone() for p in model.parameters()]
    gradreg = GradRegularizer(model, optimizer, loss_fn, block_strength=0.5, eps=1e-2, implementation="finite_diff",
                              microbatch=mb)
    assert gradreg.create_graph is False
    out = gradreg(grads, X, Y, None)
    assert out is grads  # in place, same list (modules.py:240-241)
    for p, q in zip(model.parameters(), theta_before):
        assert torch.equal(p.detach(), q)  # parameters restored exactly (modules.py:237-238)
    # oracle: forward differences in fp64 starting from the same raw gradient
    p64 = {k: v.detach().double() for k, v in model.named_parameters()}
    b64 = {k: (v.detach().clone() if v.dtype == torch.long else v.detach().double()) for k, v in model.named_buffers()}
    net = O.OracleResNet(18, b64, update_running_stats=False)
    ref, _, _ = O.forward_differences(net, p64, [g.double() for g in raw], X.double(), Y, 0.8, 0.5, 1e-2)
    err = rel(O.flat(grads), O.flat(ref))
    assert err < 0.2, err  # FD term divides operand rounding by eps_n; fp32 itself is ~5e-2 here, this path 0.11
    cos = float((O.flat(grads).double() * O.flat(ref)).sum() / (O.flat(grads).double().norm() * O.flat(ref).norm()))
    assert cos > 0.98


def test_gradregularizer_rejects_unknown_implementation():
    model = fresh()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    with pytest.raises(ValueError):
        GradRegularizer(model, opt, None, block_strength=0.5, implementation="nonsense")
    with pytest.raises(ValueError):
        GradRegularizer(model, opt, None, block_strength=0.5, implementation="autograd")  # double backward: not on the path
    g = GradRegularizer(model, opt, None, block_strength=0.0, acc_strength=0.0, implementation="nonsense")
    grads = [torch.zeros(1)]
    assert g(grads, None, None, None) is grads  # _pass, modules.py:151-153


def _cfg(mb, **extra):
    o = {"data.batch_size": mb, "hyp.sub_batch": mb, "hyp.warmup": 0, "hyp.steps": 1, "hyp.grad_clip": None}
    o.update(extra)
    return default_cfg(o)


def test_train_one_step_matches_oracle_and_stream_equals_resident():
    mb, n = 16, 48
    X, Y = O.synthetic_cifar(n)
    setup = dict(device=DEV, dtype=torch.float32)
    # resident: a TensorDataset loader like the oracle harness builds (data_preparation.py:56-72)
    ds = torch.utils.data.TensorDataset(X, Y)
    sampler = torch.utils.data.SequentialSampler(ds)
    loader = torch.utils.data.DataLoader(ds, batch_size=mb, sampler=sampler, drop_last=True)
    model = fresh()
    p64 = {k: v.detach().to(DEV, torch.float64) for k, v in model.named_parameters()}
    b64 = {k: (v.detach().to(DEV) if v.dtype == torch.long else v.detach().to(DEV, torch.float64))
           for k, v in model.named_buffers()}
    stats = train(model, loader, None, setup, _cfg(mb))
    ref = O.full_batch_step(18, p64, b64, X.to(DEV).double(), Y.to(DEV), mb, lr=0.8, block_strength=0.5, eps=1e-2)
    assert stats["train_loss"][0] == pytest.approx(float(ref["loss"]), rel=1e-4)
    assert stats["train_acc"][0] == pytest.approx(float(ref["correct"]) / n)
    assert stats["grad_norm"][0] == pytest.approx(math.sqrt(float(ref["grad_norms"].mean())), rel=1e-2)
    full = float(ref["loss"]) + 0.5 * 5e-4 * float(ref["param_norm"]) + 0.8 / 4 * 0.5 * float(ref["grad_norms"].mean())
    assert stats["full_loss"][0] == pytest.approx(full, rel=1e-2)
    for k in range(3):
        assert f"grad_norm_train_{k}" in stats
    # the SGD update used the accumulated gradient: theta1 = theta0 - lr * (g + wd*theta0)  (nesterov, first step:
    # buf = d, d + momentum*buf = 1.9 d)
    g = O.flat(ref["avg"])
    th0 = O.flat(list(p64.values()))
    th1 = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double()
    d = g + 5e-4 * th0
    expect = th0 - 0.8 * 1.9 * d
    assert rel(th1 - th0, expect - th0) < 0.2  # = the error of the accumulated regularised gradient (0.12)
    # host-streamed blocks give bit-identical accumulation
    model_a, model_b = fresh(), fresh()
    ta = Trainer(model_a, loader, None, setup, _cfg(mb))
    tb = Trainer(model_b, HostBlockLoader(X, Y, mb), None, setup, _cfg(mb, **{"impl.resident_dataset": False}))
    assert ta.resident is not None and tb.resident is None
    ta._accumulate_full_gradient()
    tb._accumulate_full_gradient()
    assert torch.equal(ta.engine.avg, tb.engine.avg)
    assert tb.engine.h2d_bytes == n * (3 * 32 * 32 * 4 + 8)


def test_grad_clip_and_lr_schedule():
    mb, n = 16, 32
    X, Y = O.synthetic_cifar(n)
    loader = HostBlockLoader(X, Y, mb)
    cfg = _cfg(mb, **{"hyp.grad_clip": 0.25, "hyp.warmup": 2, "hyp.steps": 3})
    model = fresh()
    trainer = Trainer(model, loader, None, dict(device=DEV, dtype=torch.float32), cfg)
    lrs = []
    for _ in range(3):
        lrs.append(trainer.optimizer.param_groups[0]["lr"])
        trainer.step(validate=False)
    # scheduler.py:57-66: lr = base * step / warmup during warm-up, base afterwards (cosine-4000 starts one step later)
    assert lrs == pytest.approx([0.0, 0.4, 0.8])
    assert trainer.stats["clipped_step"] == [1, 1, 1]
    assert float(trainer.engine.avg.norm()) == pytest.approx(0.25, rel=1e-3)
    assert all(math.isfinite(v) for v in trainer.stats["train_loss"])


def test_sgd_sanity_branch_matches_oracle_and_reference_golden():
    """BASELINE.json configs[4], training.py:241-286: two steps of the stochastic branch (4 blocks of 16, the
    regulariser per block, Nesterov SGD) through `train`; parameters afterwards against the oracle in fp64 (pinned to a
    fixture of the unmodified reference, tests/test_oracle_golden.py) within 1e-3, per-step losses within 1e-3."""
    import json
    import os

    import numpy as np

    mb, n, steps, lr = 16, 64, 2, 0.001
    X, Y = O.synthetic_cifar(n)
    loader = HostBlockLoader(X, Y, mb)
    cfg = _cfg(mb, **{"hyp.train_stochastic": True, "hyp.optim.lr": lr, "hyp.steps": steps, "hyp.grad_clip": None})
    model = fresh()
    p64 = {k: v.detach().to(DEV, torch.float64).clone() for k, v in model.named_parameters()}
    b64 = {k: (v.detach().to(DEV).clone() if v.dtype == torch.long else v.detach().to(DEV, torch.float64).clone())
           for k, v in model.named_buffers()}
    th0 = O.flat(list(p64.values())).clone()
    stats = train(model, loader, None, dict(device=DEV, dtype=torch.float32), cfg)
    ref = O.sgd_epochs(18, p64, b64, X.to(DEV).double(), Y.to(DEV), mb, steps, lr, block_strength=0.5, eps=1e-2)
    # the same in fp32 (the reference's own arithmetic): its distance from fp64 is the noise floor of this trajectory
    m32 = fresh()
    p32 = {k: v.detach().to(DEV).clone() for k, v in m32.named_parameters()}
    b32 = {k: v.detach().to(DEV).clone() for k, v in m32.named_buffers()}
    ref32 = O.sgd_epochs(18, p32, b32, X.to(DEV), Y.to(DEV), mb, steps, lr, block_strength=0.5, eps=1e-2)
    th_ref = O.flat(list(p64.values()))
    th_32 = O.flat(list(p32.values())).double()
    th_new = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double()
    moved = float((th_ref - th0).norm())
    e_new, e32 = float((th_new - th_ref).norm()) / moved, float((th_32 - th_ref).norm()) / moved
    l_new = max(abs(a - b) / abs(b) for a, b in zip(stats["train_loss"], ref["train_loss"]))
    l32 = max(abs(a - b) / abs(b) for a, b in zip(ref32["train_loss"], ref["train_loss"]))
    print(f"sgd branch: update error {e_new:.3e} (fp32 reference {e32:.3e}), loss error {l_new:.3e} (fp32 {l32:.3e}), "
          f"theta rel {rel(th_new, th_ref):.3e}")
    assert len(stats["train_loss"]) == steps
    assert stats["train_acc"] == pytest.approx(ref["train_acc"], abs=1.0 / n)
    assert rel(th_new, th_ref) < 1e-3                      # parameters after 8 optimizer steps
    assert e_new <= 6.0 * max(e32, 2e-3) and e_new < 0.05  # the accumulated UPDATE: 6x the fp32 reference's own error
    assert l_new <= 6.0 * max(l32, 2e-3)
    assert [v ** 2 for v in stats["grad_norm"]] == pytest.approx(ref["grad_norm_sq"], rel=2e-2)
    for name, buf in model.named_buffers():
        if not name.endswith("num_batches_tracked"):
            assert rel(buf, b64[name]) < 5e-3, name
    # and against the fixture of the unmodified reference
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r18_sgd_mb16_n64_f64.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    assert stats["train_loss"] == pytest.approx(meta["scalars"]["train_loss_steps"], rel=6.0 * max(l32, 2e-3))
    fp = O.fingerprint([q.detach() for q in model.parameters()], meta["stride"])
    assert np.linalg.norm(fp["sample"] - z["theta.sample"]) <= 1e-3 * np.linalg.norm(z["theta.sample"])


def test_unsupported_reference_options_are_refused():
    """options of the reference that are not on the B200 path raise instead of being silently ignored"""
    X, Y = O.synthetic_cifar(32)
    setup = dict(device=DEV, dtype=torch.float32)
    for key, val in [("hyp.evaluate_ema", True), ("hyp.grad_noise.additive", 0.1), ("hyp.norm_bias.strength", 0.1),
                     ("hyp.only_linear_layers_weight_decay", True), ("hyp.train_semi_stochastic", True),
                     ("impl.mixed_precision", True)]:
        with pytest.raises(ValueError):
            Trainer(fresh(), HostBlockLoader(X, Y, 16), None, setup, _cfg(16, **{key: val}))
    # a resident dataset whose last block is ragged (drop_last=False) cannot be walked out of bounds
    ds = torch.utils.data.TensorDataset(*O.synthetic_cifar(40))
    loader = torch.utils.data.DataLoader(ds, batch_size=16, sampler=torch.utils.data.SequentialSampler(ds), drop_last=False)
    with pytest.raises(ValueError):
        Trainer(fresh(), loader, None, setup, _cfg(16))
    from fullbatchtraining_b200.engine import FullBatchEngine

    eng = FullBatchEngine(fresh(), 16, groups=2)
    Xd, Yd = (t.to(DEV) for t in O.synthetic_cifar(40))
    with pytest.raises(ValueError):
        eng.accumulate_resident(Xd, Yd, 0.8, 0.5, 1e-2, count=3)


def test_resumed_step_records_the_same_statistics(tmp_path):
    """param_norm / full_loss of the first step after a resume equal those of the uninterrupted run (sum theta^2 is
    re-initialised on the device when the checkpoint is loaded)."""
    mb, n = 16, 32
    X, Y = O.synthetic_cifar(n)
    setup = dict(device=DEV, dtype=torch.float32)

    def run(steps, name):
        cfg = _cfg(mb, **{"hyp.steps": steps, "hyp.grad_clip": 0.25, "impl.checkpoint.name": name,
                          "original_cwd": str(tmp_path), "hyp.warmup": 0})
        trainer = Trainer(fresh(), HostBlockLoader(X, Y, mb), None, setup, cfg)
        while trainer.step_count < steps:
            trainer.step(validate=False)
            trainer.maybe_checkpoint()
        return trainer.stats

    full = run(3, None)
    run(2, "r.pth")
    resumed = run(3, "r.pth")
    assert resumed["param_norm"][0] == pytest.approx(full["param_norm"][2], rel=1e-6) and resumed["param_norm"][0] > 0
    assert resumed["full_loss"][0] == pytest.approx(full["full_loss"][2], rel=1e-6)
    assert full["param_norm"][0] == pytest.approx(float(sum(p.double().pow(2).sum() for p in fresh().parameters())), rel=1e-6)


def test_flat_sgd_matches_torch_sgd_with_clip():
    """FlatSGD (fb_sgd_step: clip + weight decay + Nesterov momentum in one sweep) against torch.optim.SGD +
    the reference's clip (training.py:198-211) on the same gradients, three steps."""
    from fullbatchtraining_b200.engine import FullBatchEngine
    from fullbatchtraining_b200.optim import FlatSGD, S_GNORM, S_PNORM

    model = fresh()
    eng = FullBatchEngine(model, 16)
    ref_params = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    kw = dict(lr=0.8, momentum=0.9, dampening=0.0, weight_decay=5e-4, nesterov=True)
    opt = FlatSGD(model.parameters(), **kw).bind(eng, grad_clip=0.25)
    ref = torch.optim.SGD(ref_params, **kw)
    g = torch.Generator(device="cuda").manual_seed(0)
    for step in range(3):
        grad = torch.randn(eng.numel, device=DEV, generator=g) * (1e-3 if step == 1 else 1e-2)
        eng.avg.copy_(grad)
        norm = grad.norm()
        clipped = grad * (0.25 / (norm + 1e-6)) if norm > 0.25 else grad
        off = 0
        for p in ref_params:
            p.grad = clipped[off:off + p.numel()].view_as(p).clone()
            off += p.numel()
        ref.step()
        opt.step()
        flat_ref = torch.cat([p.detach().reshape(-1) for p in ref_params])
        assert rel(eng.theta, flat_ref) < 1e-6
        assert float(eng.scal[S_GNORM]) == pytest.approx(float(norm) ** 2, rel=1e-5)
        assert float(eng.scal[S_PNORM]) == pytest.approx(float(flat_ref.double().pow(2).sum()), rel=1e-5)
        assert rel(eng.avg, clipped) < 1e-6  # param.grad clipped in place like the reference
    sd = opt.state_dict()
    assert "momentum_buffer" in sd["state"][0]


def test_gradregularizer_central_differences_and_pre_grads():
    """GradRegularizer(..., implementation='central-differences', acc_strength) with pre_grads (modules.py:266-300)."""
    model = fresh().to(DEV)
    mb = 16
    X, Y = O.synthetic_cifar(mb)
    X, Y = X.to(DEV), Y.to(DEV)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.8)
    loss_fn = LabelSmoothCrossEntropyLoss(0.0)
    model.train()
    raw = [g.clone() for g in torch.autograd.grad(loss_fn(model(X), Y), list(model.parameters()))]
    gen = torch.Generator(device="cuda").manual_seed(1)
    pre = [torch.randn(g.shape, device=DEV, generator=gen) * g.abs().mean() for g in raw]
    gradreg = GradRegularizer(model, optimizer, loss_fn, block_strength=0.5, acc_strength=0.25, eps=1e-2,
                              implementation="central-differences", microbatch=mb)
    grads = [g.clone() for g in raw]
    out = gradreg(grads, X, Y, pre)
    assert out is grads
    p64 = {k: v.detach().double() for k, v in model.named_parameters()}
    b64 = {k: (v.detach().clone() if v.dtype == torch.long else v.detach().double()) for k, v in model.named_buffers()}
    net = O.OracleResNet(18, b64, update_running_stats=False)
    ref, _, _ = O.regularize(net, p64, [g.double() for g in raw], X.double(), Y, 0.8, 0.5, 1e-2, 0.0,
                             "central-differences", [q.double() for q in pre], 0.25)
    assert rel(O.flat(grads), O.flat(ref)) < 0.2
    c = float((O.flat(grads).double() * O.flat(ref)).sum() / (O.flat(grads).double().norm() * O.flat(ref).norm()))
    assert c > 0.98


def test_device_side_augmentation_and_shuffle_pipeline():
    """uint8 HWC dataset kept in HBM: normalisation (+ crop / flip / per-step shuffle) happen inside the stem kernel.
    Without augmentation and shuffling the step equals the float-dataset path; with them it runs and changes per step."""
    mb, n = 16, 64
    g = torch.Generator().manual_seed(4)
    raw = torch.randint(0, 256, (n, 32, 32, 3), generator=g, dtype=torch.uint8)
    Y = torch.randint(0, 10, (n,), generator=g)
    cfg0 = _cfg(mb, **{"data.augmentations_train": None})
    mean = torch.tensor(cfg0.data.mean)[None, :, None, None]
    std = torch.tensor(cfg0.data.std)[None, :, None, None]
    Xf = (raw.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    setup = dict(device=DEV, dtype=torch.float32)

    def loader(x):
        ds = torch.utils.data.TensorDataset(x, Y)
        return torch.utils.data.DataLoader(ds, batch_size=mb, sampler=torch.utils.data.SequentialSampler(ds), drop_last=True)

    ta = Trainer(fresh(), loader(raw), None, setup, cfg0)
    tb = Trainer(fresh(), loader(Xf), None, setup, _cfg(mb))
    assert ta.resident[0].dtype == torch.uint8 and not ta.augment
    la, lb = ta._accumulate_full_gradient(), tb._accumulate_full_gradient()
    assert float(la) == pytest.approx(float(lb), rel=1e-5)
    assert rel(ta.engine.avg, tb.engine.avg) < 0.2  # same data up to 1 ulp of the normalisation; FD amplifies it
    # augmentation + shuffle
    cfg1 = _cfg(mb, **{"hyp.shuffle": True, "seed": 3})
    tc = Trainer(fresh(), loader(raw), None, setup, cfg1)
    assert tc.augment and tc.crop_pad == 4 and tc.flip_p == 0.5
    l1 = float(tc._accumulate_full_gradient())
    perm1, aug1 = tc.perm.clone(), tc.engine.aug_params.clone()
    l2 = float(tc._accumulate_full_gradient())
    assert math.isfinite(l1) and math.isfinite(l2)
    assert not torch.equal(perm1, tc.perm) and not torch.equal(aug1, tc.engine.aug_params)
    assert sorted(tc.perm.tolist()) == list(range(n))
    assert int(tc.engine.aug_params[:, :2].max()) <= 8 and int(tc.engine.aug_params[:, :2].min()) >= 0


def test_checkpoint_roundtrip_in_reference_format(tmp_path):
    """training/utils.py:43-70: 5-list checkpoint; resuming reproduces the uninterrupted run bit for bit."""
    mb, n = 16, 32
    X, Y = O.synthetic_cifar(n)
    setup = dict(device=DEV, dtype=torch.float32)

    def run(steps, name):
        cfg = _cfg(mb, **{"hyp.steps": steps, "hyp.grad_clip": 0.25, "impl.checkpoint.name": name,
                          "original_cwd": str(tmp_path), "hyp.warmup": 2})
        model = fresh()
        trainer = Trainer(model, HostBlockLoader(X, Y, mb), None, setup, cfg)
        while trainer.step_count < steps:
            trainer.step(validate=False)
            trainer.maybe_checkpoint()  # training.py:330-335, after the early-stop checks of the main loop
        return model, trainer

    m_full, _ = run(3, None)            # uninterrupted
    run(2, "ck.pth")                    # two steps, checkpointed
    ck = torch.load(tmp_path / "checkpoints" / "ck.pth")
    assert isinstance(ck, list) and len(ck) == 5 and ck[3] is None and ck[4] == 2
    assert list(ck[1].keys()) == list(fresh().state_dict().keys())
    assert "momentum_buffer" in ck[0]["state"][0]
    m_res, t_res = run(3, "ck.pth")     # resumes at step 2, runs the third
    assert t_res.stats["train_loss"] and len(t_res.stats["train_loss"]) == 1
    a = torch.cat([p.detach().reshape(-1) for p in m_full.parameters()])
    b = torch.cat([p.detach().reshape(-1) for p in m_res.parameters()])
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        run(2, "ck.pth")                # checkpoint is already at step 3 >= max_steps


def test_measure_floating_point_accuracy_reports_zero_drift():
    """measure_floating_point_accuracy.py / training.py:429-600: two evaluations from the same state."""
    from fullbatchtraining_b200.training import measure_implementation_noise

    mb, n = 16, 48
    X, Y = O.synthetic_cifar(n)
    rep = measure_implementation_noise(fresh(), HostBlockLoader(X, Y, mb), None, dict(device=DEV, dtype=torch.float32),
                                       _cfg(mb))
    assert rep["grad_l2"] > 0 and rep["diff_linf"] == 0.0 and rep["diff_l2"] == 0.0


@pytest.mark.parametrize("flips", [False, True], ids=["plain", "test_time_flips"])
def test_evaluate_on_kernels_matches_torch_eval(flips):
    """training.py:343-388 (`evaluate`): the kernel path (engine.forward_eval: convs on the tensor cores, eval-mode
    BatchNorm from the running statistics) against the same function through the torch modules, after one training
    step has moved the running statistics; 300 validation images = two microbatches + a zero-padded remainder."""
    from fullbatchtraining_b200.training import evaluate

    mb, n = 128, 256
    X, Y = O.synthetic_cifar(n)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(X, Y), batch_size=mb, shuffle=False)
    Xv, Yv = O.synthetic_cifar(300, seed=77)
    valid = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(Xv, Yv), batch_size=150, shuffle=False)
    setup = dict(device=DEV, dtype=torch.float32)
    cfg = _cfg(mb, **{"hyp.test_time_flips": flips})
    trainer = Trainer(fresh(), loader, valid, setup, cfg)
    trainer.step(validate=False)
    ref = evaluate(trainer.model, valid, None, setup, cfg.impl, cfg.hyp)
    got = evaluate(trainer.model, valid, None, setup, cfg.impl, cfg.hyp, engine=trainer.engine)
    assert got["valid_loss"][0] == pytest.approx(ref["valid_loss"][0], rel=2e-4)
    assert abs(got["valid_acc"][0] - ref["valid_acc"][0]) <= 1.0 / 300 + 1e-9
    # evaluation must not disturb training state: parameters, BN buffers
    before = [b.clone() for b in trainer.model.buffers()]
    evaluate(trainer.model, valid, None, setup, cfg.impl, cfg.hyp, engine=trainer.engine)
    for a, b in zip(before, trainer.model.buffers()):
        assert torch.equal(a, b)
