"""Scheduler objects of the drop-in against the reference's: same learning rates, and checkpoints cross-load
(fixture tests/golden/ref_scheduler_state.pt is the state_dict of the reference's GradualWarmupScheduler +
CosineAnnealingLR stack after 6 steps, written by oracle/make_goldens.py from the unmodified reference)."""
import math
import os

import pytest
import torch

from fullbatchtraining_b200.config import default_cfg
from fullbatchtraining_b200.schedulers import LinearWarmup, build_scheduler


def make(warmup=3, sched="cosine-4000", steps=3000, lr=0.8):
    w = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.SGD([w], lr=lr, momentum=0.9, nesterov=True, weight_decay=5e-4)
    cfg = default_cfg({"hyp.warmup": warmup, "hyp.scheduler": sched, "hyp.steps": steps, "hyp.optim.lr": lr})
    return opt, build_scheduler(opt, cfg.hyp)


def run(opt, sched, n):
    lrs = []
    for _ in range(n):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    return lrs


def test_warmup_then_cosine_matches_the_reference_trace(golden_dir):
    ref = torch.load(os.path.join(golden_dir, "ref_scheduler_state.pt"), weights_only=False)
    opt, sched = make(warmup=ref["warmup"], lr=ref["base_lr"])
    assert isinstance(sched, LinearWarmup)
    assert run(opt, sched, 6) == pytest.approx(ref["lrs"], rel=1e-12, abs=1e-15)
    # scheduler.py:57-66: lr = base * t / W during warm-up (0 at step 0), the cosine starts at its epoch 0 at t = W + 1
    assert ref["lrs"][:4] == pytest.approx([0.0, 0.8 / 3, 1.6 / 3, 0.8])
    assert ref["lrs"][5] == pytest.approx(0.8 * (1 + math.cos(math.pi / 4000)) / 2)
    # same state_dict layout as the reference's
    mine = sched.state_dict()
    assert sorted(mine.keys()) == sorted(ref["state"].keys())
    assert sorted(mine["after_scheduler"].keys()) == sorted(ref["state"]["after_scheduler"].keys())
    for k in ("multiplier", "total_epoch", "finished", "last_epoch", "base_lrs", "_step_count"):
        assert mine[k] == ref["state"][k], k
    assert run(opt, sched, 4) == pytest.approx(ref["lrs_after"], rel=1e-12)


def test_reference_checkpoint_state_loads_and_continues(golden_dir):
    ref = torch.load(os.path.join(golden_dir, "ref_scheduler_state.pt"), weights_only=False)
    opt, sched = make(warmup=ref["warmup"], lr=ref["base_lr"])
    sched.load_state_dict(ref["state"])  # what training/utils.py:61 does with a reference checkpoint
    opt.param_groups[0]["lr"] = ref["state"]["_last_lr"][0]  # the optimizer's state_dict carries the lr (utils.py:60)
    assert sched.finished and sched.last_epoch == ref["state"]["last_epoch"]
    assert run(opt, sched, 4) == pytest.approx(ref["lrs_after"], rel=1e-12)
    # and our own state round-trips
    opt2, sched2 = make(warmup=ref["warmup"], lr=ref["base_lr"])
    run(opt2, sched2, 2)
    state = sched2.state_dict()
    opt3, sched3 = make(warmup=ref["warmup"], lr=ref["base_lr"])
    sched3.load_state_dict(state)
    opt3.param_groups[0]["lr"] = state["_last_lr"][0]
    assert run(opt3, sched3, 6) == pytest.approx(run(opt2, sched2, 6), rel=1e-12)


@pytest.mark.parametrize("sched,cls", [("cosine-4000", torch.optim.lr_scheduler.CosineAnnealingLR),
                                       ("cosine-decay", torch.optim.lr_scheduler.CosineAnnealingLR),
                                       ("linear", torch.optim.lr_scheduler.MultiStepLR),
                                       ("", torch.optim.lr_scheduler.MultiStepLR)])
def test_without_warmup_the_stock_torch_scheduler_is_used(sched, cls):
    opt, s = make(warmup=0, sched=sched, steps=40)
    assert type(s) is cls  # optimizers.py:69-87: its state_dict is torch's own
    lrs = run(opt, s, 40)
    if sched == "cosine-decay":
        assert lrs[20] == pytest.approx(0.8 * (1 + math.cos(math.pi * 20 / 40)) / 2)
    if sched == "linear":  # drops at steps // 2.667, // 1.6, // 1.142
        assert lrs[13] == pytest.approx(0.8) and lrs[14] == pytest.approx(0.08) and lrs[36] == pytest.approx(0.0008)
    if sched == "":
        assert lrs == pytest.approx([0.8] * 40)
    with pytest.raises(ValueError):
        make(sched="nonsense")
