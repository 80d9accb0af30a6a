"""The oracle restatement (oracle/fb_oracle.py) must reproduce fixtures produced by the unmodified reference
(oracle/make_goldens.py -> tests/golden/*.npz).  fp64 fixtures are reproducible across machines and are checked
tightly; fp32 fixtures depend on the CPU conv backend (~1e-2 between backends) and are checked loosely."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fb_oracle as O


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def check_fp(z, prefix, tensors, stride, rtol):
    fp = O.fingerprint(tensors, stride)
    scale = z[f"{prefix}.total_norm"]
    assert abs(fp["total_norm"] - scale) <= rtol * scale, prefix
    n = np.linalg.norm(z[f"{prefix}.norms"])
    assert np.linalg.norm(fp["norms"] - z[f"{prefix}.norms"]) <= rtol * n, prefix
    s = np.linalg.norm(z[f"{prefix}.sample"])
    assert np.linalg.norm(fp["sample"] - z[f"{prefix}.sample"]) <= rtol * s, prefix


CASES = [("r18_mb16_n32_f64", 1e-9), ("r18_mb16_n32_f32", 5e-2), ("r152_mb4_n8_f64", 1e-9),
         ("r18_mb16_n48_f64_central", 1e-9), ("r18_mb16_n48_f64_legacy", 1e-9), ("r18_mb16_n48_f64_acc", 1e-9),
         ("r18_mb16_n48_f64_central_acc", 1e-9),
         pytest.param("r18_mb128_n256_f64", 1e-9, marks=pytest.mark.slow)]


@pytest.mark.parametrize("name,rtol", CASES)
def test_oracle_matches_reference_golden(golden_dir, name, rtol):
    z, meta = load(golden_dir, name)
    dt = getattr(torch, meta["dtype"])
    torch.manual_seed(0)
    p, b = O.build_resnet_state(meta["depth"], dtype=dt)
    assert sum(v.numel() for v in p.values()) == meta["num_params"]
    # initialisation is bit-identical to the reference's construct_model under the same seed
    check_fp(z, "init", list(p.values()), meta["stride"], 1e-12 if dt == torch.float64 else 1e-6)
    X, Y = O.synthetic_cifar(meta["n"], dtype=dt)
    extra = meta.get("extra", {})
    out = O.full_batch_step(meta["depth"], p, b, X, Y, meta["mb"], keep_microbatches=2, **meta["hyp"], **extra)
    check_fp(z, "avg", out["avg"], meta["stride"], rtol)
    for i, kept in enumerate(out["kept"]):
        check_fp(z, f"mb{i}.raw", kept["raw"], meta["stride"], rtol)
        check_fp(z, f"mb{i}.reg", kept["reg"], meta["stride"], rtol)
    bufs = [v for k, v in b.items() if not k.endswith("num_batches_tracked")]
    check_fp(z, "buffers", bufs, 97, max(rtol, 1e-6))
    sc = meta["scalars"]
    assert abs(float(out["loss"]) - sc["train_loss"]) <= max(rtol, 1e-6) * abs(sc["train_loss"])
    assert float(out["correct"]) / (out["K"] * meta["mb"]) == pytest.approx(sc["train_acc"])
    gn = out["grad_norms"].sqrt().tolist()
    assert np.allclose(gn, sc["grad_norm_train"], rtol=max(rtol, 1e-6))
    # training.py:95-98 full_loss = loss + wd/2 |theta|^2 + lr/4*bs*mean(grad_norms)
    full = float(out["loss"]) + 0.5 * 5e-4 * float(out["param_norm"]) \
        + meta["hyp"]["lr"] / 4 * meta["hyp"]["block_strength"] * float(out["grad_norms"].mean())
    if extra.get("acc_strength", 0.0) != 0:  # training.py:99-102
        full += meta["hyp"]["lr"] / 4 * extra["acc_strength"] * float(sum(g.pow(2).sum() for g in out["pre_grads"]))
    assert full == pytest.approx(sc["full_loss"], rel=max(rtol, 1e-6))


def test_oracle_sgd_branch_matches_reference_golden(golden_dir):
    """training.py:241-286 restated (fb_oracle.sgd_epochs) against two steps of the unmodified reference's stochastic
    branch: parameters after 8 optimizer steps, per-step loss / accuracy / gradient norm, BatchNorm statistics."""
    z, meta = load(golden_dir, "r18_sgd_mb16_n64_f64")
    dt = torch.float64
    torch.manual_seed(0)
    p, b = O.build_resnet_state(18, dtype=dt)
    check_fp(z, "init", list(p.values()), meta["stride"], 1e-12)
    X, Y = O.synthetic_cifar(meta["n"], dtype=dt)
    out = O.sgd_epochs(18, p, b, X, Y, meta["mb"], meta["extra"]["steps"], meta["hyp"]["lr"],
                       block_strength=meta["hyp"]["block_strength"], eps=meta["hyp"]["eps"])
    check_fp(z, "theta", list(p.values()), meta["stride"], 1e-9)
    sc = meta["scalars"]
    assert out["train_loss"] == pytest.approx(sc["train_loss_steps"], rel=1e-9)
    assert out["train_acc"] == pytest.approx(sc["train_acc_steps"])
    assert [v ** 0.5 for v in out["grad_norm_sq"]] == pytest.approx(sc["grad_norm_steps"], rel=1e-9)
    bufs = [v for k, v in b.items() if not k.endswith("num_batches_tracked")]
    check_fp(z, "buffers", bufs, 97, 1e-8)


def test_microbatch_count_and_drop_last():
    # data_preparation.py:68 drop_last -> K = N // mb
    torch.manual_seed(0)
    p, b = O.build_resnet_state(18, dtype=torch.float64)
    X, Y = O.synthetic_cifar(20, dtype=torch.float64)
    out = O.full_batch_step(18, p, b, X, Y, 8, lr=0.8, block_strength=0.0)
    assert out["K"] == 2 and out["grad_norms"].shape[0] == 2


def test_label_smoothing_loss_matches_torch():
    # modules.py:96-101 with smoothing 0 equals CrossEntropyLoss
    g = torch.Generator().manual_seed(1)
    z = torch.randn(7, 10, generator=g, dtype=torch.float64)
    y = torch.randint(0, 10, (7,), generator=g)
    assert torch.allclose(O.label_smooth_xent(z, y, 0.0), torch.nn.functional.cross_entropy(z, y))
    ls = O.label_smooth_xent(z, y, 0.1)
    lp = torch.log_softmax(z, -1)
    w = torch.full_like(z, 0.1 / 9)
    w[torch.arange(7), y] = 0.9
    assert torch.allclose(ls, -(w * lp).sum(-1).mean())
