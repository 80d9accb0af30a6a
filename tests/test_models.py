"""Host-side checks of the reference-facing model factory (no GPU): same module tree / state_dict keys / parameter
order / seed-0 initialisation as the reference's construct_model (pinned through the golden fingerprints, which were
produced by the unmodified reference)."""
import json
import os

import numpy as np
import pytest
import torch

from fullbatchtraining_b200 import construct_model
from oracle import fb_oracle as O


def _meta(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z, json.loads(bytes(z["meta"]).decode())


@pytest.mark.parametrize("depth,golden", [(18, "r18_mb16_n32_f32"), (152, "r152_mb4_n8_f64")])
def test_init_matches_reference(golden_dir, depth, golden):
    z, meta = _meta(golden_dir, golden)
    torch.manual_seed(0)
    model = construct_model(dict(name=f"ResNet{depth}", depth=depth, initialization="skip-residual"), 3, 10)
    params = [p.detach() for p in model.parameters()]
    assert sum(p.numel() for p in params) == meta["num_params"]
    fp = O.fingerprint(params, meta["stride"])
    assert np.allclose(fp["norms"], z["init.norms"], rtol=1e-6)
    assert np.allclose(fp["sample"], z["init.sample"], rtol=1e-6, atol=1e-9)
    # key names and order equal the oracle's reference-ordered state
    torch.manual_seed(0)
    p, b = O.build_resnet_state(depth)
    assert [n for n, _ in model.named_parameters()] == list(p.keys())
    assert [n for n, _ in model.named_buffers()] == list(b.keys())


def test_forward_matches_oracle_forward():
    torch.manual_seed(0)
    model = construct_model(dict(name="ResNet18", depth=18), 3, 10).double()
    p = {k: v.detach().clone() for k, v in model.named_parameters()}
    b = {k: v.detach().clone() for k, v in model.named_buffers()}
    x, _ = O.synthetic_cifar(4, dtype=torch.float64)
    model.train()
    ref = O.OracleResNet(18, b).forward(p, x)
    assert torch.allclose(model(x), ref, rtol=1e-10, atol=1e-12)


def test_unsupported_configs_raise():
    with pytest.raises(ValueError):
        construct_model(dict(name="VGG11", depth=11), 3, 10)
    with pytest.raises(ValueError):
        construct_model(dict(name="ResNet18", depth=18, downsample="B"), 3, 10)
    with pytest.raises(ValueError):
        construct_model(dict(name="ResNet18", depth=20), 3, 10)
