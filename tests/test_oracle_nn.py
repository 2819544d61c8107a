"""The functional fp32 NN restatement (oracle/nn_ref.py) reproduces the committed outputs of the
UNMODIFIED reference PolicyValueNet (tests/golden/nn_golden.npz) and, where /root/reference exists,
is compared live with the reference module."""
import json
import os

import numpy as np
import pytest
import torch

from matrix0_b200.model import NetConfig, parameter_shapes
from oracle import nn_ref, refload


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    d = json.loads(str(g[f"{name}_cfg"]))
    known = set(NetConfig.__dataclass_fields__)
    cfg = NetConfig(**{k: v for k, v in d.items() if k in known})
    sd = nn_ref.make_state_dict(parameter_shapes(cfg), seed=1)
    return g, cfg, sd


@pytest.mark.parametrize("name", ["small", "r24"])
def test_restatement_reproduces_reference_outputs(golden_dir, name):
    g, cfg, sd = load_case(golden_dir, name)
    with torch.no_grad():
        p, v, ssl = nn_ref.forward(sd, cfg, torch.from_numpy(g[f"{name}_x"]), return_ssl=True)
    assert np.abs(p.numpy() - g[f"{name}_logits"]).max() < 2e-5
    assert np.abs(v.numpy() - g[f"{name}_values"]).max() < 2e-5
    for t, s in ssl.items():
        assert np.abs(s.numpy() - g[f"{name}_ssl_{t}"]).max() < 2e-5


@pytest.mark.skipif(not refload.reference_available(), reason="/root/reference not present")
def test_restatement_matches_reference_module_live(golden_dir):
    resnet = refload.load_reference("model.resnet")
    g, cfg, sd = load_case(golden_dir, "small")
    d = json.loads(str(g["small_cfg"]))
    ref = resnet.PolicyValueNet.from_config(d).eval()
    ref.load_state_dict(sd, strict=False)
    x = torch.rand(5, 19, 8, 8, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        p1, v1 = ref(x)
        p2, v2 = nn_ref.forward(sd, cfg, x)
    assert (p1 - p2).abs().max() < 2e-5 and (v1 - v2).abs().max() < 2e-5
    assert sorted(k for k in ref.state_dict() if k in sd) == sorted(sd)  # every key we model exists in the reference
