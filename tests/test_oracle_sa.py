"""CPU: the oracle's SpatialAttention / UNet_SA restatement against fixtures generated from the unmodified reference
(tests/golden/make_golden_sa.py), and the drop-in UNet_SA constructor against the reference's seeded draws."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import unet_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def golden_sa():
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_sa_v1.pt"), weights_only=False)


def dropin_state(case):
    import unet.unet_model as UM
    nc, ncls, bil = case["cfg"][:3]
    torch.manual_seed(case["seed"])
    m = UM.UNet_SA(nc, ncls, bil)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("tag", ["sa_1_2_convT", "sa_3_3_bilinear"])
def test_unet_sa_constructor_and_oracle_step(golden_sa, tag):
    g = golden_sa[tag]
    nc, ncls, bil, B, H, W = g["cfg"]
    st = dropin_state(g)
    assert list(st) == g["keys"], "state_dict keys / order differ from the reference's UNet_SA"
    for k, v in st.items():
        assert float(v.double().sum()) == float(g["state_sum"][k]), f"seeded draw of {k} differs from the reference"
    img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    logits, loss, grads = O.training_step(st, img, msk, ncls, bil)
    assert O.rel_err(logits, g["logits"]) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert set(grads) == set(g["grad_norm"])
    for k, gr in grads.items():
        n = float(g["grad_norm"][k])
        assert abs(float(gr.double().norm()) - n) <= 2e-4 * max(n, 1e-12), k
        sm = gr.reshape(-1)[:: max(1, gr.numel() // 64)][:64]
        assert O.rel_err(sm, g["grad_sample"][k]) < 5e-4 or float(g["grad_sample"][k].abs().max()) < 1e-12, k


def test_gate_restatement(golden_sa):
    g = golden_sa["gate"]
    st = {"up9.attention.conv1.weight": g["w"].clone().requires_grad_(True)}
    x = g["x"].clone().requires_grad_(True)
    y = x * O.spatial_attention(st, "up9", x)
    y.backward(g["gy"])
    assert O.rel_err(y, g["y"]) < 1e-6
    assert O.rel_err(x.grad, g["gx"]) < 1e-5
    assert O.rel_err(st["up9.attention.conv1.weight"].grad, g["gw"]) < 1e-5
