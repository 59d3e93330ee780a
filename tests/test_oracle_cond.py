"""CPU: the oracle's well-conditioned state recipe and its bf16 storage model.

``golden_cond_v1.pt`` was written by tests/golden/make_golden_cond.py from the unmodified reference modules trained
with torch.optim.RMSprop; the recipe is chaotic (sign-like first RMSprop steps), so the bit-level fingerprint can
only be reproduced on a host whose oneDNN kernels round like the build container's.  The test therefore checks the
fingerprint tightly when the first tensor matches, and otherwise (a different CPU) the properties the GPU gates
rely on: the state is trained (loss, accuracy), BatchNorm is non-trivial, and reduced-precision storage lands
within north_star's own bf16 tolerances of the fp32 result."""
import os
import statistics
import warnings

import pytest
import torch

from oracle import unet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cond():
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_cond_v1.pt"), weights_only=False)


@pytest.fixture(scope="module")
def state(cond):
    r = cond["recipe"]
    n = torch.get_num_threads()
    torch.set_num_threads(8)             # oneDNN's reduction order follows the thread count; the fingerprint used 8
    try:
        return O.conditioned_state(1, 2, False, r["steps"], r["lr"], r["size"], r["batch"])
    finally:
        torch.set_num_threads(n)


def test_structured_batch_is_learnable_data():
    img, msk = O.structured_batch(2, 3, 4, 64, 48)
    assert img.shape == (2, 3, 64, 48) and msk.shape == (2, 64, 48) and msk.dtype == torch.long
    assert 0.0 <= img.min() and img.max() <= 1.0
    assert set(msk.unique().tolist()) <= {0, 1, 2, 3} and len(msk.unique()) >= 3
    # the grey level follows the class: means per class are ordered
    means = [img[:, 0][msk == c].mean().item() for c in msk.unique().tolist()]
    assert means == sorted(means)
    a, _ = O.structured_batch(2, 3, 4, 64, 48)
    assert torch.equal(a, img)


def test_conditioned_state_matches_reference_fingerprint(cond, state):
    g = cond["cond_1_2_convT"]
    r = cond["recipe"]
    same_host = all(abs(state[k].double().sum().item() - g["state_sum"][k].item()) <= 1e-9 * max(1.0, abs(g["state_sum"][k].item()))
                    for k in list(state)[:6])
    img2, msk2 = O.structured_batch(r["batch"], 1, 2, r["size"], r["size"], seed=9)
    n = torch.get_num_threads()
    torch.set_num_threads(8)
    logits, loss, grads = O.training_step({k: v.clone() for k, v in state.items()}, img2, msk2, 2, False)
    torch.set_num_threads(n)
    acc = (logits.argmax(1) == msk2).float().mean().item()
    if same_host:
        for k, v in g["state_sum"].items():
            assert abs(state[k].double().sum().item() - v.item()) <= 1e-9 * max(1.0, abs(v.item())), k
        assert O.rel_err(logits[:, :, ::8, ::8], g["logits_sample"]) < 1e-5
        assert abs(loss.item() - g["loss"].item()) < 1e-6
        for k, n in g["grad_norm"].items():
            assert abs(grads[k].norm().double().item() - n.item()) <= 2e-4 * n.item() + 1e-12, k
    else:
        warnings.warn("conditioned_state: this host's CPU kernels round differently from the build container's; "
                      "checking the state's properties instead of its fingerprint")
    assert loss.item() < 0.3 and acc > 0.97
    assert int(state["inc.double_conv.1.num_batches_tracked"]) == r["steps"]
    assert max(state[k].abs().max().item() for k in state if k.endswith(".bias") and "double_conv" in k) > 5e-3


def test_bf16_storage_model_meets_north_star_on_conditioned_state(cond, state):
    """north_star, bf16 mode: logits and gradients within 2e-2 relative, dice within 1e-3, argmax equal on >= 99.9 % of
    the pixels.  Storing the same arithmetic in bf16 achieves that on the conditioned state (and does NOT at random
    init, which is the point of having it): this is the attainable bar the CUDA path is gated against."""
    r = cond["recipe"]
    img, msk = O.structured_batch(r["batch"], 1, 2, r["size"], r["size"], seed=9)
    q = O.Rounding(torch.bfloat16)
    rl, rloss, rg = O.training_step({k: v.clone() for k, v in state.items()}, img, msk, 2, False)
    ql, qloss, qg = O.training_step({k: v.clone() for k, v in state.items()}, img, msk, 2, False, q=q)
    assert O.rel_err(ql, rl) < 2e-2
    assert (ql.argmax(1) == rl.argmax(1)).float().mean().item() >= 0.999
    assert abs(qloss.item() - rloss.item()) < 1e-3
    l2 = [O.rel_l2(qg[k], rg[k]) for k in rg]
    assert statistics.median(l2) < 2e-2 and max(l2) < 8e-2
    assert O.rel_l2(torch.cat([qg[k].reshape(-1) for k in rg]), torch.cat([rg[k].reshape(-1) for k in rg])) < 2e-2
    # contrast: the same comparison at random init on noise is an order of magnitude worse
    st0 = O.build_state(1, 2, False, seed=0)
    img0, msk0 = O.synthetic_batch(2, 1, 2, 64, 64)
    _, _, g0 = O.training_step({k: v.clone() for k, v in st0.items()}, img0, msk0, 2, False)
    _, _, q0 = O.training_step({k: v.clone() for k, v in st0.items()}, img0, msk0, 2, False, q=q)
    assert statistics.median(O.rel_l2(q0[k], g0[k]) for k in g0) > 1e-1


def test_rounding_model_is_identity_when_exact():
    st = O.build_state(1, 2, True, seed=0)
    img, msk = O.synthetic_batch(1, 1, 2, 32, 32)
    a = O.training_step({k: v.clone() for k, v in st.items()}, img, msk, 2, True)
    b = O.training_step({k: v.clone() for k, v in st.items()}, img, msk, 2, True, q=O.Rounding(None))
    assert torch.equal(a[0], b[0]) and all(torch.equal(a[2][k], b[2][k]) for k in a[2])
    x = torch.randn(64, requires_grad=True)
    y = O.Rounding(torch.bfloat16).act(x)
    assert torch.equal(y.detach(), x.detach().bfloat16().float())
    y.backward(torch.full_like(x, 1.0 + 2 ** -10))
    assert torch.equal(x.grad, torch.full_like(x, 1.0 + 2 ** -10).bfloat16().float())
    w = torch.randn(8, requires_grad=True)
    O.Rounding(torch.bfloat16).weight(w).backward(torch.full_like(w, 1.0 + 2 ** -10))
    assert torch.equal(w.grad, torch.full_like(w, 1.0 + 2 ** -10))
