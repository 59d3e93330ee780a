"""CPU: the oracle (oracle/unet_oracle.py) against the golden fixtures generated from the reference
itself by tests/golden/make_golden.py.  No access to /root/reference at run time."""
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O

CASES = {
    "unet_1_2_convT_32": (1, 2, False, 2, 32, 32),
    "unet_1_2_bilinear_32": (1, 2, True, 2, 32, 32),
    "unet_3_4_convT_48": (3, 4, False, 1, 48, 48),
}


def _check_unet(golden, tag, nc, ncls, bil, B, H, W):
    g = golden[tag]
    st = O.build_state(nc, ncls, bil, seed=0)
    for k, v in g["param_sum"].items():              # same random draws as the reference constructor
        assert abs(st[k].double().sum().item() - v.item()) < 1e-9 * max(1.0, abs(v.item())), k
    img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    logits, loss, grads = O.training_step(st, img, msk, ncls, bil)
    ref_logits = g["logits"]
    mine = logits if H <= 48 else logits[:, :, ::8, ::8]
    assert O.rel_err(mine, ref_logits) < 1e-5
    assert abs(loss.item() - g["loss"].item()) < 1e-6
    for k, n in g["grad_norm"].items():
        assert abs(grads[k].norm().double().item() - n.item()) <= 2e-4 * n.item() + 1e-9, k
        assert torch.allclose(grads[k].reshape(-1)[:32], g["grad_head"][k], rtol=2e-3, atol=2e-5 * grads[k].abs().max().item() + 1e-9), k
    for k, v in g["running"].items():
        assert torch.allclose(st[k].float(), v.float(), rtol=1e-5, atol=1e-6), k
    b = O.boundary_loss(logits, msk.float(), edge_width=5, edge_weight=7)
    assert abs(b.item() - g["boundary"].item()) < 1e-5


def test_unet_convT(golden):
    _check_unet(golden, "unet_1_2_convT_32", *CASES["unet_1_2_convT_32"])


def test_unet_bilinear(golden):
    _check_unet(golden, "unet_1_2_bilinear_32", *CASES["unet_1_2_bilinear_32"])


def test_unet_3_4(golden):
    _check_unet(golden, "unet_3_4_convT_48", *CASES["unet_3_4_convT_48"])


def test_unet_config1_256(golden):
    """BASELINE.json configs[0]: UNet(1,2,False) fp32 B=1 256x256 on CPU."""
    _check_unet(golden, "unet_1_2_convT_256_C1", 1, 2, False, 1, 256, 256)


def test_dice(golden):
    d = golden["dice"]
    for name in ("mc_loss", "bin_loss", "zero_loss"):
        c = d[name]
        x = c["input"].clone().requires_grad_(True)
        v = O.dice_loss(x, c["target"], multiclass=c["multiclass"])
        assert abs(v.item() - c["loss"].item()) < 1e-7, name
        v.backward()
        assert torch.allclose(x.grad, c["grad"], rtol=1e-5, atol=1e-9), name
    assert abs(O.dice_coeff(d["coeff_nobatch"]["input"], d["coeff_nobatch"]["target"]).item()
               - d["coeff_nobatch"]["value"].item()) < 1e-7
    assert O.dice_coeff(d["coeff_empty"]["input"], d["coeff_empty"]["target"]).item() == 1.0
    assert abs(O.multiclass_dice_coeff(d["mc_coeff_nobatch"]["input"], d["mc_coeff_nobatch"]["target"]).item()
               - d["mc_coeff_nobatch"]["value"].item()) < 1e-7


def test_multiclass_dice_is_one_global_ratio():
    """with multiclass=True the loss is a single ratio over B*C*H*W."""
    g = torch.Generator().manual_seed(0)
    p = torch.softmax(torch.randn(2, 3, 8, 8, generator=g), 1)
    t = F.one_hot(torch.randint(0, 3, (2, 8, 8), generator=g), 3).permute(0, 3, 1, 2).float()
    inter = 2 * (p * t).sum()
    expect = 1 - (inter + 1e-6) / (p.sum() + t.sum() + 1e-6)
    assert abs(O.dice_loss(p, t, multiclass=True).item() - expect.item()) < 1e-7


def test_boundary(golden):
    b = golden["boundary"]
    for name, c in b.items():
        if name == "bce_constants":
            continue
        v = O.boundary_loss(c["pred"], c["target"], c["edge_width"], c["edge_weight"])
        assert abs(v.item() - c["value"].item()) < 1e-6, name
        assert O.boundary_counts(c["pred"], c["target"], c["edge_width"]) == c["counts"], name
        cf = O.boundary_loss_counts(c["pred"], c["target"], c["edge_width"], c["edge_weight"])
        assert abs(cf - c["value"].item()) < 2e-6, name
        assert not v.requires_grad
    assert torch.allclose(torch.tensor(O._bce_constants(), dtype=torch.float64), b["bce_constants"])


def test_parts(golden):
    """Per-part fixtures: DoubleConv / Down / Up / OutConv forward + backward."""
    def run(tag, fn):
        c = golden[tag]
        st = {k: v.clone() for k, v in c["state"].items()}
        names = [k for k in st if st[k].is_floating_point() and "running" not in k]
        for k in names:
            st[k].requires_grad_(True)
        ins = [t.clone().requires_grad_(True) for t in c["inputs"]]
        out = fn(st, *ins)
        assert O.rel_err(out, c["out"]) < 1e-5, tag
        out.backward(c["gout"])
        for a, b in zip(ins, c["gin"]):
            assert O.rel_err(a.grad, b) < 1e-4, tag
        for k in names:
            assert O.rel_err(st[k].grad, c["gparams"][k]) < 1e-4, (tag, k)

    run("DoubleConv_4_8", lambda st, x: O.double_conv(st, "double_conv", x))
    run("DoubleConv_4_8_mid6", lambda st, x: O.double_conv(st, "double_conv", x))
    run("Down_4_8", lambda st, x: O.double_conv(st, "maxpool_conv.1.double_conv", F.max_pool2d(x, 2)))
    run("Down_4_8_odd", lambda st, x: O.double_conv(st, "maxpool_conv.1.double_conv", F.max_pool2d(x, 2)))

    def up_fn(bilinear):
        def f(st, x1, x2):
            st2 = {("u." + k): v for k, v in st.items()}
            return O.up(st2, "u", x1, x2, bilinear)
        return f
    run("Up_8_4_convT", up_fn(False))
    run("Up_8_4_convT_pad", up_fn(False))
    run("Up_8_4_bilinear", up_fn(True))
    run("Up_8_4_bilinear_pad", up_fn(True))
    run("OutConv_8_3", lambda st, x: F.conv2d(x, st["conv.weight"], st["conv.bias"]))
