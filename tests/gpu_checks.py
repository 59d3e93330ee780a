"""GPU parity checks of the CUDA path (through the C ABI) against the CPU oracle / plain PyTorch-CPU
fp32 references.  Used by the ``-m gpu`` pytest files and by ``tests/run_gpu_diag.py`` (which runs
every group in its own process and prints a table instead of stopping at the first failure).

Each check returns a list of (label, error, tolerance).  Tolerances (north_star): fp32 exact mode
1e-3 relative (we hold 1e-4 on single ops), TF32 1e-3 end to end, bf16 2e-2 relative, dice 1e-3.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import unet_oracle as O  # noqa: E402
from unetb200 import _lib, ops  # noqa: E402
from unetb200 import functional as UF  # noqa: E402
from unetb200 import losses as UL  # noqa: E402

DEV = "cuda"
BF, FP = torch.bfloat16, torch.float32
TOL = {FP: 2e-5, BF: 1.2e-2}


def gen(seed):
    return torch.Generator().manual_seed(seed)


def dev_nhwc(x, dtype):
    return ops.to_nhwc(x.to(DEV), dtype)


def host(t):
    return t.detach().float().cpu()


def rq(x, dtype):
    """round a CPU fp32 tensor through `dtype` (what the device stores); always a fresh tensor"""
    return x.detach().to(dtype).float().clone()


def rel(a, b):
    return O.rel_err(a, b)


def in_slice(x, dtype, extra=8):
    """Place x (CPU, NCHW) into channels [extra/2, ...) of a wider NHWC device buffer -> slice view."""
    B, C, H, W = x.shape
    buf = ops.empty_nhwc(B, C + extra, H, W, dtype, DEV)
    buf.fill_(7.0)
    s = ops.channel_slice(buf, extra // 2 if (extra // 2) % 8 == 0 else 0, C)
    s.copy_(x.to(DEV).to(dtype))
    return s


# ------------------------------------------------------------------------------------------------
# elementwise group
# ------------------------------------------------------------------------------------------------
def check_layout_ops():
    out = []
    g = gen(0)
    x = torch.randn(2, 5, 6, 7, generator=g)
    for dt_ in (FP, BF):
        a = dev_nhwc(x, dt_)                                   # NCHW contiguous -> gather
        out.append((f"gather_nchw_{dt_}", rel(host(a), rq(x, dt_)), 1e-7))
        b = dev_nhwc(x.contiguous(memory_format=torch.channels_last), dt_)
        out.append((f"gather_cl_{dt_}", rel(host(b), rq(x, dt_)), 1e-7))
    for C in (16, 5):
        x = torch.randn(2, C, 6, 7, generator=g)
        y = torch.randn(2, C, 6, 7, generator=g)
        for dt_ in (FP, BF):
            s = in_slice(x, dt_, extra=16)
            d = ops.empty_nhwc(2, C, 6, 7, dt_, DEV)
            ops.copy_channels(s, d)
            out.append((f"copy_slice_C{C}_{dt_}", rel(host(d), rq(x, dt_)), 1e-7))
            ops.add_channels_(d, dev_nhwc(y, dt_))
            out.append((f"add_C{C}_{dt_}", rel(host(d), rq(rq(x, dt_) + rq(y, dt_), dt_)), 1e-7))
            cs = ops.channel_sum(s)
            out.append((f"channel_sum_C{C}_{dt_}", rel(host(cs), rq(x, dt_).sum((0, 2, 3))), 1e-5))
            ops.zero_channels(s)
            out.append((f"zero_C{C}_{dt_}", host(s).abs().max().item(), 0.0))
    return out


def _bn_ref(y, gamma, beta, eps=1e-5):
    mean = y.mean((0, 2, 3))
    var = y.var((0, 2, 3), unbiased=False)
    invstd = 1 / torch.sqrt(var + eps)
    return mean, var, invstd


def check_bn_forward():
    out = []
    g = gen(1)
    for (B, C, H, W) in ((2, 16, 10, 12), (1, 6, 9, 7), (2, 64, 8, 8)):
        for dt_ in (FP, BF):
            y = rq(torch.randn(B, C, H, W, generator=g) * 2 + 0.5, dt_)
            gamma = torch.rand(C, generator=g) + 0.5
            beta = torch.randn(C, generator=g)
            rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
            rm0, rv0 = rm.clone(), rv.clone()
            zref = F.relu(F.batch_norm(y, rm, rv, gamma, beta, True, 0.1, 1e-5))     # updates rm, rv
            yd = in_slice(y, dt_, extra=8) if C % 8 == 0 else dev_nhwc(y, dt_)
            stats = torch.stack([y.double().sum((0, 2, 3)), (y.double() ** 2).sum((0, 2, 3))]).to(DEV).reshape(-1)
            drm, drv = rm0.to(DEV), rv0.to(DEV)
            coefs = ops.bn_finalize(stats, B * H * W, gamma.to(DEV), beta.to(DEV), 1e-5, 0.1, drm, drv, C)
            mean, var, invstd = _bn_ref(y, gamma, beta)
            out.append((f"bn_finalize_mean_{C}_{dt_}", rel(host(coefs[0]), mean), 1e-5))
            out.append((f"bn_finalize_invstd_{C}_{dt_}", rel(host(coefs[1]), invstd), 1e-5))
            out.append((f"bn_running_mean_{C}_{dt_}", rel(host(drm), rm), 1e-5))
            out.append((f"bn_running_var_{C}_{dt_}", rel(host(drv), rv), 1e-5))
            z = ops.empty_nhwc(B, C, H, W, dt_, DEV)
            p = ops.empty_nhwc(B, C, H // 2, W // 2, dt_, DEV)
            ops.bn_relu_apply(yd, coefs, z, p)
            out.append((f"bn_apply_pool_z_{C}_{dt_}", rel(host(z), zref), TOL[dt_]))
            out.append((f"bn_apply_pool_p_{C}_{dt_}", rel(host(p), F.max_pool2d(host(z), 2)), 1e-7))
            z2 = ops.empty_nhwc(B, C, H, W, dt_, DEV)
            ops.bn_relu_apply(yd, coefs, z2, None)
            out.append((f"bn_apply_z_{C}_{dt_}", rel(host(z2), host(z)), 1e-7))
            ce = ops.bn_eval_coeffs(gamma.to(DEV), beta.to(DEV), rm.to(DEV), rv.to(DEV), 1e-5, C)
            ops.bn_relu_apply(yd, ce, z2, None)
            zev = F.relu(F.batch_norm(y, rm, rv, gamma, beta, False, 0.1, 1e-5))
            out.append((f"bn_eval_{C}_{dt_}", rel(host(z2), zev), TOL[dt_]))
    return out


def check_maxpool():
    out = []
    g = gen(2)
    for (B, C, H, W) in ((2, 8, 8, 10), (1, 5, 7, 9), (2, 16, 6, 6)):
        for dt_ in (FP, BF):
            # few distinct values -> many ties: exercises the first-max rule
            x = torch.randint(0, 3, (B, C, H, W), generator=g).float()
            xr = x.clone().requires_grad_(True)
            pref = F.max_pool2d(xr, 2)
            gp = rq(torch.randn(pref.shape, generator=g), dt_)
            pref.backward(gp)
            xd = dev_nhwc(x, dt_)
            p = ops.empty_nhwc(B, C, H // 2, W // 2, dt_, DEV)
            ops.maxpool2_fwd(xd, p)
            out.append((f"maxpool_fwd_{C}_{H}x{W}_{dt_}", rel(host(p), pref.detach()), 1e-7))
            gx = ops.empty_nhwc(B, C, H, W, dt_, DEV)
            gx.fill_(3.0)
            ops.maxpool2_bwd(xd, dev_nhwc(gp, dt_), gx, accumulate=False)
            out.append((f"maxpool_bwd_ties_{C}_{H}x{W}_{dt_}", rel(host(gx), xr.grad), 1e-7))
            base = rq(torch.randn(B, C, H, W, generator=g), dt_)
            gx2 = dev_nhwc(base, dt_).clone()
            ops.maxpool2_bwd(xd, dev_nhwc(gp, dt_), gx2, accumulate=True)
            out.append((f"maxpool_bwd_acc_{C}_{H}x{W}_{dt_}", rel(host(gx2), rq(base + xr.grad, dt_)), 1e-7))
    return out


def check_bn_backward():
    out = []
    g = gen(3)
    for (B, C, H, W) in ((2, 16, 10, 12), (1, 6, 9, 7), (3, 128, 6, 6), (2, 520, 4, 4)):
        for dt_ in (FP, BF):
            for training in (True, False):
                y = rq(torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3, dt_).requires_grad_(True)
                gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
                beta = torch.randn(C, generator=g).requires_grad_(True)
                rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
                z = F.relu(F.batch_norm(y, rm.clone(), rv.clone(), gamma, beta, training, 0.1, 1e-5))
                gz = rq(torch.randn(z.shape, generator=g), dt_)
                z.backward(gz)
                yd = dev_nhwc(y.detach(), dt_)
                if training:
                    stats = torch.stack([y.detach().double().sum((0, 2, 3)),
                                         (y.detach().double() ** 2).sum((0, 2, 3))]).to(DEV).reshape(-1)
                    coefs = ops.bn_finalize(stats, B * H * W, gamma.detach().to(DEV), beta.detach().to(DEV), 1e-5,
                                            0.0, None, None, C)
                else:
                    coefs = ops.bn_eval_coeffs(gamma.detach().to(DEV), beta.detach().to(DEV), rm.to(DEV), rv.to(DEV),
                                               1e-5, C)
                gy, dg, db = ops.bn_relu_bwd(dev_nhwc(gz, dt_), yd, coefs, training)
                tag = f"{C}_{dt_}_{'train' if training else 'eval'}"
                out.append((f"bn_bwd_gy_{tag}", rel(host(gy), y.grad), TOL[dt_]))
                out.append((f"bn_bwd_dgamma_{tag}", rel(host(dg), gamma.grad), 1e-4))
                out.append((f"bn_bwd_dbeta_{tag}", rel(host(db), beta.grad), 1e-4))
    return out


def check_upsample():
    out = []
    g = gen(4)
    for (B, C, h, w, Ho, Wo) in ((2, 8, 5, 6, 10, 12), (1, 3, 4, 7, 9, 15), (2, 16, 1, 3, 3, 6), (1, 8, 6, 5, 13, 12)):
        for dt_ in (FP, BF):
            x = rq(torch.randn(B, C, h, w, generator=g), dt_).requires_grad_(True)
            up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
            dy, dx = Ho - 2 * h, Wo - 2 * w
            ref = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
            gy = rq(torch.randn(ref.shape, generator=g), dt_)
            ref.backward(gy)
            yd = ops.empty_nhwc(B, C, Ho, Wo, dt_, DEV)
            ops.zero_channels(yd)
            ops.upsample2x_fwd(dev_nhwc(x.detach(), dt_), yd, (dy // 2, dx // 2))
            out.append((f"upsample_fwd_{C}_{h}x{w}_{dt_}", rel(host(yd), ref.detach()), TOL[dt_]))
            gx = ops.empty_nhwc(B, C, h, w, dt_, DEV)
            ops.upsample2x_bwd(dev_nhwc(gy, dt_), gx, (dy // 2, dx // 2))
            out.append((f"upsample_bwd_{C}_{h}x{w}_{dt_}", rel(host(gx), x.grad), TOL[dt_]))
    return out


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def check_ce_dice():
    out = []
    g = gen(5)
    for (B, K, H, W) in ((2, 2, 16, 20), (1, 4, 9, 7), (3, 3, 32, 32)):
        for dt_ in (FP, BF):
            lg = rq(torch.randn(B, K, H, W, generator=g) * 2, dt_).requires_grad_(True)
            tg = torch.randint(0, K, (B, H, W), generator=g)
            ref = O.train_loss(lg, tg, K)
            (ref * 1.7).backward()
            x = lg.detach().to(DEV).to(dt_).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            total, parts = UL.ce_dice_loss(x, tg.to(DEV), return_parts=True)
            (total * 1.7).backward()
            out.append((f"ce_dice_loss_{K}_{dt_}", abs(total.item() - ref.item()) / abs(ref.item()), 2e-6))
            ce = F.cross_entropy(lg.detach(), tg)
            out.append((f"ce_part_{K}_{dt_}", abs(parts[1].item() - ce.item()), 2e-6))
            out.append((f"ce_dice_grad_{K}_{dt_}", rel(host(x.grad), lg.grad), TOL[dt_] if dt_ == BF else 1e-5))
    return out


def check_dice(golden):
    out = []
    d = golden["dice"]
    for name in ("mc_loss", "bin_loss", "zero_loss"):
        c = d[name]
        x = c["input"].to(DEV).requires_grad_(True)
        v = UL.dice_loss(x, c["target"].to(DEV), multiclass=c["multiclass"])
        v.backward()
        out.append((f"dice_{name}", abs(v.item() - c["loss"].item()), 1e-6))
        out.append((f"dice_{name}_grad", (host(x.grad) - c["grad"]).abs().max().item(), 1e-8 + 1e-5 * c["grad"].abs().max().item()))
    for name in ("coeff_nobatch", "coeff_empty"):
        c = d[name]
        v = UL.dice_coeff(c["input"].to(DEV), c["target"].to(DEV), reduce_batch_first=False)
        out.append((f"dice_{name}", abs(v.item() - c["value"].item()), 1e-6))
    c = d["mc_coeff_nobatch"]
    v = UL.multiclass_dice_coeff(c["input"].to(DEV), c["target"].to(DEV), reduce_batch_first=False)
    out.append(("dice_mc_coeff_nobatch", abs(v.item() - c["value"].item()), 1e-6))
    # channels_last probabilities + one-hot view, as train.py:138-142 produces them
    g = gen(6)
    lg = torch.randn(2, 3, 12, 10, generator=g).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    tg = torch.randint(0, 3, (2, 12, 10), generator=g).to(DEV)
    p = F.softmax(lg, dim=1).float()
    oh = F.one_hot(tg, 3).permute(0, 3, 1, 2).float()
    v = UL.dice_loss(p, oh, multiclass=True)
    v.backward()
    lc = lg.detach().cpu().requires_grad_(True)
    r = O.dice_loss(F.softmax(lc, dim=1), F.one_hot(tg.cpu(), 3).permute(0, 3, 1, 2).float(), multiclass=True)
    r.backward()
    out.append(("dice_train_form", abs(v.item() - r.item()), 1e-6))
    out.append(("dice_train_form_grad", rel(host(lg.grad), lc.grad), 1e-4))
    return out


def check_boundary(golden):
    out = []
    for name, c in golden["boundary"].items():
        if name == "bce_constants":
            continue
        v = UL.boundary_loss(c["pred"].to(DEV), c["target"].to(DEV), c["edge_width"], c["edge_weight"])
        out.append((f"boundary_{name}", abs(v.item() - c["value"].item()), 2e-6 * max(1.0, abs(c["value"].item()))))
        out.append((f"boundary_{name}_nograd", float(v.requires_grad), 0.0))
    # train.py call forms on NHWC logits (strided channel-1 view), fp32 and bf16, class-index targets
    g = gen(7)
    for dt_ in (FP, BF):
        lg = rq(torch.randn(4, 2, 96, 80, generator=g) * 6, dt_)
        tg = torch.randint(0, 2, (4, 96, 80), generator=g)
        tg255 = tg * 255
        xd = lg.to(DEV).to(dt_).contiguous(memory_format=torch.channels_last)
        for t_cpu, tname in ((tg.float(), "idx"), (tg255.float(), "255")):
            ref = O.boundary_loss(lg.to(dt_), t_cpu, edge_width=11, edge_weight=7)
            v = UL.boundary_loss(xd, t_cpu.to(DEV), edge_width=11, edge_weight=7)
            out.append((f"boundary_logits_{tname}_{dt_}", abs(v.item() - float(ref)), 3e-6 * max(1.0, abs(float(ref)))))
        v2 = UL.boundary_loss(xd, tg255.to(DEV), edge_width=11, edge_weight=7)                 # int64 target
        ref = O.boundary_loss(lg.to(dt_), tg255.float(), edge_width=11, edge_weight=7)
        out.append((f"boundary_int64_target_{dt_}", abs(v2.item() - float(ref)), 3e-6 * max(1.0, abs(float(ref)))))
    return out


# ------------------------------------------------------------------------------------------------
# OutConv
# ------------------------------------------------------------------------------------------------
def check_outconv():
    out = []
    g = gen(8)
    for (B, C, K, H, W) in ((2, 64, 2, 12, 10), (1, 8, 3, 7, 9), (2, 6, 4, 5, 5), (1, 16, 1, 8, 8)):
        for dt_ in (FP, BF):
            x = rq(torch.randn(B, C, H, W, generator=g), dt_).requires_grad_(True)
            w = torch.randn(K, C, 1, 1, generator=g) * 0.3
            b = torch.randn(K, generator=g)
            wq, bq = rq(w, dt_).requires_grad_(True), rq(b, dt_).requires_grad_(True)
            ref = F.conv2d(x, wq, bq)
            gl = rq(torch.randn(ref.shape, generator=g), dt_)
            ref.backward(gl)
            xd = dev_nhwc(x.detach(), dt_).requires_grad_(True)
            wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
            y = UF.OutConvFn.apply(xd, wd, bd, UF._Cfg(save=True))
            y.backward(gl.to(DEV).to(dt_))
            tag = f"{C}to{K}_{dt_}"
            out.append((f"outconv_fwd_{tag}", rel(host(y), ref.detach()), TOL[dt_]))
            out.append((f"outconv_gx_{tag}", rel(host(xd.grad), x.grad), TOL[dt_]))
            out.append((f"outconv_dw_{tag}", rel(host(wd.grad), wq.grad), 2e-4))
            out.append((f"outconv_db_{tag}", rel(host(bd.grad), bq.grad), 2e-4))
    return out


# ------------------------------------------------------------------------------------------------
# generalised convolution: SIMT and tcgen05 engines
# ------------------------------------------------------------------------------------------------
def _conv3x3_case(B, Ci, Co, H, W, dt_, algo, seed, slice_io=False):
    """fprop (+BN stats), dgrad and wgrad of a 3x3 conv through gconv with `algo`."""
    res = []
    g = gen(seed)
    x = rq(torch.randn(B, Ci, H, W, generator=g), dt_).requires_grad_(True)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5))
    wq = rq(w, dt_).requires_grad_(True)
    ref = F.conv2d(x, wq, padding=1)
    gy = rq(torch.randn(ref.shape, generator=g), dt_)
    ref.backward(gy)
    tol = 2e-5 if dt_ == FP and algo == _lib.ALGO_SIMT else (2e-3 if dt_ == FP else 1.2e-2)
    if dt_ == FP and ops.x3_mode() and algo != _lib.ALGO_SIMT:
        tol = 5e-5                     # 3xTF32: fp32-level results from the tensor cores
    tag = f"{Ci}to{Co}_{H}x{W}_{str(dt_)[6:]}_{'simt' if algo == _lib.ALGO_SIMT else 'tc'}{'_slice' if slice_io else ''}"
    xd = in_slice(x.detach(), dt_, 64) if slice_io else dev_nhwc(x.detach(), dt_)
    wdev = w.to(DEV)
    y = ops.empty_nhwc(B, Co, H, W, dt_, DEV)
    if slice_io:
        ybuf = ops.empty_nhwc(B, Co + 64, H, W, dt_, DEV)
        ybuf.fill_(5.0)
        y = ops.channel_slice(ybuf, 64, Co)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)

    def desc(xin, Cout, yout):
        d = ops.make_gconv(ops._DT[dt_], algo, B, H, W, xin.shape[1], ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(xin),
                           Cout, 1, 1, (0, 0), H, W, ops.nhwc_ld(yout))
        d.algo = algo
        return d
    used = ops.gconv_fprop(desc(xd, Co, y), xd, UF.pack3x3_fprop(wdev, dt_), None, y, stats)
    assert used == algo, f"expected algo {algo}, library used {used}"
    yh = host(y)
    res.append((f"fprop_{tag}", rel(yh, ref.detach()), tol))
    if slice_io:
        res.append((f"fprop_{tag}_untouched", (host(ops.channel_slice(ybuf, 0, 64)) - 5.0).abs().max().item(), 0.0))
    st = host(stats.float()).reshape(2, Co)
    res.append((f"stats_sum_{tag}", rel(st[0], yh.sum((0, 2, 3))), 1e-4))
    res.append((f"stats_sq_{tag}", rel(st[1], (yh ** 2).sum((0, 2, 3))), 1e-4))
    gyd = dev_nhwc(gy, dt_)
    gx = ops.empty_nhwc(B, Ci, H, W, dt_, DEV)
    ops.gconv_fprop(desc(gyd, Ci, gx), gyd, UF.pack3x3_dgrad(wdev, dt_), None, gx, None)
    res.append((f"dgrad_{tag}", rel(host(gx), x.grad), tol))
    dW = torch.empty(Co, Ci, 3, 3, device=DEV)
    used = ops.gconv_wgrad(desc(xd, Co, gyd), xd, gyd, dW, 1, 9, Ci * 9)
    assert used == algo, f"wgrad: expected algo {algo}, library used {used}"
    res.append((f"wgrad_{tag}", rel(host(dW), wq.grad), 2e-4 if tol < 1e-3 else 5e-3))
    return res


def _convT_case(B, Ci, Co, h, w_, pad, dt_, algo, seed):
    res = []
    g = gen(seed)
    x = rq(torch.randn(B, Ci, h, w_, generator=g), dt_).requires_grad_(True)
    w = torch.randn(Ci, Co, 2, 2, generator=g) / (Ci ** 0.5)
    b = torch.randn(Co, generator=g)
    wq, bq = rq(w, dt_).requires_grad_(True), rq(b, dt_).requires_grad_(True)
    H, W = 2 * h + pad[0], 2 * w_ + pad[1]
    up = F.conv_transpose2d(x, wq, bq, stride=2)
    ref = F.pad(up, [pad[1] // 2, pad[1] - pad[1] // 2, pad[0] // 2, pad[0] - pad[0] // 2])
    gy = rq(torch.randn(ref.shape, generator=g), dt_)
    ref.backward(gy)
    off = (pad[0] // 2, pad[1] // 2)
    tol = 2e-5 if dt_ == FP and algo == _lib.ALGO_SIMT else (2e-3 if dt_ == FP else 1.2e-2)
    if dt_ == FP and ops.x3_mode() and algo != _lib.ALGO_SIMT:
        tol = 5e-5                     # 3xTF32: fp32-level results from the tensor cores
    tag = f"{Ci}to{Co}_{h}x{w_}_pad{pad[0]}{pad[1]}_{str(dt_)[6:]}_{'simt' if algo == _lib.ALGO_SIMT else 'tc'}"
    xd = dev_nhwc(x.detach(), dt_)
    # destination = second half of a concat buffer [skip(Co) | up(Co)]
    cat = ops.empty_nhwc(B, 2 * Co, H, W, dt_, DEV)
    cat.fill_(0.0)
    upv = ops.channel_slice(cat, Co, Co)
    d = ops.make_gconv(ops._DT[dt_], algo, B, h, w_, Ci, ops.TAPS1, 1, (0, 0), h, w_, ops.nhwc_ld(xd), 4 * Co, 4, 2,
                       off, H, W, ops.nhwc_ld(cat))
    d.algo = algo
    used = ops.gconv_fprop(d, xd, UF.packT_fprop(w.to(DEV), dt_), b.to(DEV), upv, None)
    assert used == algo or algo == _lib.ALGO_AUTO
    res.append((f"convT_fprop_{tag}", rel(host(upv), ref.detach()), tol))
    res.append((f"convT_fprop_{tag}_skip_untouched", host(ops.channel_slice(cat, 0, Co)).abs().max().item(), 0.0))
    gcat = ops.empty_nhwc(B, 2 * Co, H, W, dt_, DEV)
    gup = ops.channel_slice(gcat, Co, Co)
    gup.copy_(gy.to(DEV).to(dt_))
    gx = ops.empty_nhwc(B, Ci, h, w_, dt_, DEV)
    dd = ops.make_gconv(ops._DT[dt_], algo, B, h, w_, Co, ops.TAPS_Q, 2, off, H, W, ops.nhwc_ld(gcat), Ci, 1, 1,
                        (0, 0), h, w_, ops.nhwc_ld(gx))
    dd.algo = algo
    ops.gconv_fprop(dd, gup, UF.packT_dgrad(w.to(DEV), dt_), None, gx, None)
    res.append((f"convT_dgrad_{tag}", rel(host(gx), x.grad), tol))
    dW = torch.empty(Ci, Co, 2, 2, device=DEV)
    d2 = ops.make_gconv(ops._DT[dt_], algo, B, h, w_, Ci, ops.TAPS1, 1, (0, 0), h, w_, ops.nhwc_ld(xd), 4 * Co, 4, 2,
                        off, H, W, ops.nhwc_ld(gcat))
    d2.algo = algo
    ops.gconv_wgrad(d2, xd, gup, dW, 0, Co * 4, 4, sq=1)
    res.append((f"convT_wgrad_{tag}", rel(host(dW), wq.grad), 2e-4 if tol < 1e-3 else 5e-3))
    return res


def check_conv_simt():
    out = []
    S = _lib.ALGO_SIMT
    out += _conv3x3_case(2, 1, 8, 9, 11, FP, S, 10)
    out += _conv3x3_case(2, 1, 64, 16, 20, FP, S, 50)      # first-layer kernels (C_in = 1, 64 channels)
    out += _conv3x3_case(2, 1, 64, 16, 20, BF, S, 51)
    out += _conv3x3_case(1, 3, 64, 12, 12, BF, S, 52)      # RGB first layer (fprop special, wgrad generic)
    out += _conv3x3_case(1, 4, 32, 7, 9, FP, S, 53)
    out += _conv3x3_case(1, 3, 20, 8, 8, FP, S, 11)
    out += _conv3x3_case(2, 16, 24, 10, 12, FP, S, 12)
    out += _conv3x3_case(2, 32, 64, 7, 9, BF, S, 13)
    out += _conv3x3_case(1, 6, 10, 5, 6, BF, S, 14)
    out += _conv3x3_case(2, 64, 64, 8, 8, FP, S, 15, slice_io=True)
    out += _convT_case(2, 16, 8, 5, 6, (0, 0), FP, S, 16)
    out += _convT_case(1, 8, 4, 4, 3, (1, 1), FP, S, 17)
    out += _convT_case(2, 32, 16, 4, 5, (1, 2), BF, S, 18)
    out += _convT_case(1, 6, 3, 3, 3, (0, 0), FP, S, 19)
    # UNet-bottleneck-like shapes: many channels, 2x2 / 4x4 pixels
    out += _conv3x3_case(2, 256, 512, 2, 2, FP, S, 40)
    out += _conv3x3_case(2, 512, 256, 4, 4, FP, S, 41)
    out += _convT_case(2, 256, 128, 2, 2, (0, 0), FP, S, 42)
    out += _convT_case(2, 128, 64, 4, 4, (0, 0), FP, S, 43)
    return out


def _narrow_case(B, Ci, Co, H, W, seed, slice_in=False):
    """3x3 conv with narrow channel counts (UNet_S / UNet_T / UNet_SA layers) in bf16: the thread-built-im2col tcgen05
    kernel -- fprop + BatchNorm statistics, dgrad (the same kernel on rotated / transposed weights) and the folded
    eval-mode BatchNorm + ReLU epilogue -- against PyTorch-CPU fp32 on the rounded operands."""
    res = []
    g = gen(seed)
    dt_ = BF
    x = rq(torch.randn(B, Ci, H, W, generator=g), dt_).requires_grad_(True)
    w = torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)
    wq = rq(w, dt_).requires_grad_(True)
    ref = F.conv2d(x, wq, padding=1)
    gy = rq(torch.randn(ref.shape, generator=g), dt_)
    ref.backward(gy)
    tag = f"narrow_{Ci}to{Co}_{B}x{H}x{W}{'_slice' if slice_in else ''}"
    xd = in_slice(x.detach(), dt_, 16) if slice_in else dev_nhwc(x.detach(), dt_)
    wdev = w.to(DEV)
    y = ops.empty_nhwc(B, Co, H, W, dt_, DEV)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)

    def desc(xin, Cout, yout):
        return ops.make_gconv(ops._DT[dt_], _lib.ALGO_AUTO, B, H, W, xin.shape[1], ops.TAPS3, 1, (0, 0), H, W,
                              ops.nhwc_ld(xin), Cout, 1, 1, (0, 0), H, W, ops.nhwc_ld(yout))
    used = ops.gconv_fprop(desc(xd, Co, y), xd, UF.pack3x3_fprop(wdev, dt_), None, y, stats)
    # C_in = 1 -> 8 / 16 / 32 channels is the CUDA-core first-layer kernel (conv_first_narrow.cu: K = 9, HBM-bound)
    want = _lib.ALGO_SIMT if (Ci == 1 and Co in (8, 16, 32)) else _lib.ALGO_TC
    res.append((f"{tag}_on_tensor_cores", 0.0 if used == want else 1.0, 0.0))
    yh = host(y)
    res.append((f"{tag}_fprop", rel(yh, ref.detach()), 1.2e-2))
    st = host(stats.float()).reshape(2, Co)
    res.append((f"{tag}_stats_sum", rel(st[0], yh.sum((0, 2, 3))), 1e-4))
    res.append((f"{tag}_stats_sq", rel(st[1], (yh ** 2).sum((0, 2, 3))), 1e-4))
    if Ci >= 8:
        gyd = dev_nhwc(gy, dt_)
        gx = ops.empty_nhwc(B, Ci, H, W, dt_, DEV)
        used = ops.gconv_fprop(desc(gyd, Ci, gx), gyd, UF.pack3x3_dgrad(wdev, dt_), None, gx, None)
        res.append((f"{tag}_dgrad_on_tensor_cores", 0.0 if used == _lib.ALGO_TC else 1.0, 0.0))
        res.append((f"{tag}_dgrad", rel(host(gx), x.grad), 1.2e-2))
    gyd = dev_nhwc(gy, dt_)
    dW = torch.empty(Co, Ci, 3, 3, device=DEV)
    used = ops.gconv_wgrad(desc(xd, Co, gyd), xd, gyd, dW, 1, 9, Ci * 9)
    res.append((f"{tag}_wgrad_on_tensor_cores", 0.0 if used == want else 1.0, 0.0))
    res.append((f"{tag}_wgrad", rel(host(dW), wq.grad), 5e-3))
    # folded eval-mode BatchNorm + ReLU
    sc, sh = torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g) * 0.2
    coefs = torch.stack([torch.zeros(Co), torch.ones(Co), sc, sh]).to(DEV).contiguous()
    z = ops.empty_nhwc(B, Co, H, W, dt_, DEV)
    d = desc(xd, Co, z)
    ok = ops.gconv_fprop_affine_relu_supported(d, xd, UF.pack3x3_fprop(wdev, dt_), z)
    res.append((f"{tag}_fold_supported", 0.0 if ok else 1.0, 0.0))
    if ok:
        ops.gconv_fprop_affine_relu(d, xd, UF.pack3x3_fprop(wdev, dt_), coefs, z)
        zref = torch.relu(ref.detach() * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
        res.append((f"{tag}_fold", rel(host(z), zref), 1.2e-2))
    return res


def _wgrad_f32_narrow_case(B, Ci, Co, H, W, seed):
    """exact-fp32 weight gradient of a narrow 3x3 layer (conv_simt_narrow.cu) against PyTorch-CPU fp32"""
    g = gen(seed)
    x = torch.randn(B, Ci, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)).requires_grad_(True)
    ref = F.conv2d(x, w, padding=1)
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    xd, gyd = dev_nhwc(x.detach(), FP), dev_nhwc(gy, FP)
    dW = torch.empty(Co, Ci, 3, 3, device=DEV)
    d = ops.make_gconv(ops._DT[FP], _lib.ALGO_SIMT, B, H, W, Ci, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(xd), Co, 1, 1,
                       (0, 0), H, W, ops.nhwc_ld(gyd))
    ops.gconv_wgrad(d, xd, gyd, dW, 1, 9, Ci * 9)
    res = [(f"wgrad_f32_narrow_{Ci}to{Co}_{B}x{H}x{W}", rel(host(dW), w.grad), 2e-5)]
    # forward (+ BatchNorm statistics) and data gradient through the same exact-fp32 CUDA-core kernel
    y = ops.empty_nhwc(B, Co, H, W, FP, DEV)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)
    ops.gconv_fprop(d, xd, UF.pack3x3_fprop(w.detach().to(DEV), FP), None, y, stats)
    yh = host(y)
    res.append((f"fprop_f32_narrow_{Ci}to{Co}_{B}x{H}x{W}", rel(yh, ref.detach()), 2e-5))
    st = host(stats.float()).reshape(2, Co)
    res.append((f"fprop_f32_narrow_{Ci}to{Co}_{B}x{H}x{W}_stats_sum", rel(st[0], yh.sum((0, 2, 3))), 1e-4))
    res.append((f"fprop_f32_narrow_{Ci}to{Co}_{B}x{H}x{W}_stats_sq", rel(st[1], (yh ** 2).sum((0, 2, 3))), 1e-4))
    gx = ops.empty_nhwc(B, Ci, H, W, FP, DEV)
    dd = ops.make_gconv(ops._DT[FP], _lib.ALGO_SIMT, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gyd), Ci, 1, 1,
                        (0, 0), H, W, ops.nhwc_ld(gx))
    ops.gconv_fprop(dd, gyd, UF.pack3x3_dgrad(w.detach().to(DEV), FP), None, gx, None)
    res.append((f"dgrad_f32_narrow_{Ci}to{Co}_{B}x{H}x{W}", rel(host(gx), x.grad), 2e-5))
    return res


def _narrow_tf32_case(B, Ci, Co, H, W, seed):
    """fp32 tensors through the TF32 form of the TMA-staged narrow kernel (fprop + statistics, dgrad, folded eval
    BatchNorm + ReLU) against PyTorch-CPU fp32; the weight gradient of these layers is the exact CUDA-core kernel."""
    res = []
    g = gen(seed)
    x = torch.randn(B, Ci, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)).requires_grad_(True)
    ref = F.conv2d(x, w, padding=1)
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    tag = f"narrow_tf32_{Ci}to{Co}_{B}x{H}x{W}"
    xd, gyd, wdev = dev_nhwc(x.detach(), FP), dev_nhwc(gy, FP), w.detach().to(DEV)
    y = ops.empty_nhwc(B, Co, H, W, FP, DEV)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)

    def desc(xin, Cout, yout):
        return ops.make_gconv(ops._DT[FP], _lib.ALGO_PREFER_TC, B, H, W, xin.shape[1], ops.TAPS3, 1, (0, 0), H, W,
                              ops.nhwc_ld(xin), Cout, 1, 1, (0, 0), H, W, ops.nhwc_ld(yout))
    used = ops.gconv_fprop(desc(xd, Co, y), xd, UF.pack3x3_fprop(wdev, FP), None, y, stats)
    res.append((f"{tag}_on_tensor_cores", 0.0 if used == _lib.ALGO_TC else 1.0, 0.0))
    yh = host(y)
    res.append((f"{tag}_fprop", rel(yh, ref.detach()), 2e-3))
    st = host(stats.float()).reshape(2, Co)
    res.append((f"{tag}_stats_sum", rel(st[0], yh.sum((0, 2, 3))), 1e-4))
    res.append((f"{tag}_stats_sq", rel(st[1], (yh ** 2).sum((0, 2, 3))), 1e-4))
    gx = ops.empty_nhwc(B, Ci, H, W, FP, DEV)
    used = ops.gconv_fprop(desc(gyd, Ci, gx), gyd, UF.pack3x3_dgrad(wdev, FP), None, gx, None)
    res.append((f"{tag}_dgrad_on_tensor_cores", 0.0 if used == _lib.ALGO_TC else 1.0, 0.0))
    res.append((f"{tag}_dgrad", rel(host(gx), x.grad), 2e-3))
    dW = torch.empty(Co, Ci, 3, 3, device=DEV)
    ops.gconv_wgrad(desc(xd, Co, gyd), xd, gyd, dW, 1, 9, Ci * 9)
    res.append((f"{tag}_wgrad_exact", rel(host(dW), w.grad), 2e-5 if min(Ci, Co) >= 16 else 2e-4))
    sc, sh = torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g) * 0.2
    coefs = torch.stack([torch.zeros(Co), torch.ones(Co), sc, sh]).to(DEV).contiguous()
    z = ops.empty_nhwc(B, Co, H, W, FP, DEV)
    d = desc(xd, Co, z)
    ok = ops.gconv_fprop_affine_relu_supported(d, xd, UF.pack3x3_fprop(wdev, FP), z)
    res.append((f"{tag}_fold_supported", 0.0 if ok else 1.0, 0.0))
    if ok:
        ops.gconv_fprop_affine_relu(d, xd, UF.pack3x3_fprop(wdev, FP), coefs, z)
        zref = torch.relu(ref.detach() * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
        res.append((f"{tag}_fold", rel(host(z), zref), 2e-3))
    return res


def check_conv_narrow():
    out = []
    out += _narrow_case(2, 16, 16, 20, 24, 101)
    out += _narrow_case(1, 16, 32, 17, 9, 102)                     # ragged: the last tile is partial
    out += _narrow_case(2, 32, 32, 16, 16, 103, slice_in=True)     # the input is a slice of a concat buffer
    out += _narrow_case(1, 32, 64, 12, 20, 104)
    out += _narrow_case(1, 64, 32, 24, 16, 105)
    out += _narrow_case(2, 32, 16, 33, 7, 106)
    out += _narrow_case(2, 8, 8, 16, 24, 107)                      # UNet_T: N = 8 rides an N = 16 MMA
    out += _narrow_case(1, 16, 8, 10, 10, 108)
    out += _narrow_case(2, 1, 16, 20, 20, 109)                     # first layers of the light variants
    out += _narrow_case(3, 1, 16, 200, 300, 118)                   # ... several tiles per block, ragged in both directions
    out += _narrow_case(2, 1, 8, 37, 530, 119)
    out += _narrow_case(2, 1, 32, 64, 100, 126)
    out += _narrow_case(1, 3, 16, 24, 24, 127)                     # RGB input stays on the im2col kernel
    out += _narrow_case(1, 3, 8, 18, 14, 110)
    out += _narrow_case(1, 8, 16, 40, 40, 111)
    # the TMA-staged kernels (conv_halo.cu): several tiles per CTA (ring wrap, both accumulator sets), every channel pair
    out += _narrow_case(8, 16, 16, 256, 256, 112)
    out += _narrow_case(4, 32, 32, 200, 136, 113)
    out += _narrow_case(2, 64, 16, 64, 48, 114)
    out += _narrow_case(2, 16, 64, 40, 56, 115)
    out += _narrow_case(3, 64, 32, 130, 100, 116)
    out += _narrow_case(2, 32, 16, 72, 88, 117, slice_in=True)
    # 8-channel tensors (UNet_T) ride the 16-channel instantiation: TMA zero-fills the upper half of the 32-byte rows
    out += _narrow_case(4, 8, 8, 200, 136, 131)
    out += _narrow_case(2, 16, 8, 130, 100, 132)
    out += _narrow_case(2, 8, 32, 64, 48, 133)
    # the same layers without autocast: exact-fp32 weight gradient on the CUDA cores, several tiles per block, ragged
    for i, (B, Ci, Co, H, W) in enumerate([(2, 16, 16, 40, 70), (8, 16, 16, 256, 256), (2, 16, 32, 33, 20), (1, 32, 16, 9, 50),
                                           (3, 32, 32, 64, 48), (1, 64, 32, 24, 17), (2, 32, 64, 16, 16), (1, 16, 64, 20, 36),
                                           (1, 64, 16, 10, 9), (2, 8, 8, 37, 41), (1, 8, 16, 24, 24), (2, 16, 8, 16, 48),
                                           (1, 8, 32, 20, 20), (1, 64, 8, 9, 9), (4, 8, 8, 128, 128)]):
        out += _wgrad_f32_narrow_case(B, Ci, Co, H, W, 140 + i)
    # ... and with TF32 allowed: fprop / dgrad on the tensor cores from fp32 tensors (64 / 128-byte swizzle rows)
    for i, (B, Ci, Co, H, W) in enumerate([(2, 16, 16, 40, 70), (8, 16, 16, 256, 256), (2, 16, 32, 33, 20), (1, 32, 16, 9, 50),
                                           (4, 32, 32, 200, 136), (2, 8, 8, 37, 41), (1, 8, 16, 24, 24), (2, 32, 8, 16, 48)]):
        out += _narrow_tf32_case(B, Ci, Co, H, W, 160 + i)
    # narrow ConvTranspose2d (conv_halo_t.cu): partial tiles in both directions, rows past the image (h % 8 != 0) that
    # the 5-D quadrant view reads from the next image, several tiles per CTA, padded destination (fprop only)
    out += _convT_case(2, 32, 16, 24, 40, (0, 0), BF, _lib.ALGO_TC, 120)
    out += _convT_case(3, 64, 32, 20, 16, (0, 0), BF, _lib.ALGO_TC, 121)
    out += _convT_case(1, 32, 32, 9, 7, (0, 0), BF, _lib.ALGO_TC, 122)
    out += _convT_case(2, 64, 16, 8, 16, (0, 0), BF, _lib.ALGO_TC, 123)
    out += _convT_case(16, 32, 16, 128, 128, (0, 0), BF, _lib.ALGO_TC, 124)
    out += _convT_case(2, 32, 16, 5, 6, (1, 2), BF, _lib.ALGO_AUTO, 125)
    out += _convT_case(2, 16, 8, 24, 40, (0, 0), BF, _lib.ALGO_TC, 134)          # UNet_T's last up-sampler (padded instantiation)
    out += _convT_case(3, 16, 8, 9, 7, (0, 0), BF, _lib.ALGO_TC, 135)
    out += _convT_case(8, 16, 8, 128, 128, (0, 0), BF, _lib.ALGO_TC, 136)
    return out


def _guarded_nhwc(B, Cc, H, W, dtype, guard=8192, fill=3.0):
    """Logical [B, C, H, W] NHWC view in the middle of a flat buffer whose `guard` elements either side hold a
    sentinel: an out-of-bounds store of a kernel shows up as a changed guard (compute-sanitizer is not available on
    the GPU pool)."""
    n = B * H * W * Cc
    flat = torch.full((guard + n + guard,), fill, dtype=dtype, device=DEV)
    view = flat[guard:guard + n].view(B, H, W, Cc).permute(0, 3, 1, 2)
    return flat, view, guard


def _guards_intact(flat, guard, fill=3.0):
    return 0.0 if bool((flat[:guard] == fill).all()) and bool((flat[-guard:] == fill).all()) else 1.0


def check_narrow_bounds():
    """Stores of the TMA-staged narrow kernels stay inside their tensors on ragged shapes (partial tiles in both
    directions, rows of a tile past the image, the last tile of the last image)."""
    res = []
    g = gen(130)
    for (B, Ci, Co, H, W) in [(2, 16, 16, 20, 24), (1, 16, 32, 17, 9), (3, 64, 32, 35, 50), (2, 1, 8, 37, 530), (2, 1, 16, 9, 130),
                              (1, 32, 64, 5, 3), (2, 8, 8, 21, 19), (1, 16, 8, 33, 40)]:
        x = dev_nhwc(torch.randn(B, Ci, H, W, generator=g), BF)
        w = torch.randn(Co, Ci, 3, 3, generator=g).to(DEV) / (3 * Ci ** 0.5)
        flat, y, gd = _guarded_nhwc(B, Co, H, W, BF)
        stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)
        d = ops.make_gconv(ops._DT[BF], _lib.ALGO_AUTO, B, H, W, Ci, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(x), Co, 1, 1,
                           (0, 0), H, W, ops.nhwc_ld(y))
        ops.gconv_fprop(d, x, UF.pack3x3_fprop(w, BF), None, y, stats)
        torch.cuda.synchronize()
        res.append((f"bounds_fprop_{Ci}to{Co}_{B}x{H}x{W}", _guards_intact(flat, gd), 0.0))
        ref = F.conv2d(host(x), rq(w.cpu(), BF), padding=1)
        res.append((f"bounds_fprop_{Ci}to{Co}_{B}x{H}x{W}_values", rel(host(y), ref), 1.2e-2))
        if Ci >= 8:
            gy = dev_nhwc(torch.randn(B, Co, H, W, generator=g), BF)
            flat2, gx, gd2 = _guarded_nhwc(B, Ci, H, W, BF)
            dd = ops.make_gconv(ops._DT[BF], _lib.ALGO_AUTO, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gy), Ci, 1, 1,
                                (0, 0), H, W, ops.nhwc_ld(gx))
            ops.gconv_fprop(dd, gy, UF.pack3x3_dgrad(w, BF), None, gx, None)
            torch.cuda.synchronize()
            res.append((f"bounds_dgrad_{Ci}to{Co}_{B}x{H}x{W}", _guards_intact(flat2, gd2), 0.0))
    for (B, Ci, Co, h, w_) in [(2, 32, 16, 9, 7), (1, 64, 32, 3, 20), (3, 32, 32, 8, 16), (2, 16, 8, 9, 7)]:
        x = dev_nhwc(torch.randn(B, Ci, h, w_, generator=g), BF)
        wT = torch.randn(Ci, Co, 2, 2, generator=g).to(DEV) / (Ci ** 0.5)
        b = torch.randn(Co, generator=g).to(DEV)
        H, W = 2 * h, 2 * w_
        flat, up, gd = _guarded_nhwc(B, Co, H, W, BF)
        d = ops.make_gconv(ops._DT[BF], _lib.ALGO_AUTO, B, h, w_, Ci, ops.TAPS1, 1, (0, 0), h, w_, ops.nhwc_ld(x), 4 * Co, 4, 2,
                           (0, 0), H, W, ops.nhwc_ld(up))
        ops.gconv_fprop(d, x, UF.packT_fprop(wT, BF), b, up, None)
        flat2, gx, gd2 = _guarded_nhwc(B, Ci, h, w_, BF)
        gup = dev_nhwc(torch.randn(B, Co, H, W, generator=g), BF)
        dd = ops.make_gconv(ops._DT[BF], _lib.ALGO_AUTO, B, h, w_, Co, ops.TAPS_Q, 2, (0, 0), H, W, ops.nhwc_ld(gup), Ci, 1, 1,
                            (0, 0), h, w_, ops.nhwc_ld(gx))
        ops.gconv_fprop(dd, gup, UF.packT_dgrad(wT, BF), None, gx, None)
        torch.cuda.synchronize()
        res.append((f"bounds_convT_fprop_{Ci}to{Co}_{B}x{h}x{w_}", _guards_intact(flat, gd), 0.0))
        res.append((f"bounds_convT_dgrad_{Ci}to{Co}_{B}x{h}x{w_}", _guards_intact(flat2, gd2), 0.0))
    return res


def check_conv_tc_fprop_small():
    """First contact with the tcgen05 engine: one tile, one N block."""
    return _conv3x3_case(1, 64, 64, 8, 16, BF, _lib.ALGO_TC, 20)


def check_conv_tc():
    out = []
    T = _lib.ALGO_TC
    out += _conv3x3_case(2, 64, 64, 16, 24, BF, T, 21)
    out += _conv3x3_case(1, 128, 128, 20, 12, BF, T, 22)
    out += _conv3x3_case(2, 128, 256, 9, 7, BF, T, 23)
    out += _conv3x3_case(1, 256, 512, 6, 6, BF, T, 24)
    out += _conv3x3_case(2, 64, 128, 2, 2, BF, T, 25)
    out += _conv3x3_case(2, 64, 64, 16, 16, BF, T, 26, slice_io=True)
    out += _conv3x3_case(2, 64, 64, 40, 36, BF, T, 60)          # several 16x16 tiles, ragged edges
    out += _conv3x3_case(1, 128, 256, 33, 17, BF, T, 61)
    out += _conv3x3_case(3, 64, 128, 20, 8, BF, T, 62)          # W <= 8: stacked sub-tile geometry
    out += _conv3x3_case(1, 192, 64, 48, 48, BF, T, 63)         # 3 channel chunks, odd unit count
    out += _convT_case(2, 128, 64, 6, 10, (0, 0), BF, T, 27)
    out += _convT_case(1, 256, 128, 5, 4, (1, 1), BF, T, 28)
    out += _convT_case(1, 512, 256, 4, 4, (0, 0), BF, T, 29)
    return out


def _conv3x3_cl_case(B, Ci, Co, H, W, dt_, seed):
    """3x3 conv with a channels_last parameter (what `model.to(memory_format=channels_last)` gives train.py:262):
    the dgrad operand is packed by the transposing kernel and the weight gradient is written in the parameter's
    own layout by the coalesced split reduction (few / >= 8 / >= 32 splits depending on the shape)."""
    res = []
    g = gen(seed)
    x = rq(torch.randn(B, Ci, H, W, generator=g), dt_).requires_grad_(True)
    w = torch.randn(Co, Ci, 3, 3, generator=g) / ((Ci * 9) ** 0.5)
    wq = rq(w, dt_).requires_grad_(True)
    ref = F.conv2d(x, wq, padding=1)
    gy = rq(torch.randn(ref.shape, generator=g), dt_)
    ref.backward(gy)
    tol = 2e-3 if dt_ == FP else 1.2e-2
    tag = f"cl_{Ci}to{Co}_{B}x{H}x{W}_{str(dt_)[6:]}"
    algo = _lib.ALGO_TC
    xd, gyd = dev_nhwc(x.detach(), dt_), dev_nhwc(gy, dt_)
    wcl = w.to(DEV).contiguous(memory_format=torch.channels_last)
    assert wcl.stride(1) == 1

    def desc(xin, Cout, yout):
        d = ops.make_gconv(ops._DT[dt_], algo, B, H, W, xin.shape[1], ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(xin),
                           Cout, 1, 1, (0, 0), H, W, ops.nhwc_ld(yout))
        d.algo = algo
        return d
    y = ops.empty_nhwc(B, Co, H, W, dt_, DEV)
    ops.gconv_fprop(desc(xd, Co, y), xd, UF.pack3x3_fprop(wcl, dt_), None, y, None)
    res.append((f"fprop_{tag}", rel(host(y), ref.detach()), tol))
    gx = ops.empty_nhwc(B, Ci, H, W, dt_, DEV)
    ops.gconv_fprop(desc(gyd, Ci, gx), gyd, UF.pack3x3_dgrad(wcl, dt_), None, gx, None)
    res.append((f"dgrad_{tag}", rel(host(gx), x.grad), tol))
    dW = torch.empty_like(wcl)
    so, si, skh, skw = dW.stride()
    assert skh == 3 * skw and si == 1
    ops.gconv_wgrad(desc(xd, Co, gyd), xd, gyd, dW, skw, si, so)
    res.append((f"wgrad_{tag}", rel(host(dW), wq.grad), 5e-3))
    return res


def check_conv_layouts():
    out = []
    out += _conv3x3_cl_case(2, 64, 64, 64, 64, BF, 70)        # 2 wgrad tiles -> >= 32 splits
    out += _conv3x3_cl_case(2, 128, 128, 32, 48, BF, 71)      # 8 <= splits < 32
    out += _conv3x3_cl_case(1, 256, 512, 16, 16, BF, 72)      # CTA-pair wgrad, few splits
    out += _conv3x3_cl_case(3, 512, 256, 16, 8, BF, 73)       # CTA-pair wgrad, W <= 8, odd batch (a pair with one tile)
    out += _conv3x3_cl_case(1, 64, 128, 24, 40, FP, 74)       # tf32
    return out


def check_conv_tc_tf32():
    out = []
    T = _lib.ALGO_TC
    out += _conv3x3_case(2, 64, 64, 10, 12, FP, T, 31)
    out += _conv3x3_case(1, 64, 128, 16, 16, FP, T, 32)
    out += _conv3x3_case(1, 128, 64, 7, 9, FP, T, 33, slice_io=True)
    out += _convT_case(2, 64, 64, 6, 5, (0, 0), FP, T, 34)
    out += _convT_case(1, 128, 64, 4, 4, (1, 0), FP, T, 35)
    return out


# ------------------------------------------------------------------------------------------------
# parts (modules) against the golden part fixtures, and the full network against the oracle
# ------------------------------------------------------------------------------------------------
def _part_case(golden, tag, build, mode):
    """mode: 'fp32' (exact SIMT), 'bf16' (autocast)."""
    c = golden[tag]
    mod = build().to(DEV)
    mod.load_state_dict({k: v.clone() for k, v in c["state"].items()})
    mod.train()
    ins = [t.to(DEV).requires_grad_(True) for t in c["inputs"]]
    tol = 2e-4 if mode == "fp32" else 4e-2
    os.environ["UNET_B200_PRECISION"] = "fp32"
    with torch.autocast("cuda", enabled=(mode == "bf16")):
        out = mod(*ins)
    out.backward(c["gout"].to(DEV).to(out.dtype))
    res = [(f"{tag}_{mode}_out", rel(host(out), c["out"]), tol)]
    # bf16 on a few hundred pixels: ReLU-mask flips dominate the gradients (see tests/gpu_e2e.py), so the
    # backward is only sanity-checked in relative L2 there; fp32 is held to max-rel 2e-4.
    err = rel if mode == "fp32" else O.rel_l2
    gtol = tol if mode == "fp32" else 0.6
    for i, (a, b) in enumerate(zip(ins, c["gin"])):
        res.append((f"{tag}_{mode}_gin{i}", err(host(a.grad), b), gtol))
    for k, p in mod.named_parameters():
        res.append((f"{tag}_{mode}_g_{k}", err(host(p.grad), c["gparams"][k]), gtol))
    for k, v in mod.state_dict().items():
        if "running" in k:
            res.append((f"{tag}_{mode}_{k}", rel(host(v), _updated_running(c, k)), tol))
    return res


def _updated_running(c, key):
    """The fixture stores the state *after* the reference's forward (state_dict is taken post-step)."""
    return c["state"][key]


def check_parts(golden, mode="fp32"):
    from unet import unet_parts as P
    out = []
    # the fixture state was saved after the forward pass, so running stats in it are already updated;
    # parity of those is covered by the full-network checks.  Reset them here.
    def fresh(build):
        def f():
            return build()
        return f
    specs = [("DoubleConv_4_8", lambda: P.DoubleConv(4, 8)), ("DoubleConv_4_8_mid6", lambda: P.DoubleConv(4, 8, 6)),
             ("Down_4_8", lambda: P.Down(4, 8)), ("Down_4_8_odd", lambda: P.Down(4, 8)),
             ("Up_8_4_convT", lambda: P.Up(8, 4, bilinear=False)), ("Up_8_4_convT_pad", lambda: P.Up(8, 4, bilinear=False)),
             ("Up_8_4_bilinear", lambda: P.Up(8, 4, bilinear=True)), ("Up_8_4_bilinear_pad", lambda: P.Up(8, 4, bilinear=True)),
             ("OutConv_8_3", lambda: P.OutConv(8, 3))]
    for tag, build in specs:
        res = _part_case(golden, tag, fresh(build), mode)
        out += [r for r in res if "running" not in r[0]]
    return out


def unet_step_gpu(model, img, msk, amp, boundary_coeff=0.0, fused=True):
    """One training step on the GPU through the drop-in modules; returns logits, loss, grads (CPU)."""
    model.zero_grad(set_to_none=True)
    x = img.to(DEV).contiguous(memory_format=torch.channels_last)
    t = msk.to(DEV)
    with torch.autocast("cuda", enabled=amp):
        logits = model(x)
        if fused:
            loss = UL.training_criterion(logits, t, boundary_coeff=boundary_coeff)
        else:      # exactly the reference's train.py:137-142 composition, dice through the drop-in
            from utils.dice_score import dice_loss
            loss = F.cross_entropy(logits, t)
            loss = loss + dice_loss(F.softmax(logits, dim=1).float(),
                                    F.one_hot(t, model.n_classes).permute(0, 3, 1, 2).float(), multiclass=True)
    loss.backward()
    grads = {k: host(p.grad) for k, p in model.named_parameters()}
    return host(logits), float(loss), grads


def check_unet(nc, ncls, bilinear, B, H, W, mode, fused=True, boundary_coeff=0.0):
    """mode: 'fp32' | 'tf32' | 'bf16'."""
    import unet
    tag = f"unet{nc}_{ncls}_{'bil' if bilinear else 'convT'}_{B}x{H}x{W}_{mode}{'' if fused else '_unfused'}"
    st = O.build_state(nc, ncls, bilinear, seed=0)
    img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    ref_st = {k: v.clone() for k, v in st.items()}
    r_logits, r_loss, r_grads = O.training_step(ref_st, img, msk, ncls, bilinear, boundary_coeff=boundary_coeff)
    model = unet.UNet(nc, ncls, bilinear)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    os.environ["UNET_B200_PRECISION"] = "tf32" if mode == "tf32" else "fp32"
    logits, loss, grads = unet_step_gpu(model, img, msk, amp=(mode == "bf16"), boundary_coeff=boundary_coeff, fused=fused)
    tol = {"fp32": 1e-3, "tf32": 1e-3, "bf16": 2e-2}[mode]
    res = [(f"{tag}_logits", rel(logits, r_logits), tol),
           (f"{tag}_loss", abs(loss - float(r_loss)) / abs(float(r_loss)), tol),
           (f"{tag}_argmax_mismatch", (logits.argmax(1) != r_logits.argmax(1)).float().mean().item(),
            1e-3 if mode != "bf16" else 1e-2)]
    worst, worst_k = 0.0, ""
    l2 = 0.0
    table = []
    for k, gr in r_grads.items():
        e = rel(grads[k], gr)
        e2 = O.rel_l2(grads[k], gr)
        table.append((e, e2, k))
        l2 = max(l2, e2)
        if e > worst:
            worst, worst_k = e, k
    if os.environ.get("UNETB200_TEST_VERBOSE"):
        for e, e2, k in table:
            print(f"      grad {k:<50s} max-rel {e:.3e}  l2-rel {e2:.3e}")
    res.append((f"{tag}_grad_worst[{worst_k}]", worst, tol * (1 if mode == "fp32" else 5)))
    res.append((f"{tag}_grad_worst_l2", l2, tol * (1 if mode == "fp32" else 5)))
    sd = model.state_dict()
    rw = max(rel(host(sd[k]), ref_st[k]) for k in sd if "running" in k)
    res.append((f"{tag}_running_stats", rw, tol))
    nbt = all(int(sd[k]) == int(ref_st[k]) for k in sd if "tracked" in k)
    res.append((f"{tag}_num_batches_tracked", 0.0 if nbt else 1.0, 0.0))
    return res


def torch_gpu_step(st, img, msk, ncls, bilinear, mode):
    """The reference's own GPU path (PyTorch ATen/cuDNN through the oracle's functional restatement) --
    used ONLY to calibrate how far a reduced-precision GPU run sits from the CPU fp32 oracle."""
    names = O.param_names(st)
    leaves = {k: st[k].detach().clone().to(DEV).requires_grad_(True) for k in names}
    work = {k: (leaves[k] if k in leaves else v.clone().to(DEV)) for k, v in st.items()}
    x = img.to(DEV).contiguous(memory_format=torch.channels_last)
    t = msk.to(DEV)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = (mode == "tf32")
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            logits = O.unet_forward(work, x, bilinear, True)
            loss = O.train_loss(logits, t, ncls)
        grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return host(logits), float(loss), {k: host(g) for k, g in zip(names, grads)}


def calibrate(nc, ncls, bilinear, B, H, W, mode):
    """Errors vs the CPU fp32 oracle of (a) torch's GPU path and (b) ours, and (c) ours vs torch-GPU."""
    import statistics
    import unet
    tag = f"calib{nc}_{ncls}_{'bil' if bilinear else 'convT'}_{B}x{H}x{W}_{mode}"
    st = O.build_state(nc, ncls, bilinear, seed=0)
    img, msk = O.synthetic_batch(B, nc, ncls, H, W)
    r_logits, r_loss, r_grads = O.training_step({k: v.clone() for k, v in st.items()}, img, msk, ncls, bilinear)
    t_logits, t_loss, t_grads = torch_gpu_step(st, img, msk, ncls, bilinear, mode)
    model = unet.UNet(nc, ncls, bilinear)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    os.environ["UNET_B200_PRECISION"] = "tf32" if mode == "tf32" else "fp32"
    m_logits, m_loss, m_grads = unet_step_gpu(model, img, msk, amp=(mode == "bf16"))
    inf = float("inf")
    res = []
    for who, lg, ls, gr in (("torch_gpu", t_logits, t_loss, t_grads), ("ours", m_logits, m_loss, m_grads)):
        l2 = [O.rel_l2(gr[k], r_grads[k]) for k in r_grads]
        mx = [rel(gr[k], r_grads[k]) for k in r_grads]
        res += [(f"{tag}_{who}_logits_maxrel", rel(lg, r_logits), inf),
                (f"{tag}_{who}_loss_rel", abs(ls - float(r_loss)) / abs(float(r_loss)), inf),
                (f"{tag}_{who}_argmax_mismatch", (lg.argmax(1) != r_logits.argmax(1)).float().mean().item(), inf),
                (f"{tag}_{who}_grad_l2_median", statistics.median(l2), inf),
                (f"{tag}_{who}_grad_l2_worst", max(l2), inf),
                (f"{tag}_{who}_grad_maxrel_median", statistics.median(mx), inf),
                (f"{tag}_{who}_grad_maxrel_worst", max(mx), inf)]
    l2 = [O.rel_l2(m_grads[k], t_grads[k]) for k in r_grads]
    res += [(f"{tag}_ours_vs_torch_gpu_logits_maxrel", rel(m_logits, t_logits), inf),
            (f"{tag}_ours_vs_torch_gpu_grad_l2_median", statistics.median(l2), inf),
            (f"{tag}_ours_vs_torch_gpu_grad_l2_worst", max(l2), inf)]
    return res


def check_optim():
    """FusedRMSprop (+ folded clip_grad_norm_) against torch.optim.RMSprop + torch.nn.utils.clip_grad_norm_."""
    from unetb200.optim import FusedRMSprop
    out = []
    g = gen(90)
    shapes = [(64, 1, 3, 3), (128, 64, 3, 3), (64,), (7,), (2, 64, 1, 1), (256, 128, 2, 2), (3, 5, 3, 3), (1024, 17)]
    for mom, wd, clip in ((0.999, 1e-8, 1.0), (0.0, 0.0, None), (0.9, 1e-2, 0.05)):
        ps_a, ps_b = [], []
        for i, sh in enumerate(shapes):
            w = torch.randn(sh, generator=g)
            if len(sh) == 4 and i % 2 == 1:
                w = w.contiguous(memory_format=torch.channels_last)
            ps_a.append(torch.nn.Parameter(w.clone().to(DEV)))
            ps_b.append(torch.nn.Parameter(w.clone().to(DEV)))
            if len(sh) == 4 and i % 2 == 1:
                ps_a[-1].data = ps_a[-1].data.contiguous(memory_format=torch.channels_last)
                ps_b[-1].data = ps_b[-1].data.contiguous(memory_format=torch.channels_last)
        oa = torch.optim.RMSprop(ps_a, lr=1e-3, alpha=0.99, eps=1e-8, weight_decay=wd, momentum=mom, foreach=True)
        ob = FusedRMSprop(ps_b, lr=1e-3, alpha=0.99, eps=1e-8, weight_decay=wd, momentum=mom)
        norms = []
        for it in range(3):
            for pa, pb in zip(ps_a, ps_b):
                gr = torch.randn(pa.shape, generator=g).to(DEV) * (10.0 if it == 1 else 0.1)
                pa.grad = torch.empty_like(pa).copy_(gr)
                pb.grad = torch.empty_like(pb).copy_(gr)
            if clip is not None:
                na = torch.nn.utils.clip_grad_norm_(ps_a, clip)
                oa.step()
                nb = ob.step(clip_max_norm=clip)
                norms.append(abs(float(na) - float(nb)) / float(na))
            else:
                oa.step()
                ob.step()
        tag = f"mom{mom}_wd{wd}_clip{clip}"
        out.append((f"rmsprop_w_{tag}", max(rel(host(b), host(a)) for a, b in zip(ps_a, ps_b)), 1e-5))
        out.append((f"rmsprop_sq_{tag}", max(rel(host(ob.state[b]["square_avg"]), host(oa.state[a]["square_avg"]))
                                             for a, b in zip(ps_a, ps_b)), 1e-5))
        if mom > 0:
            out.append((f"rmsprop_buf_{tag}", max(rel(host(ob.state[b]["momentum_buffer"]), host(oa.state[a]["momentum_buffer"]))
                                                  for a, b in zip(ps_a, ps_b)), 1e-5))
        if norms:
            out.append((f"clip_total_norm_{tag}", max(norms), 1e-6))
    return out


def check_conv_bnfold():
    """Inference form of conv3x3 -> BatchNorm(running statistics) -> ReLU folded into the tcgen05 epilogue
    (unetb200_gconv_fprop_affine_relu) against PyTorch-CPU fp32 on the same (dtype-rounded) inputs, and against the
    unfused two-kernel path.  Covers both N blocks (64 / 128), the 8x32 tile variant (W <= 8), ragged tile edges,
    a destination that is a channel slice of a wider concat buffer, the fused max-pool follow-up, bf16 and tf32."""
    out = []
    g = gen(41)
    os.environ["UNET_B200_PRECISION"] = "tf32"
    try:
        for (B, Ci, Co, H, W, dt_, sliced, pool) in ((2, 64, 64, 40, 36, BF, False, False), (1, 128, 128, 32, 48, BF, True, True),
                                                    (2, 64, 128, 16, 8, BF, False, True), (1, 256, 64, 24, 24, BF, True, False),
                                                    (1, 64, 128, 20, 28, FP, False, True), (3, 128, 256, 8, 8, BF, False, False),
                                                    (2, 64, 64, 18, 22, BF, False, True), (1, 64, 128, 17, 21, BF, False, True),
                                                    # narrow layers: the pool rides in the TMA-staged kernel's epilogue too
                                                    (2, 16, 16, 40, 36, BF, False, True), (1, 32, 32, 18, 22, BF, False, True),
                                                    (1, 16, 8, 17, 21, BF, False, True), (2, 32, 64, 16, 24, BF, True, True)):
            x = rq(torch.randn(B, Ci, H, W, generator=g), dt_ if dt_ == BF else FP)
            w = torch.randn(Co, Ci, 3, 3, generator=g) * (2.0 / (9 * Ci)) ** 0.5
            bn = torch.nn.BatchNorm2d(Co)
            with torch.no_grad():
                bn.weight.copy_(0.5 + torch.rand(Co, generator=g))
                bn.bias.copy_(0.3 * torch.randn(Co, generator=g))
                bn.running_mean.copy_(0.2 * torch.randn(Co, generator=g))
                bn.running_var.copy_(0.5 + torch.rand(Co, generator=g))
            bn.eval()
            wq = rq(w, dt_) if dt_ == BF else w
            with torch.no_grad():
                ref = F.relu(bn(F.conv2d(x, wq, padding=1)))
                ref_pool = F.max_pool2d(ref, 2)
            bnd = torch.nn.BatchNorm2d(Co).to(DEV)
            bnd.load_state_dict(bn.state_dict())
            bnd.eval()
            xd = dev_nhwc(x, dt_)
            wd = w.to(DEV)
            res = {}
            for fold in (True, False):
                dst = None
                if sliced:
                    buf = ops.empty_nhwc(B, 2 * Co, H, W, dt_, DEV)
                    buf.fill_(7.0)
                    dst = ops.channel_slice(buf, 0, Co)
                with torch.no_grad():
                    y, z, pooled, _, _ = UF.conv_bn_relu_fwd(xd, wd, bnd, False, out=dst, want_pool=pool, fold=fold)
                if fold:
                    out.append((f"bnfold_taken_{Ci}_{Co}_{H}x{W}_{dt_}", 0.0 if y is None else 1.0, 0))
                    if sliced:
                        out.append((f"bnfold_slice_untouched_{Ci}_{Co}", float((host(buf[:, Co:]) != 7.0).sum().item()), 0))
                res[fold] = (host(z), host(pooled) if pool else None)
            tol = TOL[BF] if dt_ == BF else 2e-3          # fp32 storage runs on tf32 tensor cores here
            tag = f"{Ci}_{Co}_{H}x{W}_{dt_}{'_slice' if sliced else ''}"
            out.append((f"bnfold_vs_fp32_{tag}", rel(res[True][0], ref), tol))
            out.append((f"bnfold_vs_unfused_{tag}", rel(res[True][0], res[False][0]), tol))
            if pool:
                out.append((f"bnfold_pool_vs_fp32_{tag}", rel(res[True][1], ref_pool), tol))
                # the pooled tensor is exactly the max-pool of the z that was stored
                zs = res[True][0]
                out.append((f"bnfold_pool_exact_{tag}", float((F.max_pool2d(zs, 2) != res[True][1]).sum().item()), 0))
            out.append((f"bnfold_relu_{tag}", float((res[True][0] < 0).sum().item()), 0))
    finally:
        os.environ.pop("UNET_B200_PRECISION", None)
    return out


def check_conv_tc_x3():
    """UNET_B200_PRECISION=tf32x3: the same tcgen05 kind::tf32 kernels on hi/lo-split operands must give fp32-level
    results (5e-5 instead of plain TF32's 2e-3) for fprop, dgrad, wgrad of the 3x3 convs and of ConvTranspose."""
    old = os.environ.get("UNET_B200_PRECISION")
    os.environ["UNET_B200_PRECISION"] = "tf32x3"
    try:
        out = []
        T = _lib.ALGO_TC
        out += _conv3x3_case(2, 64, 64, 10, 12, FP, T, 31)
        out += _conv3x3_case(1, 64, 128, 16, 16, FP, T, 32)
        out += _conv3x3_case(1, 128, 64, 7, 9, FP, T, 33, slice_io=True)
        out += _conv3x3_case(1, 256, 256, 16, 24, FP, T, 36)
        out += _convT_case(2, 64, 64, 6, 5, (0, 0), FP, T, 34)
        out += _convT_case(1, 128, 64, 4, 4, (1, 0), FP, T, 35)
        return [(n + "_x3", e, t) for n, e, t in out]
    finally:
        if old is None:
            os.environ.pop("UNET_B200_PRECISION", None)
        else:
            os.environ["UNET_B200_PRECISION"] = old


def _dgrad_bnbwd_case(B, Ci, Co, H, W, dt_, seed, slice_y=False):
    """dgrad of conv3x3 (Co -> Ci channels) with the BatchNorm-backward reduction of the layer below fused into its
    epilogue: gx must equal the plain dgrad bit for bit, the sums must equal unetb200_bn_relu_bwd_reduce on that gx
    (fp32 partials, different order: 1e-5) and a CPU fp64 evaluation of the same formula on the stored gx."""
    res = []
    if dt_ == FP and ops.x3_mode():
        return res                      # the 3xTF32 split runs the plain kernels
    g = gen(seed)
    gy = rq(torch.randn(B, Co, H, W, generator=g), dt_)
    w = torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Co ** 0.5)
    yprev = rq(torch.randn(B, Ci, H, W, generator=g), dt_)
    gamma, beta = torch.rand(Ci, generator=g) + 0.5, torch.randn(Ci, generator=g) * 0.3
    tag = f"{Co}to{Ci}_{B}x{H}x{W}_{str(dt_)[6:]}{'_slice' if slice_y else ''}"
    gyd = dev_nhwc(gy, dt_)
    yd = in_slice(yprev, dt_, 64) if slice_y else dev_nhwc(yprev, dt_)
    stats = torch.stack([yprev.double().sum((0, 2, 3)), (yprev.double() ** 2).sum((0, 2, 3))]).reshape(-1).to(DEV)
    coefs = ops.bn_finalize(stats, B * H * W, gamma.to(DEV), beta.to(DEV), 1e-5, 0.0, None, None, Ci)
    wd = UF.pack3x3_dgrad(w.to(DEV), dt_)
    algo = _lib.ALGO_TC if dt_ == BF else _lib.ALGO_PREFER_TC
    gx0 = ops.empty_nhwc(B, Ci, H, W, dt_, DEV)
    gx1 = ops.empty_nhwc(B, Ci, H, W, dt_, DEV)
    d = ops.make_gconv(ops._DT[dt_], algo, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gyd),
                       Ci, 1, 1, (0, 0), H, W, ops.nhwc_ld(gx0))
    ops.gconv_fprop(d, gyd, wd, None, gx0, None, kind="dgrad")
    ok = ops.gconv_dgrad_bnbwd_supported(d, gyd, wd, gx1)
    res.append((f"bnbwd_supported_{tag}", 0.0 if ok else 1.0, 0.0))
    if not ok:
        return res
    sums = ops.gconv_dgrad_bnbwd(d, gyd, wd, gx1, yd, coefs)
    res.append((f"bnbwd_gx_bitequal_{tag}", (host(gx1) - host(gx0)).abs().max().item(), 0.0))
    ref = torch.zeros((2, Ci), dtype=torch.float64, device=DEV)
    L = ops.lib()
    _lib.check(L.unetb200_bn_relu_bwd_reduce(ops._p(gx0), ops.nhwc_ld(gx0), ops._p(yd), ops.nhwc_ld(yd), ops._p(coefs[2]),
                                             ops._p(coefs[3]), ops._p(coefs[0]), ops._p(coefs[1]), ops._p(ref), ops.dt(yd),
                                             B, H, W, Ci, ops._stream()), "bn_relu_bwd_reduce")
    c = host(coefs).double()
    gxh, yh = host(gx0).double(), yprev.double()
    sh = lambda v: v.reshape(1, -1, 1, 1)  # noqa: E731
    mask = (torch.addcmul(sh(c[3]).float(), yprev, sh(c[2]).float()) > 0).double()   # fp32 fma like the kernels (ties aside)
    gm = gxh * mask
    cpu = torch.stack([gm.sum((0, 2, 3)), (gm * (yh - sh(c[0])) * sh(c[1])).sum((0, 2, 3))])
    scale = cpu.abs().max().item() + 1e-30
    res.append((f"bnbwd_sums_vs_kernel_{tag}", (sums.cpu() - ref.cpu()).abs().max().item() / scale, 1e-5))
    res.append((f"bnbwd_sums_vs_cpu_{tag}", (sums.cpu() - cpu).abs().max().item() / scale, 2e-4))
    return res


def _outconv_bnbwd_case(B, C, K, H, W, dt_, seed):
    """OutConv backward with the BatchNorm-backward reduction of the stage below fused in: gx / dW / db as without
    the fusion, sums equal to unetb200_bn_relu_bwd_reduce on the stored gx."""
    res = []
    g = gen(seed)
    yprev = rq(torch.randn(B, C, H, W, generator=g), dt_)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    w = (torch.randn(K, C, generator=g) * 0.3).to(DEV)
    gl = dev_nhwc(rq(torch.randn(B, K, H, W, generator=g), dt_), dt_)
    yd = dev_nhwc(yprev, dt_)
    stats = torch.stack([yprev.double().sum((0, 2, 3)), (yprev.double() ** 2).sum((0, 2, 3))]).reshape(-1).to(DEV)
    coefs = ops.bn_finalize(stats, B * H * W, gamma.to(DEV), beta.to(DEV), 1e-5, 0.0, None, None, C)
    x = ops.empty_nhwc(B, C, H, W, dt_, DEV)
    ops.bn_relu_apply(yd, coefs, x, None)
    tag = f"{C}to{K}_{B}x{H}x{W}_{str(dt_)[6:]}"
    outs = []
    was, ops.OUTCONV_FUSE = ops.OUTCONV_FUSE, True            # opt-in in the product (slower at the path's shape)
    for below in (None, (yd, coefs)):
        gx = ops.empty_nhwc(B, C, H, W, dt_, DEV)
        dw = torch.empty((K, C, 1, 1), device=DEV)
        db = torch.empty(K, device=DEV)
        glp = gl.permute(0, 2, 3, 1).contiguous()          # packed NHWC logits gradient [B, H, W, K]
        sums = ops.outconv_bwd(x, w, glp, gx, dw, db, below=below)
        outs.append((gx, dw, db, sums))
    ops.OUTCONV_FUSE = was
    (gx0, dw0, db0, s0), (gx1, dw1, db1, s1) = outs
    res.append((f"outconv_bnbwd_fused_{tag}", 0.0 if (s0 is None and s1 is not None) else 1.0, 0.0))
    if s1 is None:
        return res
    res.append((f"outconv_bnbwd_gx_bitequal_{tag}", (host(gx1) - host(gx0)).abs().max().item(), 0.0))
    res.append((f"outconv_bnbwd_dw_{tag}", rel(host(dw1), host(dw0)), 1e-5))
    res.append((f"outconv_bnbwd_db_{tag}", rel(host(db1), host(db0)), 1e-5))
    ref = torch.zeros((2, C), dtype=torch.float64, device=DEV)
    _lib.check(ops.lib().unetb200_bn_relu_bwd_reduce(ops._p(gx0), ops.nhwc_ld(gx0), ops._p(yd), ops.nhwc_ld(yd),
                                                     ops._p(coefs[2]), ops._p(coefs[3]), ops._p(coefs[0]), ops._p(coefs[1]),
                                                     ops._p(ref), ops.dt(yd), B, H, W, C, ops._stream()), "bn_relu_bwd_reduce")
    scale = ref.abs().max().item() + 1e-30
    res.append((f"outconv_bnbwd_sums_{tag}", (s1 - ref).abs().max().item() / scale, 1e-5))
    return res


def _pool_bnreduce_case(B, C, H, W, dt_, seed):
    """Accumulating max-pool backward fused with the BatchNorm-backward reduction: gx bit-equal to the plain kernel,
    sums equal to unetb200_bn_relu_bwd_reduce on that gx."""
    res = []
    g = gen(seed)
    y = rq(torch.randn(B, C, H, W, generator=g), dt_)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    yd = dev_nhwc(y, dt_)
    stats = torch.stack([y.double().sum((0, 2, 3)), (y.double() ** 2).sum((0, 2, 3))]).reshape(-1).to(DEV)
    coefs = ops.bn_finalize(stats, B * H * W, gamma.to(DEV), beta.to(DEV), 1e-5, 0.0, None, None, C)
    x = ops.empty_nhwc(B, C, H, W, dt_, DEV)
    ops.bn_relu_apply(yd, coefs, x, None)
    gp = dev_nhwc(rq(torch.randn(B, C, H // 2, W // 2, generator=g), dt_), dt_)
    g0 = rq(torch.randn(B, C, H, W, generator=g), dt_)
    ga, gb = in_slice(g0, dt_, 64), in_slice(g0, dt_, 64)            # the skip half of a wider concat-gradient buffer
    ops.maxpool2_bwd(x, gp, ga, accumulate=True)
    ref = torch.zeros((2, C), dtype=torch.float64, device=DEV)
    _lib.check(ops.lib().unetb200_bn_relu_bwd_reduce(ops._p(ga), ops.nhwc_ld(ga), ops._p(yd), ops.nhwc_ld(yd), ops._p(coefs[2]),
                                                     ops._p(coefs[3]), ops._p(coefs[0]), ops._p(coefs[1]), ops._p(ref),
                                                     ops.dt(yd), B, H, W, C, ops._stream()), "bn_relu_bwd_reduce")
    sums = ops.maxpool2_bwd_bnreduce(x, gp, gb, yd, coefs)
    tag = f"{C}_{B}x{H}x{W}_{str(dt_)[6:]}"
    res.append((f"pool_bnreduce_gx_bitequal_{tag}", (host(gb) - host(ga)).abs().max().item(), 0.0))
    scale = ref.abs().max().item() + 1e-30
    res.append((f"pool_bnreduce_sums_{tag}", (sums - ref).abs().max().item() / scale, 1e-5))
    return res


def check_dgrad_bnbwd():
    out = []
    out += _pool_bnreduce_case(2, 64, 16, 24, BF, 91)
    out += _pool_bnreduce_case(1, 128, 9, 7, BF, 92)               # odd sizes: the last row / column has no window
    out += _pool_bnreduce_case(2, 16, 10, 12, FP, 93)
    out += _pool_bnreduce_case(1, 6, 8, 8, FP, 94)                  # scalar path
    out += _outconv_bnbwd_case(2, 64, 2, 24, 20, BF, 81)
    out += _outconv_bnbwd_case(1, 64, 4, 33, 17, BF, 82)
    out += _outconv_bnbwd_case(2, 16, 3, 9, 11, FP, 83)
    out += _dgrad_bnbwd_case(2, 64, 64, 16, 24, BF, 71)
    out += _dgrad_bnbwd_case(1, 128, 128, 20, 12, BF, 72)
    out += _dgrad_bnbwd_case(2, 128, 256, 9, 7, BF, 73)            # ragged tiles: rows outside the M grid must not count
    # narrow layers: the TMA-staged kernel's epilogue (conv_halo.cu, MODE 2), partial tiles, several tiles per CTA,
    # 8-channel tensors, yprev as a channel slice
    out += _dgrad_bnbwd_case(2, 16, 16, 40, 70, BF, 170)
    out += _dgrad_bnbwd_case(8, 16, 16, 256, 256, BF, 171)
    out += _dgrad_bnbwd_case(2, 32, 32, 33, 20, BF, 172)
    out += _dgrad_bnbwd_case(1, 16, 32, 9, 50, BF, 173)
    out += _dgrad_bnbwd_case(2, 32, 64, 24, 24, BF, 174)
    out += _dgrad_bnbwd_case(2, 8, 8, 37, 41, BF, 175)
    out += _dgrad_bnbwd_case(2, 32, 16, 16, 48, BF, 176, slice_y=True)
    out += _dgrad_bnbwd_case(3, 64, 128, 20, 8, BF, 74)            # stacked sub-tile geometry (W <= 8)
    out += _dgrad_bnbwd_case(2, 64, 64, 40, 36, BF, 75, slice_y=True)
    out += _dgrad_bnbwd_case(1, 256, 128, 33, 17, BF, 76)
    out += _dgrad_bnbwd_case(2, 64, 64, 16, 24, FP, 77)            # tf32 engine, fp32 storage
    out += _dgrad_bnbwd_case(1, 128, 64, 19, 21, FP, 78)
    return out


GROUPS = {
    "dgrad_bnbwd": lambda gd: check_dgrad_bnbwd(),
    "layout": lambda gd: check_layout_ops(),
    "bn_fwd": lambda gd: check_bn_forward(),
    "maxpool": lambda gd: check_maxpool(),
    "bn_bwd": lambda gd: check_bn_backward(),
    "upsample": lambda gd: check_upsample(),
    "ce_dice": lambda gd: check_ce_dice(),
    "dice": lambda gd: check_dice(gd),
    "boundary": lambda gd: check_boundary(gd),
    "outconv": lambda gd: check_outconv(),
    "conv_simt": lambda gd: check_conv_simt(),
    "conv_narrow": lambda gd: check_conv_narrow(),
    "narrow_bounds": lambda gd: check_narrow_bounds(),
    "conv_tc_first": lambda gd: check_conv_tc_fprop_small(),
    "conv_tc": lambda gd: check_conv_tc(),
    "conv_tc_tf32": lambda gd: check_conv_tc_tf32(),
    "conv_tc_x3": lambda gd: check_conv_tc_x3(),
    "conv_layouts": lambda gd: check_conv_layouts(),
    "optim": lambda gd: check_optim(),
    "conv_bnfold": lambda gd: check_conv_bnfold(),
    "parts_fp32": lambda gd: check_parts(gd, "fp32"),
    "parts_bf16": lambda gd: check_parts(gd, "bf16"),
    "calib_small": lambda gd: sum((calibrate(1, 2, False, 2, 128, 128, m) for m in ("fp32", "tf32", "bf16")), []),
    "calib_large": lambda gd: sum((calibrate(1, 2, False, 4, 256, 256, m) for m in ("fp32", "tf32", "bf16")), []),
    "calib_bil": lambda gd: sum((calibrate(1, 2, True, 2, 256, 256, m) for m in ("tf32", "bf16")), []),
}


def all_groups():
    """Op-level groups (this module) + end-to-end UNet gates (gpu_e2e)."""
    import gpu_e2e
    d = dict(GROUPS)
    d.update(gpu_e2e.GROUPS)
    return d


def load_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_v1.pt"), weights_only=False)
