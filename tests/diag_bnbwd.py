"""Diagnostic: run-to-run determinism and correctness of the fused dgrad + BatchNorm-backward reduction."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G
from unetb200 import ops, _lib
from unetb200 import functional as UF
DEV = "cuda"
BF = torch.bfloat16
for (B, Ci, Co, H, W) in ((2, 256, 256, 16, 16), (2, 512, 512, 8, 8), (2, 128, 128, 32, 32), (2, 512, 256, 16, 16), (2, 1024, 512, 8, 8)):
    g = G.gen(5)
    gy = G.rq(torch.randn(B, Co, H, W, generator=g), BF)
    w = torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Co ** 0.5)
    yprev = G.rq(torch.randn(B, Ci, H, W, generator=g), BF)
    gamma, beta = torch.rand(Ci, generator=g) + 0.5, torch.randn(Ci, generator=g) * 0.3
    gyd, yd = G.dev_nhwc(gy, BF), G.dev_nhwc(yprev, BF)
    stats = torch.stack([yprev.double().sum((0, 2, 3)), (yprev.double() ** 2).sum((0, 2, 3))]).reshape(-1).to(DEV)
    coefs = ops.bn_finalize(stats, B * H * W, gamma.to(DEV), beta.to(DEV), 1e-5, 0.0, None, None, Ci)
    wd = UF.pack3x3_dgrad(w.to(DEV), BF)
    outs = []
    for rep in range(6):
        junk = torch.randn(1 << 22, device=DEV)            # perturb the allocator / memory contents
        gx = ops.empty_nhwc(B, Ci, H, W, BF, DEV)
        d = ops.make_gconv(ops._DT[BF], _lib.ALGO_TC, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gyd),
                           Ci, 1, 1, (0, 0), H, W, ops.nhwc_ld(gx))
        sums = ops.gconv_dgrad_bnbwd(d, gyd, wd, gx, yd, coefs)
        torch.cuda.synchronize()
        outs.append((sums.cpu().clone(), gx.float().cpu().clone()))
        del junk
    ref = torch.zeros((2, Ci), dtype=torch.float64, device=DEV)
    _lib.check(ops.lib().unetb200_bn_relu_bwd_reduce(ops._p(gx), ops.nhwc_ld(gx), ops._p(yd), ops.nhwc_ld(yd), ops._p(coefs[2]),
               ops._p(coefs[3]), ops._p(coefs[0]), ops._p(coefs[1]), ops._p(ref), ops.dt(yd), B, H, W, Ci, ops._stream()), "r")
    ref = ref.cpu()
    sc = ref.abs().max().item()
    print(f"{B}x{Co}->{Ci} {H}x{W}:", " ".join(f"[run{r} sums-vs-run0 {(s - outs[0][0]).abs().max().item() / sc:.1e} gx {(x - outs[0][1]).abs().max().item():.1e} vs-ref {(s - ref).abs().max().item() / sc:.1e}]" for r, (s, x) in enumerate(outs)))
