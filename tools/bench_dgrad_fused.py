#!/usr/bin/env python
"""Micro-benchmark: plain 3x3 dgrad vs dgrad with the BatchNorm-backward reduction fused into its epilogue, and the
separate reduction pass, at the full-size layer shapes (B=16).  CUDA events, inputs > L2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
from unetb200 import _lib, ops  # noqa: E402
from unetb200 import functional as UF  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
shapes = [(16, 64, 64, 512), (16, 128, 128, 256), (16, 64, 128, 512), (16, 256, 256, 128)]   # B, Ci (dgrad N), Co, H
for (B, Ci, Co, H) in shapes:
    W = H
    gy = ops.empty_nhwc(B, Co, H, W, BF, DEV).normal_()
    yprev = ops.empty_nhwc(B, Ci, H, W, BF, DEV).normal_()
    gx = ops.empty_nhwc(B, Ci, H, W, BF, DEV)
    w = torch.randn(Co, Ci, 3, 3, device=DEV) / (3 * Co ** 0.5)
    wd = UF.pack3x3_dgrad(w, BF)
    stats = torch.zeros(2 * Ci, dtype=torch.float64, device=DEV)
    stats[Ci:] = B * H * W
    coefs = ops.bn_finalize(stats, B * H * W, None, None, 1e-5, 0.0, None, None, Ci)
    d = ops.make_gconv(ops._DT[BF], _lib.ALGO_TC, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gy),
                       Ci, 1, 1, (0, 0), H, W, ops.nhwc_ld(gx))

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    sums = torch.zeros((2, Ci), dtype=torch.float64, device=DEV)
    L = ops.lib()

    def reduce_pass():
        _lib.check(L.unetb200_bn_relu_bwd_reduce(ops._p(gx), ops.nhwc_ld(gx), ops._p(yprev), ops.nhwc_ld(yprev), ops._p(coefs[2]),
                                                 ops._p(coefs[3]), ops._p(coefs[0]), ops._p(coefs[1]), ops._p(sums), ops.dt(yprev),
                                                 B, H, W, Ci, ops._stream()), "reduce")
    t_plain = timed(lambda: ops.gconv_fprop(d, gy, wd, None, gx, None, kind="dgrad"))
    t_fused = timed(lambda: ops.gconv_dgrad_bnbwd(d, gy, wd, gx, yprev, coefs))
    t_red = timed(reduce_pass)
    gf = 2.0 * B * H * W * Ci * 9 * Co / 1e9
    print(f"dgrad {Co}->{Ci} @{H}x{W}: plain {t_plain:.3f} ms ({gf / t_plain:.0f} TF/s)  fused {t_fused:.3f} ms  "
          f"separate reduce {t_red:.3f} ms  -> fused saves {t_plain + t_red - t_fused:+.3f} ms", flush=True)
