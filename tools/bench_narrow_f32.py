#!/usr/bin/env python
"""fp32 narrow layers (no autocast): TF32 halo fprop / dgrad and the exact-fp32 CUDA-core weight gradient, per layer."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
from unetb200 import _lib, ops  # noqa: E402
from unetb200 import functional as UF  # noqa: E402

DEV, FP = "cuda", torch.float32
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shapes = [(16, 16, 512), (32, 16, 512), (16, 32, 256), (32, 32, 256), (64, 32, 256), (32, 64, 128)]


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


for (Ci, Co, H) in shapes:
    W = H
    x = ops.empty_nhwc(B, Ci, H, W, FP, DEV).normal_()
    gy = ops.empty_nhwc(B, Co, H, W, FP, DEV).normal_()
    y = ops.empty_nhwc(B, Co, H, W, FP, DEV)
    gx = ops.empty_nhwc(B, Ci, H, W, FP, DEV)
    w = torch.randn(Co, Ci, 3, 3, device=DEV) / (3 * Ci ** 0.5)
    wf, wd = UF.pack3x3_fprop(w, FP), UF.pack3x3_dgrad(w, FP)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)
    dW = torch.empty(Co, Ci, 3, 3, device=DEV)
    A = _lib.ALGO_PREFER_TC
    df = ops.make_gconv(ops._DT[FP], A, B, H, W, Ci, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(x), Co, 1, 1, (0, 0), H, W, ops.nhwc_ld(y))
    dd = ops.make_gconv(ops._DT[FP], A, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gy), Ci, 1, 1, (0, 0), H, W, ops.nhwc_ld(gx))
    tf = timed(lambda: ops.gconv_fprop(df, x, wf, None, y, stats))
    td = timed(lambda: ops.gconv_fprop(dd, gy, wd, None, gx, None))
    tw = timed(lambda: ops.gconv_wgrad(df, x, gy, dW, 1, 9, Ci * 9))
    gfma = B * H * W * 9.0 * Ci * Co / 1e9
    print(f"{Ci:3d}->{Co:3d} @{H}x{W} B={B} fp32: fprop+stats {tf:.3f} ms  dgrad {td:.3f} ms  wgrad(+reduce) {tw:.3f} ms "
          f"({gfma / tw:.1f} TFMA/s; HBM floor {B * H * W * (Ci + Co) * 4 / 6.5e12 * 1e3:.3f} ms)", flush=True)
