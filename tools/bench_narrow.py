#!/usr/bin/env python
"""Micro-benchmark of the narrow-channel 3x3 layers of UNet_S / UNet_T (B=16): fprop (+ BatchNorm statistics), dgrad
and wgrad per layer, CUDA events, inputs > L2 rotate over 3 buffers.  Prints the HBM floor (bytes / measured copy peak)
next to each.  UNETB200_NO_HALO=1 selects the thread-built-im2col kernels (conv_narrow.cu) for an A/B run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
from unetb200 import _lib, ops  # noqa: E402
from unetb200 import functional as UF  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
PEAK = 6.5e12
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shapes = [(1, 16, 512), (1, 8, 512), (16, 16, 512), (32, 16, 512), (16, 32, 256), (32, 32, 256), (64, 32, 256), (32, 64, 128), (8, 8, 512), (16, 8, 512)]


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
    x = ops.empty_nhwc(B, Ci, H, W, BF, DEV).normal_()
    gy = ops.empty_nhwc(B, Co, H, W, BF, DEV).normal_()
    y = ops.empty_nhwc(B, Co, H, W, BF, DEV)
    gx = ops.empty_nhwc(B, Ci, H, W, BF, DEV)
    w = torch.randn(Co, Ci, 3, 3, device=DEV) / (3 * Ci ** 0.5)
    wf, wd = UF.pack3x3_fprop(w, BF), UF.pack3x3_dgrad(w, BF)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device=DEV)
    dW = torch.empty(Co, Ci, 3, 3, device=DEV)
    df = ops.make_gconv(ops._DT[BF], _lib.ALGO_AUTO, B, H, W, Ci, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(x), Co, 1, 1, (0, 0), H, W, ops.nhwc_ld(y))
    dd = ops.make_gconv(ops._DT[BF], _lib.ALGO_AUTO, B, H, W, Co, ops.TAPS3, 1, (0, 0), H, W, ops.nhwc_ld(gy), Ci, 1, 1, (0, 0), H, W, ops.nhwc_ld(gx))
    floor = B * H * W * (Ci + Co) * 2 / PEAK * 1e3
    tf = timed(lambda: ops.gconv_fprop(df, x, wf, None, y, stats))
    td = timed(lambda: ops.gconv_fprop(dd, gy, wd, None, gx, None)) if Ci >= 8 else float("nan")
    tw = timed(lambda: ops.gconv_wgrad(df, x, gy, dW, 1, 9, Ci * 9))
    print(f"{Ci:3d}->{Co:3d} @{H}x{W} B={B}: fprop+stats {tf:.3f} ms  dgrad {td:.3f} ms  wgrad(+reduce) {tw:.3f} ms   HBM floor {floor:.3f} ms", flush=True)
