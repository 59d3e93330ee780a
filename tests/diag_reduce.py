"""Diagnostic: unetb200_wgrad_reduce_multi against unetb200_wgrad_reduce on random partials."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
from unetb200 import _lib, ops
L = ops.lib()
DEV = "cuda"
cases = [(3, 9, 4, 8, 8, "oihw"), (5, 9, 64, 64, 64, "cl"), (2, 9, 128, 256, 256, "cl"), (7, 1, 128, 256, 64, "convT_cl"), (148, 9, 64, 64, 64, "oihw")]
jobs = (_lib.ReduceJob * len(cases))()
keep, refs = [], []
for j, (splits, ntaps, Cin, N, Cq, lay) in enumerate(cases):
    K = ntaps * Cin
    part = torch.randn(splits, K, N, device=DEV)
    if lay == "oihw":      # dst[co][c][t]
        dst = torch.zeros(N, Cin, ntaps, device=DEV); st, sc, sq, sn = 1, ntaps, 0, Cin * ntaps
    elif lay == "cl":      # dst[co][t][c]
        dst = torch.zeros(N, ntaps, Cin, device=DEV); st, sc, sq, sn = Cin, 1, 0, ntaps * Cin
    else:                  # convT channels_last [ci][q][co]
        dst = torch.zeros(Cin, N // Cq, Cq, device=DEV); st, sc, sq, sn = 0, N, Cq, 1
    ref = torch.zeros_like(dst)
    _lib.check(L.unetb200_wgrad_reduce(ops._p(part), splits, ntaps, Cin, N, Cq, ops._p(ref), st, sc, sq, sn, 0, ops._stream()), "ref")
    jobs[j].partials, jobs[j].dst = part.data_ptr(), dst.data_ptr()
    jobs[j].st, jobs[j].sc, jobs[j].sq, jobs[j].sn = st, sc, sq, sn
    jobs[j].splits, jobs[j].ntaps, jobs[j].Cin, jobs[j].N, jobs[j].Cq, jobs[j].accumulate = splits, ntaps, Cin, N, Cq, 0
    keep.append((part, dst)); refs.append(ref)
_lib.check(L.unetb200_wgrad_reduce_multi(jobs, len(cases), ops._stream()), "multi")
torch.cuda.synchronize()
for (c, (part, dst), ref) in zip(cases, keep, refs):
    print(c, "max abs diff", (dst - ref).abs().max().item(), "ref max", ref.abs().max().item())
