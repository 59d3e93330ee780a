"""Diagnostic: run-to-run and side-stream-vs-single-stream gradient differences for one bf16 backward."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200")); sys.path.insert(0, ROOT)
from oracle import unet_oracle as O
import unet
from unetb200 import losses as UL, ops
DEV = torch.device("cuda:0")
amp = (sys.argv[1] != "fp32") if len(sys.argv) > 1 else True
os.environ["UNET_B200_PRECISION"] = "fp32"
st = O.build_state(1, 2, False, seed=0)
img, msk = O.synthetic_batch(2, 1, 2, 64, 64)
x = img.to(DEV).contiguous(memory_format=torch.channels_last); t = msk.to(DEV)
def grads_of(side):
    ops._SIDE_ON = side
    ops._WGRAD_SIDE = side
    m = unet.UNet(1, 2, False); m.load_state_dict(st)
    m = m.to(DEV).to(memory_format=torch.channels_last).train()
    with torch.autocast("cuda", enabled=amp):
        loss = UL.training_criterion(m(x), t, boundary_coeff=0.2)
    loss.backward(); torch.cuda.synchronize()
    return {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()}, float(loss)
runs = [("single", False), ("single", False), ("side", True), ("side", True), ("single", False)]
res = [grads_of(s) for _, s in runs]
base = res[0][0]
for (name, _), (g, l) in zip(runs, res):
    errs = sorted(((O.rel_l2(g[k], base[k]), k) for k in base), reverse=True)
    print(f"{name:7s} loss {l:.6f}  worst rel-L2 vs run0: " + ", ".join(f"{e:.2e} {k}" for e, k in errs[:4]))
    if errs[0][0] > 1e-4 and os.environ.get("DIAG_ALL"):
        for k in reversed(list(base)):           # backward order
            print(f"      {O.rel_l2(g[k], base[k]):.2e} {k}")
