#!/usr/bin/env python
"""Host-side profile (cProfile) of eager training steps of a light variant: where the Python / ctypes launch path
spends its time when the GPU is faster than the launches (UNet_S: ~230 launches per step)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
import unet.unet_model as UM  # noqa: E402
from unetb200 import losses as UL  # noqa: E402
from unetb200.optim import FusedRMSprop  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "UNet_S"
dev = torch.device("cuda:0")
m = getattr(UM, name)(1, 2, False).to(dev).to(memory_format=torch.channels_last).train()
opt = FusedRMSprop(m.parameters(), lr=1e-5, weight_decay=1e-8, momentum=0.999)
x = torch.rand(16, 1, 512, 512, device=dev).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 2, (16, 512, 512), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", enabled=True):
        loss = UL.training_criterion(m(x), t, boundary_coeff=0.2, edge_width=51, edge_weight=7)
    loss.backward()
    opt.step(clip_max_norm=1.0)


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
st.sort_stats("cumulative").print_stats(45)
