#!/usr/bin/env python
"""Training-step time of the width variants (UNet_S / UNet_T / UNet_SA, reference unet_model.py:52-189; UNet_S is what
train.py:253 builds by default) at B = 16, 512x512, bf16, as a CUDA graph and with eager launches, with the per-kernel-class profile."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
import unet.unet_model as UM  # noqa: E402
from unetb200 import losses as UL, ops  # noqa: E402
from unetb200.optim import FusedRMSprop  # noqa: E402

dev = torch.device("cuda:0")
names = sys.argv[1:] or ["UNet_S", "UNet_SA", "UNet_T", "UNet"]
for name in names:
    torch.manual_seed(0)
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
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # the same step replayed as a CUDA graph (what bench.py times): eager launches are CPU-bound on the light variants
    from unetb200.graph import GraphedStep

    def gstep(xx, tt):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", enabled=True):
            loss = UL.training_criterion(m(xx), tt, boundary_coeff=0.2, edge_width=51, edge_weight=7)
        loss.backward()
        opt.step(clip_max_norm=1.0)
        return loss.detach()
    gs = GraphedStep(gstep, (x, t), warmup=2)
    for _ in range(3):
        gs.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        gs.replay()
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / 20
    del gs
    with ops.profile() as rec:
        step()
        torch.cuda.synchronize()
    agg = {}
    for nm, s, e, fl, nb in rec:
        k = nm.split("[")[0]
        agg[k] = agg.get(k, 0.0) + s.elapsed_time(e)
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:8]
    print(f"{name}: graph {gms:.2f} ms/step ({16 / gms * 1e3:.0f} img/s), eager {ms:.2f} ms/step;  " + ", ".join(f"{k} {v:.2f}" for k, v in top), flush=True)
    if os.environ.get("PER_SHAPE"):
        full = {}
        for nm, s, e, fl, nb in rec:
            d = full.setdefault(nm, [0.0, 0, 0.0])
            d[0] += s.elapsed_time(e); d[1] += 1; d[2] += fl
        for nm, (t, n, fl) in sorted(full.items(), key=lambda kv: -kv[1][0])[:24]:
            print(f"    {nm:60s} {t:7.3f} ms  x{n}  {fl / t / 1e9 if t else 0:7.1f} TF/s")
