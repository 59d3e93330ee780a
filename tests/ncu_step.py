"""One profiled UNet training step (BASELINE.json configs[1]) for Nsight Compute:
    ncu --profile-from-start off ... python tests/ncu_step.py [batch] [size] [UNet|UNet_S|UNet_T|UNet_SA]
Two warm-up steps run outside the capture range; cudaProfilerStart/Stop bracket the third."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
import unet  # noqa: E402
from unetb200 import losses as UL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
torch.manual_seed(0)
NAME = sys.argv[3] if len(sys.argv) > 3 else "UNet"
import unet.unet_model as _UM  # noqa: E402
model = getattr(_UM, NAME)(1, 2, False).to(dev).to(memory_format=torch.channels_last).train()
from unetb200.optim import FusedRMSprop  # noqa: E402
opt = FusedRMSprop(model.parameters(), lr=1e-5, weight_decay=1e-8, momentum=0.999)
x = torch.rand(B, 1, S, S, device=dev).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 2, (B, S, S), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", enabled=True):
        loss = UL.training_criterion(model(x), t, boundary_coeff=0.2, edge_width=51, edge_weight=7)
    loss.backward()
    opt.step(clip_max_norm=1.0)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
