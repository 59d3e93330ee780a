"""Inference throughput of BASELINE.json configs[4] on one B200: UNet(3, 4, bilinear=False).eval(), B=8, 3x1024x1024,
bf16 autocast, predict.py-style (forward -> identity-size bilinear resize -> argmax), random-init weights, synthetic
input resident in HBM.  Times the forward with eval-mode BatchNorm + ReLU folded into the conv epilogues (default)
against the unfolded path (UNETB200_NO_BN_FOLD=1: conv, then a separate BatchNorm-apply pass), both with CUDA events.

    python tools/bench_infer.py > profiles/r1_infer_bench.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
import unet  # noqa: E402
from unetb200 import eval_tail as UE  # noqa: E402
from unetb200 import ops  # noqa: E402

B, S = 8, 1024
FWD_GF = 12334.07          # algorithmic conv GFLOP per B=8 forward at configs[4]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = unet.UNet(3, 4, False).to(dev).to(memory_format=torch.channels_last).eval()
x = torch.rand(B, 3, S, S, device=dev).contiguous(memory_format=torch.channels_last)


def run(tail):
    with torch.inference_mode(), torch.autocast("cuda", enabled=True):
        lg = model(x)
        return UE.resize_argmax(lg, (S, S)) if tail else lg


def timed(tail, iters=10):
    for _ in range(3):
        run(tail)
    torch.cuda.synchronize()
    n0 = ops.LAUNCHES
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        run(tail)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, (ops.LAUNCHES - n0) // iters


if "--once" in sys.argv:                 # a single forward + tail, for an ncu capture
    run(True)
    torch.cuda.synchronize()
    sys.exit(0)

out = {"workload": f"UNet(3,4,False).eval() bf16, B={B}, {S}x{S} (BASELINE.json configs[4]), random init, synthetic",
       "device": torch.cuda.get_device_name(0), "results": {}}
labels = {}
for name, env in (("bn_folded", None), ("bn_separate_pass", "1")):
    if env:
        os.environ["UNETB200_NO_BN_FOLD"] = env
    else:
        os.environ.pop("UNETB200_NO_BN_FOLD", None)
    labels[name] = run(True).clone()
    ms, launches = timed(False)
    ms_tail, _ = timed(True)
    out["results"][name] = {"forward_ms": ms, "img_per_s": B / ms * 1e3, "conv_tflops": FWD_GF / ms,
                            "forward_plus_predict_tail_ms": ms_tail, "launches_per_forward": launches}
os.environ.pop("UNETB200_NO_BN_FOLD", None)
# where the folded forward spends its time: CUDA events around every C-ABI call of one forward
with ops.profile() as rec:
    run(True)
torch.cuda.synchronize()
prof = ops.summarize_profile(rec)
classes = {}
for name, d in prof.items():
    key = name.split("[")[0].split("@")[0]
    c = classes.setdefault(key, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
    for k in c:
        c[k] += d[k]
tot = sum(c["ms"] for c in classes.values())
out["profile_one_forward"] = {k: {"ms": round(c["ms"], 4), "calls": c["calls"], "share": round(c["ms"] / tot, 4),
                                  "tflops": round(c["flops"] / c["ms"] / 1e9, 1) if c["flops"] else None,
                                  "gbs": round(c["bytes"] / c["ms"] / 1e6, 1) if c["bytes"] else None}
                              for k, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"])}
out["label_agreement_folded_vs_separate"] = (labels["bn_folded"] == labels["bn_separate_pass"]).float().mean().item()
out["speedup_from_fold"] = out["results"]["bn_separate_pass"]["forward_ms"] / out["results"]["bn_folded"]["forward_ms"]
print(json.dumps(out, indent=1))
