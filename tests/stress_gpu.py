"""Determinism / race hunting on the GPU: repeat the same backward pass and report which gradients vary
between runs.  Usage: python tests/stress_gpu.py [unet|bnbwd] [repeats]"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gpu_checks as G  # noqa: E402
from gpu_checks import DEV, O, ops  # noqa: E402


def unet_repeat(n, bilinear=True, size=32, mode="fp32"):
    import unet
    os.environ["UNET_B200_PRECISION"] = "fp32" if mode != "tf32" else "tf32"
    st = O.build_state(1, 2, bilinear, seed=0)
    img, msk = O.synthetic_batch(2, 1, 2, size, size)
    _, _, ref = O.training_step({k: v.clone() for k, v in st.items()}, img, msk, 2, bilinear)
    model = unet.UNet(1, 2, bilinear)
    model.load_state_dict(st)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    first = None
    bad_runs = 0
    for it in range(n):
        _, _, grads = G.unet_step_gpu(model, img, msk, amp=(mode == "bf16"), fused=(it % 2 == 0))
        worst = max((O.rel_err(grads[k], ref[k]), k) for k in ref)
        if first is None:
            first = grads
        var = [(O.rel_err(grads[k], first[k]), k) for k in ref]
        vmax = max(var)
        flag = "BAD" if worst[0] > 1e-3 else "ok "
        bad_runs += worst[0] > 1e-3
        print(f"run {it:3d} {flag} worst-vs-oracle {worst[0]:.3e} [{worst[1]}]  max-vs-run0 {vmax[0]:.3e} [{vmax[1]}]")
        if worst[0] > 1e-3:
            for e, k in sorted(var, reverse=True)[:6]:
                d = (grads[k] - ref[k]).abs().reshape(grads[k].shape[0], -1).max(1).values
                idx = torch.nonzero(d > 1e-3 * ref[k].abs().max()).flatten().tolist()
                print(f"        {k:<48s} vs-run0 {e:.3e} vs-oracle {O.rel_err(grads[k], ref[k]):.3e} bad rows {idx[:16]}{'...' if len(idx) > 16 else ''} of {grads[k].shape[0]}")
    print(f"unet_repeat: {bad_runs}/{n} bad runs")


def bnbwd_repeat(n):
    g = G.gen(0)
    B, C, H, W = 2, 512, 4, 4
    y = torch.randn(B, C, H, W, generator=g)
    gz = torch.randn(B, C, H, W, generator=g)
    yd, gzd = G.dev_nhwc(y, torch.float32), G.dev_nhwc(gz, torch.float32)
    stats = torch.stack([y.double().sum((0, 2, 3)), (y.double() ** 2).sum((0, 2, 3))]).to(DEV).reshape(-1)
    coefs = ops.bn_finalize(stats, B * H * W, None, None, 1e-5, 0.0, None, None, C)
    ref = None
    bad = 0
    junk = []
    for it in range(n):
        junk.append(torch.randn(1 + (it * 37) % 5000, device=DEV))
        if len(junk) > 7:
            junk.pop(0)
        gy, dg, db = ops.bn_relu_bwd(gzd, yd, coefs, True)
        cur = (gy.clone(), dg.clone(), db.clone())
        if ref is None:
            ref = cur
        errs = [O.rel_err(a, b) for a, b in zip(cur, ref)]
        if max(errs) > 1e-5:
            bad += 1
            print(f"bnbwd run {it}: gy {errs[0]:.3e} dgamma {errs[1]:.3e} dbeta {errs[2]:.3e}")
    print(f"bnbwd_repeat: {bad}/{n} deviating runs")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "unet"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    if what == "bnbwd":
        bnbwd_repeat(n)
    else:
        unet_repeat(n, bilinear=(os.environ.get("STRESS_BILINEAR", "1") == "1"),
                    size=int(os.environ.get("STRESS_SIZE", "32")), mode=os.environ.get("STRESS_MODE", "fp32"))
