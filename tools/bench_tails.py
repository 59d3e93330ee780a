"""Micro-benchmark of the kernels either side of the step on one B200:
evaluate tail (evaluate.py:111-117), predict tail (predict.py:26-27), uint8 input pipeline (data_loading.py:65-132).

    python tools/bench_tails.py [--once] > profiles/r1_tails_bench.json

For each op: algorithmic bytes (one read of every input, one write of every output) / average launch time measured
with CUDA events on the launching stream, against MEASURED_PEAKS.json's copy bandwidth, with the reference's own
expression (the ATen chain evaluate.py / predict.py / data_loading.py + train.py:113-114 run) timed beside it on the
same device.  Inputs rotate through enough distinct buffers to exceed twice the 126 MB L2, so every launch reads HBM.
``--once`` runs each kernel a single time (for an ncu capture)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200"))
from unetb200 import data as UD  # noqa: E402
from unetb200 import eval_tail as UE  # noqa: E402
from unetb200 import ops  # noqa: E402

ONCE = "--once" in sys.argv
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
PEAK = float(peaks.get("hbm_gbs", 6530.3))
L2_BYTES = 126e6


def timed(fn, sets, iters):
    """average DEVICE ms per call of fn(*sets[i]): one pass over all input sets is captured in a CUDA graph (the
    C-ABI calls only enqueue on the stream they are given) and replayed, so the ~40 us of Python / ctypes / allocator
    time per call does not hide the kernel time; CUDA events on the replaying stream."""
    for i in range(min(3, len(sets))):
        fn(*sets[i])
    torch.cuda.synchronize()
    if ONCE:
        return 0.0
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            keep = [fn(*a) for a in sets]
        reps = max(1, iters // len(sets))
        graph.replay()
        side.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(side)
        for _ in range(reps):
            graph.replay()
        e.record(side)
        side.synchronize()
    del keep
    return s.elapsed_time(e) / (reps * len(sets))


def timed_eager(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def ncopies(nbytes):
    return max(2, int(2 * L2_BYTES / nbytes) + 1)


def report(name, nbytes, ours_ms, ref_ms, note):
    return {"op": name, "algorithmic_bytes": nbytes, "ms": ours_ms, "gbs": nbytes / ours_ms / 1e6 if ours_ms else None,
            "frac_of_hbm_peak": nbytes / ours_ms / 1e6 / PEAK if ours_ms else None, "reference_chain_ms": ref_ms,
            "speedup_vs_reference_chain": (ref_ms / ours_ms) if ours_ms and ref_ms else None, "workload": note}


def main():
    out = []
    g = torch.Generator().manual_seed(0)
    iters = 1 if ONCE else 60

    # ---- evaluate tail: B=16, 3 classes, 512x512, bf16 NHWC logits (autocast output), float32 mask (evaluate.py:49)
    B, C, H, W = 16, 3, 512, 512
    nb = B * H * W * (C * 2 + 4 + 8)
    sets = []
    for _ in range(ncopies(nb)):
        lg = torch.randn(B, C, H, W, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        mt = torch.randint(0, 3, (B, H, W), generator=g).to(dev).float()
        sets.append((lg, mt))

    def ours(lg, mt):
        return UE.argmax_class_dice(lg, mt, c=2)

    def ref(lg, mt):
        idx = lg.argmax(dim=1)
        pred_c = (idx == 2).float()
        true_c = (mt == 2).float()
        inter = 2 * (pred_c * true_c).sum(dim=(-1, -2))
        sets_sum = pred_c.sum(dim=(-1, -2)) + true_c.sum(dim=(-1, -2))
        sets_sum = torch.where(sets_sum == 0, inter, sets_sum)
        return idx, ((inter + 1e-6) / (sets_sum + 1e-6)).mean()
    a, b = ours(*sets[0]), ref(*sets[0])
    assert torch.equal(a[0], b[0]) and abs(a[1].item() - b[1].item()) < 1e-6
    out.append(report("eval_counts (argmax + class-2 dice, int64 labels)", nb, timed(ours, sets, iters),
                      timed(ref, sets, iters), f"B={B} C={C} {H}x{W} bf16 NHWC logits, fp32 mask"))
    nb8 = B * H * W * (C * 2 + 4 + 1)

    def ours8(lg, mt):
        return UE.argmax_class_dice(lg, mt, c=2, index_dtype=torch.uint8)
    out.append(report("eval_counts (uint8 labels)", nb8, timed(ours8, sets, iters), None if ONCE else out[-1]["reference_chain_ms"],
                      "same, label map written as uint8 (what evaluate.py:128 converts it to)"))
    del sets

    # ---- predict tail: BASELINE configs[4] logits, B=8, 4 classes, 1024x1024 bf16, identity resize (predict.py:26)
    B, C, H, W = 8, 4, 1024, 1024
    nb = B * H * W * (C * 2 + 8)
    sets = [(torch.randn(B, C, H, W, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last),)
            for _ in range(ncopies(nb))]

    def ours_p(lg):
        return UE.resize_argmax(lg, (H, W))

    def ref_p(lg):
        return F.interpolate(lg, (H, W), mode="bilinear").argmax(dim=1)
    assert torch.equal(ours_p(*sets[0]), ref_p(*sets[0]))
    out.append(report("resize_argmax (identity resize)", nb, timed(ours_p, sets, iters), timed(ref_p, sets, iters),
                      f"B={B} C={C} {H}x{W} bf16 NHWC logits -> int64 labels"))
    sets = [(s[0][:, :, ::2, ::2].contiguous(memory_format=torch.channels_last),) for s in sets]
    nbu = B * (C * 2 * (H // 2) * (W // 2) + 8 * H * W)
    out.append(report("resize_argmax (2x upsample)", nbu, timed(ours_p, sets, iters), timed(ref_p, sets, iters),
                      f"B={B} C={C} {H // 2}x{W // 2} -> {H}x{W}"))
    del sets

    # ---- input pipeline: BASELINE configs[1] batch, 16 x 512 x 512 gray + masks, mixed rotations
    B, H, W = 16, 512, 512
    nb = B * H * W * (2 + 4) + B * H * W * (1 + 8)          # image bytes are read twice (any > 1 pass + convert)
    rots = [i % 4 for i in range(B)]
    drot = torch.tensor(rots, dtype=torch.int32, device=dev)
    sets = []
    for _ in range(ncopies(nb)):
        im = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, generator=g).to(dev)
        mk = torch.tensor([0, 128, 255], dtype=torch.uint8)[torch.randint(0, 3, (B, H, W), generator=g)].to(dev)
        sets.append((im, mk))
    L = ops.lib()

    def ours_d(im, mk):
        o = torch.empty((B, H, W, 1), dtype=torch.float32, device=dev)
        fl = torch.empty(B, dtype=torch.int32, device=dev)
        mo = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        ops._run("preprocess_image_u8", L.unetb200_preprocess_image_u8, ops._p(im), B, H, W, 1, ops._p(drot), 0, ops._p(o),
                 ops._p(fl), ops._stream(), kernels=2)
        ops._run("preprocess_mask_u8", L.unetb200_preprocess_mask_u8, ops._p(mk), B, H, W, ops._p(drot), 0, None, ops._p(mo),
                 ops._stream())
        return o, mo

    def ref_d(im, mk):
        # the reference's arithmetic moved to the device as plain torch ops (it runs them in numpy on the host)
        o = torch.stack([torch.rot90(im[i], rots[i]) for i in range(B)]).float() / 255.0
        m = torch.stack([torch.rot90(mk[i], rots[i]) for i in range(B)])
        mo = (m == 255).long() * 2 + (m == 128).long()
        return o, mo
    a, b = ours_d(*sets[0]), ref_d(*sets[0])
    assert torch.equal(a[1], b[1]) and (a[0].reshape(B, H, W) - b[0]).abs().max().item() < 1e-6
    out.append(report("preprocess_image_u8 + preprocess_mask_u8", nb, timed(ours_d, sets, iters), timed(ref_d, sets, iters),
                      f"B={B} {H}x{W} gray uint8 + mask, quarter turns {rots[:4]}..., device-resident bytes"))
    # end to end from pinned host bytes (what a loader hands over) vs the reference's host fp32/int64 batch + .to(device)
    if not ONCE:
        him = sets[0][0].cpu().pin_memory()
        hmk = sets[0][1].cpu().pin_memory()

        def e2e_ours():
            return UD.preprocess_batch(him, hmk, rots, device=dev)
        hf = (torch.stack([torch.rot90(him[i], rots[i]) for i in range(B)]).float() / 255.0).unsqueeze(1).pin_memory()
        hl = torch.stack([torch.rot90(hmk[i], rots[i]) for i in range(B)]).long().pin_memory()

        def e2e_ref():          # train.py:113-114 on an already pre-processed host batch (host preprocessing NOT counted)
            return (hf.to(device=dev, dtype=torch.float32, memory_format=torch.channels_last, non_blocking=True),
                    hl.to(device=dev, dtype=torch.long, non_blocking=True))
        t_o = timed_eager(e2e_ours, 30)
        t_r = timed_eager(e2e_ref, 30)
        out.append({"op": "input batch host -> HBM, ready for the step", "ms": t_o, "reference_chain_ms": t_r,
                    "speedup_vs_reference_chain": t_r / t_o, "h2d_bytes": B * H * W * 2 + 4 * B,
                    "reference_h2d_bytes": B * H * W * 12,
                    "workload": "pinned uint8 image+mask -> H2D -> kernels, vs H2D of the fp32 image + int64 mask the "
                                "reference's loader produced on the host (its host-side numpy work excluded)"})
    print(json.dumps({"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": PEAK, "results": out}, indent=1))


if __name__ == "__main__":
    main()
