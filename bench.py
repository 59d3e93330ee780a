#!/usr/bin/env python
"""bench.py -- UNet 512x512 training step on N B200s (BASELINE.json configs[1] / [3]).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU fp32 path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1

One step = forward (bf16 autocast) + loss (CE + dice + 0.2*boundary_loss, the train.py:137-147 form)
+ backward (+ NCCL gradient all-reduce when N > 1) + clip_grad_norm_ + RMSprop step (train.py:80,153-159)
on a batch of 16 synthetic 1x512x512 images per GPU.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "unet-medical-image-contour-segmentation_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "unet512_train_images_per_sec"
UNIT = "img/s"
# algorithmic conv FLOPs per image, fwd+bwd, UNet(1,2,False) at 512x512 (SURVEY.md section 8d / BASELINE.md section 4)
GFLOP_PER_IMG = {(False, 512): 1154.004, (True, 512): 957.509}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--bilinear", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--per-layer", action="store_true", help="add a per-layer conv table (`layers`) to the JSON line")
    ap.add_argument("--torch-optim", action="store_true",
                    help="torch.optim.RMSprop + clip_grad_norm_ instead of the fused multi-tensor kernels (A/B)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the whole step as one CUDA graph (auto: try, fall back to eager launches)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return dict(hbm_gbs=float(d["hbm_gbs"]), tflops_burst=float(d["bf16_tflops"]),
                        tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
        except Exception:  # noqa: BLE001
            pass
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the oracle: torch-CPU fp32 restatement of the reference step)
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(size, bilinear, steps, warmup, budget_s):
    """img/s of the reference's CPU fp32 step on B=1 samples of the workload; bounded by budget_s."""
    from oracle import unet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.build_state(1, 2, bilinear, seed=0)
    img, msk = O.synthetic_batch(1, 1, 2, size, size)
    times = []
    t_start = time.perf_counter()
    done_warm = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.training_step({k: v.clone() for k, v in st.items()}, img, msk, 2, bilinear, boundary_coeff=0.2)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        else:
            done_warm += 1
        elapsed = time.perf_counter() - t_start
        if times and elapsed + dt > budget_s:
            break
        if not times and done_warm >= 1 and elapsed + 2 * dt > budget_s:
            warmup = done_warm          # cut the warm-up short: keep at least one timed step
    if not times:
        times = [dt]
    med = statistics.median(times)
    return 1.0 / med, len(times), done_warm, med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, n, nw, med = cpu_reference_rate(args.size, args.bilinear, args.steps, args.warmup, budget_s=150.0)
    cores = os.cpu_count() or 1
    sample = (f"oracle port of the reference CPU fp32 step (unet_parts/unet_model/dice/boundary via torch-CPU), "
              f"B=1 samples of the {args.size}x{args.size} workload, {nw} warm-up + {n} timed steps, median")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": nw, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"UNet(1,2,bilinear={args.bilinear}) training step, {args.size}x{args.size}, "
                               "CE+dice+0.2*boundary_loss (BASELINE.json configs[1])",
                   "per_step_sample": "1 image"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [s.strip() for s in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            # "under load": drop the lowest quartile (ramp-up samples)
            s = sorted(sm)
            out["sm_mhz"] = statistics.median(s[len(s) // 4:])
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def run_ours(args):
    import torch.distributed as dist
    import unet
    from unetb200 import ddp, ops
    from unetb200 import losses as UL

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the UNet hot path has no CPU fallback (use --impl reference for "
                         "the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=__import__("datetime").timedelta(seconds=180))
    B, S = args.batch, args.size

    torch.manual_seed(0)
    model = unet.UNet(1, 2, args.bilinear).to(dev).to(memory_format=torch.channels_last).train()
    if world > 1:
        ddp.broadcast_module_state(model)
    bucket_mb = int(os.environ.get("UNETB200_DDP_BUCKET_MB", "256"))   # one bucket: the step runs as graphs, nothing overlaps it
    reducer = ddp.GradAllReducer(model, bucket_bytes=bucket_mb << 20) if world > 1 else None
    if args.torch_optim:
        opt = torch.optim.RMSprop(model.parameters(), lr=1e-5, weight_decay=1e-8, momentum=0.999, foreach=True,
                                  capturable=(args.graph != "off"))

        def clip_and_step():
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
    else:
        from unetb200.optim import FusedRMSprop      # same arithmetic (tests: optim group), two launches
        opt = FusedRMSprop(model.parameters(), lr=1e-5, weight_decay=1e-8, momentum=0.999)

        def clip_and_step():
            opt.step(clip_max_norm=1.0)

    gi = torch.Generator().manual_seed(1 + 1000 * rank)
    gm = torch.Generator().manual_seed(2 + 1000 * rank)
    img_h = torch.rand(B, 1, S, S, generator=gi).pin_memory()
    msk_h = torch.randint(0, 2, (B, S, S), generator=gm, dtype=torch.long).pin_memory()
    img_d = img_h.to(dev).contiguous(memory_format=torch.channels_last)
    msk_d = msk_h.to(dev)

    def step(x, t):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", enabled=True):
            logits = model(x)
            loss = UL.training_criterion(logits, t, boundary_coeff=0.2, edge_width=51, edge_weight=7)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        clip_and_step()
        return loss

    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch():
        """H2D copy of one step's inputs from pinned host memory on the copy stream (a loader's prefetch)."""
        with torch.cuda.stream(copy_stream):
            x = img_h.to(device=dev, dtype=torch.float32, non_blocking=True, memory_format=torch.channels_last)
            t = msk_h.to(device=dev, dtype=torch.long, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x, t, ev

    def e2e_steps(n):
        """n end-to-end steps: every step's inputs cross PCIe inside the loop (the copy of step i+1 overlaps the
        compute of step i, as train.py's DataLoader(pin_memory=True) + non_blocking copies allow) and every
        step ends with the D2H read of its loss (train.py:163)."""
        nxt = prefetch()
        out = 0.0
        for i in range(n):
            x, t, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            x.record_stream(torch.cuda.current_stream())
            t.record_stream(torch.cuda.current_stream())
            if i + 1 < n:
                nxt = prefetch()
            out = step(x, t).item()
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    first_loss = None
    for _ in range(args.warmup):
        lw = step(img_d, msk_d)
        if first_loss is None:
            first_loss = float(lw.detach())
        del lw      # a live loss tensor keeps its autograd graph (and default-stream AccumulateGrad nodes) alive,
        #             which breaks the CUDA-graph capture below (cudaErrorStreamCaptureImplicit)
    sync_all()
    # ---- whole-step CUDA graph (falls back to eager launches if capture is not possible) -------
    eager_step = step
    graphed, graph_note = None, "off"
    l0 = ops.LAUNCHES
    eager_step(img_d, msk_d)             # (result dropped at once, see above)
    launches_per_step = ops.LAUNCHES - l0
    graphed_b = None
    if args.graph != "off":
        try:
            from unetb200.graph import GraphedStep
            if world == 1:
                graphed = GraphedStep(eager_step, (img_d, msk_d), warmup=2)
            else:
                # the NCCL all-reduce stays outside the captured regions:
                #   graph A = forward + loss + backward + gradients packed into the buckets
                #   eager   = mean all-reduce of the buckets
                #   graph B = clip_grad_norm_ + RMSprop step on the bucket views
                # no autograd hooks in this mode: a hook keeps the parameter's AccumulateGrad node (created on the
                # default stream by the eager warm-up) alive, and autograd would then make the legacy stream depend
                # on the capturing stream (cudaErrorStreamCaptureImplicit)
                reducer.manual = True
                reducer.remove()

                def part_a(x, t):
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", enabled=True):
                        loss = UL.training_criterion(model(x), t, boundary_coeff=0.2, edge_width=51, edge_weight=7)
                    loss.backward()
                    reducer.pack_all()
                    return loss

                def part_b():
                    clip_and_step()
                    return None

                graphed = GraphedStep(part_a, (img_d, msk_d), warmup=2)
                reducer.allreduce_all()
                reducer.point_grads()
                graphed_b = GraphedStep(part_b, (), warmup=1)
            graph_note = "on" if world == 1 else "on (backward graph | eager NCCL all-reduce | optimizer graph)"
        except Exception as exc:  # noqa: BLE001
            if args.graph == "on":
                raise
            import traceback
            sys.stderr.write(f"[rank {rank}] CUDA graph capture failed:\n" + traceback.format_exc())
            graphed, graph_note = None, f"capture failed, eager launches: {type(exc).__name__}: {exc}"[:200]
            graphed_b = None
            if reducer is not None:
                reducer.remove()
                reducer = ddp.GradAllReducer(model, bucket_bytes=bucket_mb << 20)       # back to hook-driven eager mode
            torch.cuda.synchronize()
        if world > 1:                                     # every rank must take the same path
            flag = torch.tensor([1 if graphed is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if flag.item() == 0 and graphed is not None:
                graphed = graphed_b = None
                reducer.remove()
                reducer = ddp.GradAllReducer(model, bucket_bytes=bucket_mb << 20)
    if graphed is not None:
        def step(x, t):                                   # noqa: F811  (x, t are already in the static buffers when equal)
            if x is not graphed.static_inputs[0]:
                graphed.load(x, t)
            loss = graphed.replay()
            if graphed_b is not None:
                reducer.allreduce_all()
                graphed_b.replay()
            return loss
        img_d, msk_d = graphed.static_inputs
        for _ in range(2):
            step(img_d, msk_d)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: step(img_d, msk_d), args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else {}
    e2e_steps(2)
    ms_e2e = timed(lambda: e2e_steps(args.steps), 1)
    last_loss = float(step(img_d, msk_d).detach())
    step = eager_step                                     # the instrumented pass below times individual launches
    if reducer is not None and reducer.manual:
        reducer.remove()
        reducer = ddp.GradAllReducer(model, bucket_bytes=bucket_mb << 20)

    total_imgs = B * world * args.steps
    value = total_imgs / (ms / 1e3)
    e2e_value = total_imgs / (ms_e2e / 1e3)
    peaks = measured_peaks()
    gf_img = GFLOP_PER_IMG.get((args.bilinear, S))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"UNet(1,2,bilinear={args.bilinear}) bf16 training step, batch {B}/GPU, {S}x{S}, "
                               "CE+dice+0.2*boundary_loss(51,7), clip_grad_norm, RMSprop "
                               "(BASELINE.json configs[1]; configs[3] when n_gpus > 1)",
                   "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": "activations per step (~10 GB) far exceed the 126 MB L2; no flush needed",
                   "random_init": True},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": img_h.numel() * 4 + msk_h.numel() * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "cuda_graph": graph_note,
        "clocks": clocks,
        "first_loss": first_loss, "final_loss": last_loss,
        # the step really trains: same synthetic batch every step, so the loss must not blow up (a scheduling bug
        # that feeds the optimizer stale gradients shows here)
        "loss_sane": bool(first_loss is not None and last_loss == last_loss and last_loss < 1.2 * first_loss),
    }
    if gf_img:
        conv_tf = gf_img * 1e9 * value / world / 1e12
        line["step_conv_tflops_per_gpu"] = conv_tf
        line["step_frac_of_bf16_peak"] = conv_tf / peaks["tflops_sustained"]

    # ---- per-kernel roofline: instrumented pass (CUDA events around every C-ABI call) ----------
    if world > 1 and not args.no_profile:
        # every rank must take part in the instrumented steps (they all-reduce); only rank 0 records
        if rank != 0:
            for _ in range(2):
                step(img_d, msk_d)
            torch.cuda.synchronize()
    if rank == 0 and not args.no_profile:
        os.environ["UNETB200_PROFILE_SHAPES"] = "1"
        with ops.profile() as rec:
            for _ in range(2):
                step(img_d, msk_d)
        torch.cuda.synchronize()
        os.environ.pop("UNETB200_PROFILE_SHAPES", None)
        per_layer = ops.summarize_profile(rec)
        summ = {}
        for name, d in per_layer.items():                 # class = name without the [M=..,N=..,K=..] tag
            cls = name.split("[")[0]
            e = summ.setdefault(cls, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
            for k in e:
                e[k] += d[k]
        if args.per_layer:
            line["layers"] = {n: {"ms": d["ms"] / 2, "tflops": d["flops"] / (d["ms"] * 1e-3) / 1e12}
                              for n, d in sorted(per_layer.items(), key=lambda kv: -kv[1]["ms"]) if "[" in n}
        tot_ms = sum(d["ms"] for d in summ.values()) or 1.0
        kernels = {}
        for name, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
            ent = {"ms_per_step": d["ms"] / 2, "share": d["ms"] / tot_ms, "launches_per_step": d["calls"] // 2}
            if d["flops"] > 0:
                ent["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
                ent["frac_of_peak"] = ent["tflops"] / peaks["tflops_sustained"]
            elif d["bytes"] > 0:
                ent["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9
                ent["frac_of_peak"] = ent["gbs"] / peaks["hbm_gbs"]
            kernels[name] = ent
        line["kernels"] = kernels
        dom = next(iter(kernels))
        kd, sd = kernels[dom], summ[dom]
        traffic, traffic_src = None, None
        try:                                  # DRAM bytes per launch of that kernel class from the committed ncu pass
            import glob
            tj = sorted(glob.glob(os.path.join(ROOT, "profiles", "kernel_traffic_r*.json")))[-1]
            with open(tj) as f:
                tdata = json.load(f)
            ent_t = tdata["per_class"].get(dom)
            if ent_t is None and dom in ("conv_fprop_tc", "conv_dgrad_tc"):
                ent_t = tdata["per_class"].get("conv_fprop_tc+conv_dgrad_tc")
            if ent_t:
                traffic, traffic_src = ent_t["dram_bytes_per_launch"], os.path.basename(tj) + ": " + tdata["source"]
        except Exception:  # noqa: BLE001
            pass
        if "tflops" in kd:
            line["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": kd["tflops"],
                                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": kd["frac_of_peak"],
                                "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu)", "traffic_source": traffic_src,
                                "algorithmic_flops_per_launch": sd["flops"] / sd["calls"],
                                "peak_source": peaks["source"] + " (sustained bf16 cuBLAS)",
                                "launches": sd["calls"], "avg_launch_ms": sd["ms"] / sd["calls"]}
        else:
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kd.get("gbs"), "peak": peaks["hbm_gbs"],
                                "unit": "GB/s", "frac": kd.get("frac_of_peak"), "traffic": traffic,
                                "traffic_unit": "DRAM bytes per launch (ncu)", "traffic_source": traffic_src,
                                "peak_source": peaks["source"], "launches": sd["calls"],
                                "avg_launch_ms": sd["ms"] / sd["calls"]}
        hbm = [(n, k) for n, k in kernels.items() if "gbs" in k]
        if hbm:
            hn, hk = hbm[0]
            line["roofline_hbm"] = {"bound": "hbm", "kernel": hn, "achieved": hk["gbs"], "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": hk["frac_of_peak"], "peak_source": peaks["source"] + " (copy)",
                                    "launches": summ[hn]["calls"], "avg_launch_ms": summ[hn]["ms"] / summ[hn]["calls"]}
        conv_ms = sum(d["ms"] for n, d in summ.items() if n.startswith("conv_")) / 2
        conv_fl = sum(d["flops"] for n, d in summ.items() if n.startswith("conv_")) / 2
        if conv_ms > 0:
            line["conv_tensor_util"] = {"tflops": conv_fl / (conv_ms * 1e-3) / 1e12,
                                        "frac_of_sustained_peak": conv_fl / (conv_ms * 1e-3) / 1e12 / peaks["tflops_sustained"],
                                        "frac_of_burst_peak": conv_fl / (conv_ms * 1e-3) / 1e12 / peaks["tflops_burst"],
                                        "conv_ms_per_step": conv_ms}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n, nw, med = cpu_reference_rate(S, args.bilinear, steps=3, warmup=1, budget_s=40.0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"oracle port (torch-CPU fp32) of the same step on B=1 {S}x{S} samples, "
                                          f"{nw} warm-up + {n} timed, median {med:.2f} s/step"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
