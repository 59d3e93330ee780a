#!/usr/bin/env python
"""Summarise a per-launch ncu CSV of one training step (tests/ncu_step.py under
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv`):

    python tools/ncu_step_summary.py gpurun_out/step.csv profiles/rNN_step_metrics.md profiles/kernel_traffic_rNN.json
"""
import collections
import csv
import json
import re
import sys

src, out_md, out_json = sys.argv[1:4]
rows = list(csv.reader([ln for ln in open(src) if ln.startswith('"')]))
hdr = rows[0]
ik, im, iv, iid, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"name": r[ik]})
    d[r[im]] = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1)


def short(n):
    n = re.sub(r"\(.*$", "", n).replace("void ", "").replace("ub::", "").replace("(int)", "")
    return n[-84:]


agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(short(d["name"]), dict(n=0, t=0.0, b=0.0, tp=0.0))
    t = d["gpu__time_duration.sum"]
    a["n"] += 1
    a["t"] += t
    a["b"] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    a["tp"] += d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] * t
tot = sum(a["t"] for a in agg.values())
lines = ["# One training step under ncu (UNet(1,2) bf16, B=16, 512x512, eager launches; tests/ncu_step.py)",
         f"# {len(per)} launches, {tot / 1000:.2f} ms of kernel time.  Per-launch times are serialised and cold-cache: compare SHARES;",
         "# tensor % is time-weighted over the launches of the row; DRAM = read + write.", "",
         "| kernel | launches | total us | share | tensor pipe % | DRAM MB / launch | DRAM GB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    lines.append(f"| `{k}` | {a['n']} | {a['t']:.1f} | {100 * a['t'] / tot:.1f}% | {a['tp'] / a['t'] if a['t'] else 0:.1f} | "
                 f"{a['b'] / a['n'] / 1e6:.1f} | {a['b'] / a['t'] / 1e3 if a['t'] else 0:.0f} |")
open(out_md, "w").write("\n".join(lines) + "\n")
CLS = [("conv_wgrad_tc", r"tc[234]_wgrad_kernel"), ("conv_fprop_tc+conv_dgrad_tc", r"tc3_conv_kernel"),
       ("conv_convT_fprop_tc+conv_convT_dgrad_tc", r"tc2_fprop_kernel"), ("bn_relu_bwd_apply", r"bn_relu_bwd_apply_kernel"),
       ("bn_relu_bwd_reduce", r"bn_relu_bwd_reduce_kernel"), ("bn_relu_apply_pool", r"bn_relu_apply_pool_kernel"),
       ("bn_relu_apply", r"bn_relu_apply_kernel"), ("maxpool2_bwd", r"maxpool2_bwd_kernel"), ("outconv_fwd", r"outconv_fwd"),
       ("outconv_bwd", r"outconv_bwd_vec"), ("wgrad_reduce", r"wgrad_reduce_multi"), ("rmsprop_step", r"rmsprop_step_kernel"),
       ("pack_weights", r"pack_multi_kernel"), ("conv_fprop_simt", r"first_tc_kernel|first_conv_fprop"),
       ("conv_wgrad_simt", r"first_conv_wgrad")]
out = {"source": f"{out_md} (ncu per-launch metrics over one training step, tests/ncu_step.py, B=16 512x512)", "per_class": {}}
for name, pat in CLS:
    sel = [d for d in per.values() if re.search(pat, d["name"])]
    if sel:
        b = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in sel)
        out["per_class"][name] = {"launches_per_step": len(sel), "dram_bytes_per_launch": b / len(sel), "dram_bytes_per_step": b}
json.dump(out, open(out_json, "w"), indent=1)
print("\n".join(lines[:28]))
