#!/usr/bin/env python
"""Turn Nsight Compute output into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
      per-kernel totals of one profiled training step (`ncu --metrics gpu__time_duration.sum --csv`)
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_full.md
      the roofline-relevant raw metrics of every launch in a `--set full` report (needs `ncu` on PATH)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("ub::", "").replace("at::native::", "")
    return name[-90:]


def launches(path):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        e = agg.setdefault(short(r[ik]), [0, 0.0])
        e[0] += 1
        e[1] += float(r[iv].replace(",", "")) / 1e3
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# Launch list of one training step (UNet(1,2) bf16, B=16, 512x512): {n} launches, {tot / 1e3:.2f} ms of kernel time")
    print("# (ncu --metrics gpu__time_duration.sum --clock-control none: serialised, cold-cache -- compare SHARES)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")


def full(path):
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none: raw metrics per launch ({path.split('/')[-1]})\n")
    for r in rows[2:]:
        print(f"## `{short(r[hdr.index('Kernel Name')])}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print(f"| **traffic = dram read + write** | {(rd * mul[ur] + wr * mul[uw]) / 1e6:.1f} | MB |")
        except (ValueError, KeyError):
            pass
        print()


def metrics(path):
    """Long-format CSV of a few metrics for every launch of a step -> per-kernel table + per-launch list."""
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    iid, ik, im, iu, iv, ig = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size"))
    per = collections.OrderedDict()
    mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0, "cycle": 1.0}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        e = per.setdefault(r[iid], {"name": short(r[ik]), "grid": r[ig]})
        e[r[im]] = float(r[iv].replace(",", "")) * mul.get(r[iu], 1.0)
    T, RD, WR, TP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    agg = collections.OrderedDict()
    for e in per.values():
        a = agg.setdefault(e["name"], dict(n=0, us=0.0, rd=0.0, wr=0.0, tp=0.0))
        a["n"] += 1
        a["us"] += e.get(T, 0.0)
        a["rd"] += e.get(RD, 0.0)
        a["wr"] += e.get(WR, 0.0)
        a["tp"] += e.get(TP, 0.0) * e.get(T, 0.0)
    tot = sum(a["us"] for a in agg.values())
    print(f"# One training step (UNet(1,2) bf16, B=16, 512x512) under ncu: {len(per)} launches, {tot / 1e3:.2f} ms of kernel time")
    print("# ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum,sm__pipe_tensor_cycles_active... --clock-control none")
    print("# (serialised, cold-cache replays: compare SHARES; traffic = DRAM bytes read + written, summed over the launches)\n")
    print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | traffic MB / launch | achieved DRAM GB/s | tensor pipe active % (time-weighted) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        print(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | "
              f"{(a['rd'] + a['wr']) / 1e6 / a['n']:.1f} | {gbs:.0f} | {a['tp'] / a['us'] if a['us'] else 0:.1f} |")
    print("\n## Every launch, in order\n")
    print("| # | kernel | grid | us | DRAM read MB | DRAM write MB | tensor % |\n|---:|---|---|---:|---:|---:|---:|")
    for i, e in per.items():
        print(f"| {i} | `{e['name'][:70]}` | {e['grid']} | {e.get(T, 0):.1f} | {e.get(RD, 0) / 1e6:.1f} | {e.get(WR, 0) / 1e6:.1f} | {e.get(TP, 0):.1f} |")


if __name__ == "__main__":
    {"launches": launches, "full": full, "metrics": metrics}[sys.argv[1]](sys.argv[2])
