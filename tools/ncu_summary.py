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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
