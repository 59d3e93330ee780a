#!/usr/bin/env python
"""Side-by-side per-kernel-class table of bench.py JSON lines:  python tools/cmp_bench.py a.json b.json ..."""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `| head` is a normal way to read this table

def load(path):
    lines = [ln for ln in open(path) if ln.lstrip().startswith("{")]
    return json.loads(lines[-1])


runs = [load(p) for p in sys.argv[1:]]
print("value     ", "  ".join(f"{r['value']:9.1f}" for r in runs), " img/s")
print("ms/step   ", "  ".join(f"{r['ms_per_step']:9.3f}" for r in runs))
print("sm_mhz    ", "  ".join(f"{r.get('clocks', {}).get('sm_mhz', 0):9.0f}" for r in runs))
print("launches  ", "  ".join(f"{r.get('gpu_launches', 0) / max(r['steps'], 1):9.0f}" for r in runs))
names = []
for r in runs:
    for k in r.get("kernels", {}):
        if k not in names:
            names.append(k)
names.sort(key=lambda k: -max(r.get("kernels", {}).get(k, {}).get("ms_per_step", 0) for r in runs))
for k in names:
    cells = []
    for r in runs:
        v = r.get("kernels", {}).get(k)
        cells.append(f"{v['ms_per_step']:7.3f}/{v['launches_per_step']:<3d}{v.get('frac_of_peak', 0):5.2f}" if v else " " * 15)
    print(f"{k:26s}", "  ".join(cells))
tot = [sum(v["ms_per_step"] for v in r.get("kernels", {}).values()) for r in runs]
print(f"{'sum of event-timed':26s}", "  ".join(f"{t:15.3f}" for t in tot))
