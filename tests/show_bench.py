"""Pretty-print a bench.py JSON line: python tests/show_bench.py gpurun_out/bench.log [nlayers]"""
import json
import sys

line = [x for x in open(sys.argv[1]) if x.startswith("{")][-1]
d = json.loads(line)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print("value", round(d["value"], 1), d["unit"], " ms/step", round(d["ms_per_step"], 2), " e2e", round(d["e2e"]["value"], 1),
      " launches", d.get("gpu_launches"), " clocks", d.get("clocks"))
print("conv", d.get("conv_tensor_util"), " step frac", d.get("step_frac_of_bf16_peak"))
for k, v in d.get("kernels", {}).items():
    perf = f"{v['tflops']:.0f} TF" if "tflops" in v else (f"{v['gbs']:.0f} GB/s" if "gbs" in v else "")
    print(f"{k:28s} {v['ms_per_step']:8.3f} ms  {v['share'] * 100:5.1f}%  n={v['launches_per_step']:3d}  {perf}")
for k, v in list(d.get("layers", {}).items())[:n]:
    print(f"{k:60s} {v['ms']:7.3f} ms {v['tflops']:7.0f} TF")
