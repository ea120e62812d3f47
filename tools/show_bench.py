"""Pretty-print a bench.py JSON line.  Usage: python tools/show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1),
      "launches", d["gpu_launches"])
print(d["stage_ms_rank0"], "occ", round(d["mean_batch_occupancy"], 1), "steps/s", round(d["steps_per_s_rank0"], 1))
print("profiled_step_ms", d["roofline"]["profiled_step_ms"], "roofline", {k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "frac")})
for k, v in d["kernel_classes"].items():
    print(f"  {k:24s} {v}")
for k, v in list(d.get("gemm_shapes", {}).items())[:28]:
    print(f"  {k:36s} {v}  us/launch={1e3 * v['ms'] / v['launches']:.1f}")
print(d.get("latency"))
