"""Print the per-shape / per-kernel timing table written by `bench.py --layers`."""
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/layers.json"))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("gemm ms/step %.2f   tail ms/step %.2f" % (d["gemm_ms_per_step"], d["tail_ms_per_step"]))
agg = {}
for r in d["gemm"]:
    k = r["tag"].split()[0]
    t, f = agg.get(k, (0, 0))
    agg[k] = (t + r["ms_per_step"], f + r["gflop_per_step"])
for k, (t, f) in agg.items():
    print("  %-6s %7.2f ms  %7.0f TF/s" % (k, t, f / t))
for r in d["gemm"][:top]:
    print("%7.3f ms x%3.0f %6.0f TF/s %7.0f GF  %s" % (r["ms_per_step"], r["launches_per_step"], r["tflops"],
                                                     r["gflop_per_step"], r["tag"]))
print()
byname = {}
for r in d["tail"]:
    k = r["kernel"].split("(")[0]
    t, c = byname.get(k, (0, 0))
    byname[k] = (t + r["ms_per_step"], c + r["launches_per_step"])
for k, (t, c) in sorted(byname.items(), key=lambda kv: -kv[1][0]):
    print("%7.3f ms x%3.0f %s" % (t, c, k))
print()
for r in d["tail"][:top]:
    print("%7.3f ms x%3.0f %s" % (r["ms_per_step"], r["launches_per_step"], r["kernel"]))
