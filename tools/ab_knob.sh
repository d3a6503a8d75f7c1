#!/bin/bash
# cache-hint experiments on the InstanceNorm passes (tg_debug_knob bits), micro-benchmark + step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/ab_knob.txt
: > $O
for k in 0 1 2 3 4; do
  timeout 120 python tools/tail_bench.py --knob $k >> $O 2>&1
done
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" >> $O <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:28s} value {d['value']:7.1f}  e2e {d['e2e']['value']:7.1f}  ms {d['ms_per_step']:6.2f}  tail {d['roofline_tail']['frac']:.3f}  conv {d['roofline']['frac']:.3f}  wgrad {d['roofline_wgrad']['frac']:.3f}  mhz {d['clocks']['sm_mhz']}")
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/ab_{name}.err").read()[-600:])
PY
}
run knob0 TG_KNOB=0
run knob1 TG_KNOB=1
run knob5 TG_KNOB=5
run knob7 TG_KNOB=7
run knob0b TG_KNOB=0
cat $O
