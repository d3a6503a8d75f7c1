#!/bin/bash
# A/B of the round-2 scheduling changes on one box: serpentine order, deferred weight gradient, slim InstanceNorm
# backward beside the weight-gradient GEMM. Usage (GPU box): bash tools/ab_slim.sh
# The "wt4" variants need a second build of the library with shallower weight-gradient pipelines (more shared memory
# left for the slim ring), made here before the call:
#   mkdir -p ab && (cd tactile_gan_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 \
#     -Xcompiler -fPIC -shared -DTG_WT_STAGES=4 -DTG_WG256_STAGES=3 -o ../../ab/libtg_wt4.so tg_api.cu tg_tail.cu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/ab_slim.txt
: > $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -3 >> $O
echo "== tail_bench default policy" >> $O
timeout 120 python tools/tail_bench.py >> $O 2>&1
echo "== tail_bench ring everywhere (TG_STREAM=2)" >> $O
TG_STREAM=2 timeout 120 python tools/tail_bench.py >> $O 2>&1
echo "== tail_bench slim alone" >> $O
timeout 120 python tools/tail_bench.py --slim >> $O 2>&1
echo "== slim_probe (default lib, 32 KiB)" >> $O
timeout 120 python tools/slim_probe.py >> $O 2>&1
echo "== slim_probe (wgrad 4/3 stages, 72 KiB)" >> $O
TG_LIB_PATH=$PWD/ab/libtg_wt4.so TG_SLIM_KB=72 timeout 120 python tools/slim_probe.py >> $O 2>&1
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
run base        TG_SLIM=0 TG_WGRAD_DEFER=0 TG_SERP=0
run ring_all    TG_SLIM=0 TG_WGRAD_DEFER=0 TG_SERP=0 TG_STREAM=2
run defer_slim  TG_SLIM=1 TG_WGRAD_DEFER=1 TG_SERP=0 TG_SLIM_MIN_RATIO=1.0
run wt4_slim    TG_SLIM=1 TG_WGRAD_DEFER=1 TG_SERP=0 TG_SLIM_MIN_RATIO=1.0 TG_LIB_PATH=$PWD/ab/libtg_wt4.so TG_SLIM_KB=72
run base2       TG_SLIM=0 TG_WGRAD_DEFER=0 TG_SERP=0
cat $O
