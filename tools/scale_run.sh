#!/bin/bash
# Multi-GPU measurements on one 8-GPU box (gpurun --gpus 8 -- bash tools/scale_run.sh): BASELINE.json configs[2] / [4].
#   1. weak scaling, 32 images per GPU (what the driver's SCALE run measures): N = 1, 8
#   2. configs[2] exactly: global batch 256 sharded over 2 / 4 / 8 GPUs (128 / 64 / 32 per GPU) -- strong scaling
#   3. N = 8 with NCCL limited to 8 CTAs (do the ring CTAs compete with the 148-CTA persistent GEMM grids?)
#   4. configs[4] on 8 GPUs: generator-only inference, one replica per GPU
mkdir -p gpurun_out
run() {  # N batch tag [env...]
  local n=$1 b=$2 tag=$3; shift 3
  if [ "$n" = 1 ]; then
    env "$@" python bench.py --gpus 1 --batch $b --steps 10 --warmup 3 --no-cpu --no-cudnn 2>/dev/null | grep '^{' > gpurun_out/r02_scale_$tag.json
  else
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --batch $b --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_scale_$tag.json
  fi
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_scale_$tag.json"))
c = d.get("comm") or {}
print("$tag: N=%d batch/GPU=%d  %.1f img/s  %.2f ms/step  e2e %.1f  sm %s MHz  comm wait min/max over ranks %s / %s" % (
    d["n_gpus"], $b, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"],
    {k: round(v, 3) for k, v in (c.get("wait_ms_per_step_min_over_ranks") or {}).items()},
    {k: round(v, 3) for k, v in (c.get("wait_ms_per_step_max_over_ranks") or {}).items()}))
PY
}
if [ "$1" = "n8only" ]; then run 8 32 weak_n8_b; exit 0; fi
run 1 32 weak_n1
run 8 32 weak_n8
run 8 32 weak_n8_nccl8cta NCCL_MAX_CTAS=8
run 2 128 strong256_n2
run 4 64 strong256_n4
for args in "--batch 64" "--batch 512" "--size 512 --batch 16"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29790 \
    bench.py --gpus 8 --workload infer $args --steps 4 --warmup 3 --no-cpu 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('infer 8 GPUs %-24s %9.1f img/s %8.2f ms  e2e %9.1f img/s' % ('$args', d['value'], d['ms_per_step'], d['e2e']['value']))"
done
