#!/bin/bash
# BASELINE.json configs[4]: test.py inference sweep on one B200 (generator forward only), batch 1..512 at 256^2 and 512^2.
# usage (on the GPU box): bash tools/infer_sweep.sh > gpurun_out/infer_sweep.txt
for args in "--batch 1" "--batch 8" "--batch 64" "--batch 256" "--batch 512" \
            "--size 512 --batch 1" "--size 512 --batch 16" "--size 512 --batch 128"; do
  python bench.py --workload infer $args --steps 4 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
r=d.get('roofline',{}); t=d.get('roofline_tail',{})
print('%-26s %8.1f img/s %8.2f ms  %6.1f TFLOP/s  e2e %8.1f img/s  conv %.0f TF/s (%.2f)  tail %.0f GB/s (%.2f)' % (
  '$args', d['value'], d['ms_per_step'], d['step_tflops'], d['e2e']['value'], r.get('achieved',0), r.get('frac',0), t.get('achieved',0), t.get('frac',0)))"
done
