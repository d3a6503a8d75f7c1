#!/bin/bash
# Round-end evidence run on ONE B200 (gpurun): GPU tests, smoke, the headline bench line (+ reference arm), the ncu
# launch list of the same command, one `ncu --set full` capture of the dominant conv kernel, and the other
# BASELINE.json configs. Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r01.log 2>&1; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo ref rc=$?
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:igemm_rows_kernel -c 8 -f -o gpurun_out/prof_r01_rows \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rows.log 2>&1; echo full rc=$?
: > gpurun_out/other_configs.txt
for args in "--gen BCDUNet --batch 64" "--gen UNet --batch 32" "--version 1"; do
  echo "== $args" >> gpurun_out/other_configs.txt
  python bench.py $args --steps 4 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('%.1f img/s %.2f ms %.1f TF/s e2e %.1f' % (d['value'], d['ms_per_step'], d['step_tflops'], d['e2e']['value']),
      {k:(round(d[k]['achieved'],1), round(d[k]['frac'],3), round(d[k]['share_of_step'],3)) for k in d if k.startswith('roofline')})" >> gpurun_out/other_configs.txt
done
cat gpurun_out/other_configs.txt
