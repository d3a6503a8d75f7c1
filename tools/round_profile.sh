#!/bin/bash
# Round-end evidence run on ONE B200 (gpurun -- bash tools/round_profile.sh): GPU tests, smoke, measured parity at
# BASELINE shapes, the headline bench line (+ cuDNN leg + CPU arm), the ncu launch list and tensor-pipe pass of the same
# command, `ncu --set full` captures of the dominant kernels in situ, the other BASELINE.json configs and the inference
# sweep. Outputs land in gpurun_out/ as r02_*; they are copied to profiles/ by hand.
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/r02_pytest_gpu.log; tail -2 $O/r02_pytest_gpu.log
python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; tail -3 $O/r02_smoke.log
python tools/parity_report.py > $O/r02_parity_report.txt 2>&1; echo parity rc=$?
python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo bench rc=$?
python bench.py --impl reference --steps 8 --warmup 2 > $O/r02_bench_reference_arm.json 2>/dev/null; echo ref rc=$?
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn"
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none \
    -c 6000 --csv --log-file $O/r02_ncu_launches_step.csv $CMD > $O/ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:'igemm_rows_kernel|wgrad_taps' --launch-skip 200 -c 10 -f \
    -o $O/r02_prof_gemm $CMD > $O/ncu_gemm.log 2>&1; echo full gemm rc=$?
ncu --set full --clock-control none --import-source on -k regex:'in_stream|in_bwd_reduce|in_act_fwd' --launch-skip 150 -c 14 -f \
    -o $O/r02_prof_tail $CMD > $O/ncu_tail.log 2>&1; echo full tail rc=$?
: > $O/r02_other_configs.txt
for args in "--gen BCDUNet --batch 64" "--gen UNet --batch 32" "--gen UNet --batch 4" "--version 1" "--batch 64"; do
  echo "== $args" >> $O/r02_other_configs.txt
  python bench.py $args --steps 6 --warmup 3 --no-cpu --no-cudnn 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('%.1f img/s %.2f ms %.1f TF/s e2e %.1f' % (d['value'], d['ms_per_step'], d['step_tflops'], d['e2e']['value']),
      {k:(round(d[k]['achieved'],1), round(d[k]['frac'],3), round(d[k]['share_of_step'],3)) for k in d if k.startswith('roofline')})" >> $O/r02_other_configs.txt
done
cat $O/r02_other_configs.txt
bash tools/infer_sweep.sh > $O/r02_infer_sweep.txt 2>&1; cat $O/r02_infer_sweep.txt
