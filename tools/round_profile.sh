#!/bin/bash
# Round-end evidence run on ONE B200 (gpurun): GPU tests, smoke, the headline bench line, the ncu launch list of the
# same command and one `ncu --set full` capture of the dominant conv kernel. Outputs land in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r01.log 2>&1; echo bench rc=$?
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:igemm_rows_kernel -c 6 -f -o gpurun_out/prof_r01_rows \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rows.log 2>&1; echo full rc=$?
