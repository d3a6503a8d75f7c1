mkdir -p gpurun_out
for p2p in 1 0 1 0; do
TG_P2P=$p2p python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2971$p2p bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/bench8_p2p$p2p.err | grep "^{" > gpurun_out/r02_bench_8gpu_p2p$p2p.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_8gpu_p2p$p2p.json")); print("p2p=$p2p", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,3) for k,v in d["comm"]["wait_ms_per_step_min_over_ranks"].items()}, {k: round(v,3) for k,v in d["comm"]["wait_ms_per_step_max_over_ranks"].items()})
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 tools/p2p_bench.py 2>&1 | grep "^world" | tee gpurun_out/p2p_bench_8gpu.txt
