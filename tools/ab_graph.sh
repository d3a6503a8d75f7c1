#!/bin/bash
# whole-step CUDA graph at the headline batch: TG_STEP_GRAPH / TG_PDL variants on one box
cd "$(dirname "$0")/.."
for v in "TG_STEP_GRAPH=0" "TG_STEP_GRAPH=1" "TG_STEP_GRAPH=1 TG_PDL=2" "TG_STEP_GRAPH=0" "TG_STEP_GRAPH=1" "TG_STEP_GRAPH=1 TG_PDL=2"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
