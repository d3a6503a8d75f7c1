#!/bin/bash
# generic ABAB of environment switches on one box: bash tools/ab_env.sh "A=1" "A=0" ... (each variant run in the given order)
cd "$(dirname "$0")/.."
for v in "$@"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],2), 'tail', round(d['roofline_tail']['frac'],3), d['clocks']['sm_mhz'])"
done
