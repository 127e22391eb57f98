#!/bin/bash
run() {
  for w in tilt a3 a4 align adof tilt_1m; do timeout 200 python bench.py --workload $w --steps 2000 --warmup 20 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
}
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
run pdl
export PPK_NO_PDL=1
run nopdl
