#!/bin/bash
run() {
  for w in a4 tilt; do timeout 200 python bench.py --workload $w --steps 2000 --warmup 20 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
}
run now
PPK_LIB=$PWD/scratch/ab/372ec20/libppk.so run before
run now2
PPK_LIB=$PWD/scratch/ab/372ec20/libppk.so run before2
