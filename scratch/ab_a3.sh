#!/bin/bash
for v in a3_8_1 a3_8_2 a3_4_1 a3_16_4 current; do
L=$PWD/scratch/ab/$v/libppk.so; [ "$v" = current ] && L=$PWD/isaacgym_b200/_lib/libppk.so
PPK_LIB=$L timeout 200 python bench.py --workload a3 --steps 4000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
