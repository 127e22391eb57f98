#!/bin/bash
for v in a4_16_2 a4_16_4 a4_16_1 a4_8_2; do
PPK_LIB=$PWD/scratch/ab/$v/libppk.so timeout 200 python bench.py --workload a4 --steps 2000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
