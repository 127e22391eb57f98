#!/bin/bash
for v in a4_8_4 a4_4_2 a4_8_1; do
PPK_LIB=$PWD/scratch/ab/$v/libppk.so timeout 200 python bench.py --workload a4 --steps 2000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
for v in tilt_16_4 tilt_16_2 tilt_8_2 tilt_8_4 tilt_16_1 tilt_32_2; do
for w in tilt tilt_1m; do
PPK_LIB=$PWD/scratch/ab/$v/libppk.so timeout 200 python bench.py --workload $w --steps 1000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done; done
