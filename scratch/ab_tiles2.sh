#!/bin/bash
for v in a4_4_1 a4_12_1 a4_12_2 a4_8_1; do
[ -f scratch/ab/$v/libppk.so ] || continue
PPK_LIB=$PWD/scratch/ab/$v/libppk.so timeout 200 python bench.py --workload a4 --steps 2000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
for v in tilt_12_2 tilt_20_2 tilt_12_1 tilt_8_1 tilt_24_2 tilt_24_4 tilt_16_2; do
[ -f scratch/ab/$v/libppk.so ] || continue
for w in tilt tilt_1m; do
PPK_LIB=$PWD/scratch/ab/$v/libppk.so timeout 200 python bench.py --workload $w --steps 1000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done; done
