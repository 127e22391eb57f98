#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for v in adof_4_2 adof_4_4 adof_4_1 adof_8_2 adof_8_8 adof_16_4; do
[ -f scratch/ab/$v/libppk.so ] || continue
PPK_LIB=$PWD/scratch/ab/$v/libppk.so timeout 200 python bench.py --workload adof --steps 1000 --warmup 20 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
for w in tilt a3 a4 align nes adof; do timeout 200 python bench.py --workload $w --steps 2000 --warmup 20 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('current', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done
