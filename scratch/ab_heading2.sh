#!/bin/bash
run() { for w in tilt a4 adof tilt_1m; do timeout 200 python bench.py --workload $w --steps 2000 --warmup 20 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['variant'], d['config']['envs_per_gpu'], round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4))"; done; }
python scratch/err_probe.py; run faithful
export PPK_NVCC_EXTRA=-DPPK_HALF_ANGLE_HEADING
python -m isaacgym_b200.build --force > /dev/null 2>&1
python scratch/err_probe.py; run halfangle
