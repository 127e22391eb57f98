#!/bin/bash
# Larger-than-config env counts per GPU: how far the fixed launch cost is amortised.
# usage (GPU box): bash tools/sweep_sizes.sh
for w in "align2 65536" "a4 262144" "a4 524288" "align2 524288" "adof 262144" "tilt 2097152" "nes 1048576"; do
  set -- $w
  timeout 300 python bench.py --workload $1 --envs-per-gpu $2 --no-extras --steps 4000 --sets 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', $2, round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],3), round(d['value']/1e9,3))"
done
