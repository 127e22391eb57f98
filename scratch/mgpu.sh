#!/bin/bash
# 2-GPU check of the sharded bench + reference arm
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4000 --warmup 20 > gpurun_out/bench_2gpu.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/bench_2gpu.log | cut -c1-900
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
