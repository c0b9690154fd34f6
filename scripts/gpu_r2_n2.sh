#!/bin/bash
# 2 GPUs: driver-form launch of both arms
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
echo "== ours, 2 GPUs"; timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/r2_n2_full.json | cut -c1-260
echo "== reference arm, 2 ranks"; timeout 900 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-260
