#!/bin/bash
# 8 GPUs: BASELINE configs[4] (100k mixed stream) as one pass and as 4 gathered segments
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
echo "== mixed 100k on 8 GPUs, 1 segment"; timeout 900 $TR bench.py --gpus 8 --workload mixed --clouds 100000 --steps 1 --warmup 2 2>$OUT/r2_n8_mixed1.err | tail -1 | tee $OUT/r2_n8_mixed1.json | cut -c1-300
echo "== mixed 100k on 8 GPUs, 4 segments"; timeout 900 $TR bench.py --gpus 8 --workload mixed --clouds 100000 --steps 4 --warmup 2 2>$OUT/r2_n8_mixed.err | tail -1 | tee $OUT/r2_n8_mixed.json | cut -c1-300
