#!/bin/bash
# 8 GPUs of one box: BASELINE configs[4] (100k mixed stream, strong scaling, timed gather) and the default workload (weak scaling)
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
echo "== mixed 100k on 8 GPUs"; timeout 900 $TR bench.py --gpus 8 --workload mixed --clouds 100000 --steps 4 --warmup 2 2>$OUT/r2_n8_mixed.err | tail -1 | tee $OUT/r2_n8_mixed.json | cut -c1-1800
tail -3 $OUT/r2_n8_mixed.err
echo "== full (weak scaling) on 8 GPUs"; timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/r2_n8_full.err | tail -1 | tee $OUT/r2_n8_full.json | cut -c1-2500
tail -3 $OUT/r2_n8_full.err
echo "== full on 1 GPU of the same box"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee $OUT/r2_n8_full_n1.json | cut -c1-600
