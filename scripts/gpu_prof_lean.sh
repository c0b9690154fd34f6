#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sa_lean_kernel' -s 4 -c 2 -f -o $OUT/r2a_lean \
    python bench.py --stages forward --steps 1 --warmup 3 --no-cpu-baseline --batch 128 > $OUT/r2a_ncu.log 2>&1
tail -3 $OUT/r2a_ncu.log
ls -la $OUT/*.ncu-rep
