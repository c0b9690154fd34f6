#!/bin/bash
# One gpurun call: parity tests, smoke, goldens from the reference kernels, bench, ncu launch list + full capture.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== golden" ; timeout 300 python tests/golden/make_ops_golden.py $OUT/ops_ref_b200.npz 2>&1 | tail -3
echo "== bench" ; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/${TAG}_bench.log
echo "== bench reference-faithful nsample=64" ; timeout 600 python bench.py --nsample 64 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ns64.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
echo "== ncu full (dominant kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_kernel -s 6 -c 2 -f -o $OUT/${TAG}_prof_sa \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 64 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
