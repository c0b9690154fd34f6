#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list + full capture.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag] [quick|noprof]
TAG=${1:-r01}
MODE=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/${TAG}_bench.log
echo "== bench forward-only" ; timeout 600 python bench.py --stages forward --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_fwd.log
if [ "$MODE" == "all" ]; then
echo "== bench reference-faithful (nsample 64, 10000 hyp)" ; timeout 900 python bench.py --nsample 64 --hyp 10000 --no-cpu-baseline --steps 5 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ref_settings.log
echo "== bench --impl reference" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/${TAG}_bench_reference.log
fi
if [ "$MODE" != "noprof" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
echo "== ncu full (dominant kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain2_kernel|gemm_tc_kernel|single_score|joint_lm_kernel|joint_init_kernel|joint_refit|fps_kernel' -s 30 -c 16 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 64 > $OUT/${TAG}_ncu_full.log 2>&1
echo "== ncu dram traffic of the chain kernels at the bench batch (256)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"chain2_kernel" -s 24 -c 6 --csv --log-file $OUT/${TAG}_chain_traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_traffic.log 2>&1
fi
if [ "$MODE" == "all" ]; then
echo "== pipelined timeline" ; timeout 300 python scripts/timeline_probe.py 2>&1 | tail -30 | tee $OUT/${TAG}_timeline.log
fi
ls -la $OUT | tail -20
