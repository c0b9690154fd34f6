#!/bin/bash
# round 2 final validation: full GPU suite, smoke, driver-form bench (ours + reference arm), the other BASELINE workloads
TAG=r02; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1800 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench (driver form)" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/${TAG}_bench.err | tail -1 | tee $OUT/${TAG}_bench.json | cut -c1-400
echo "== bench --impl reference (driver form)" ; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | tee $OUT/${TAG}_bench_reference.json | cut -c1-400
echo "== bench forward" ; timeout 600 python bench.py --workload forward --no-cpu-baseline 2>/dev/null | tail -1 | tee $OUT/${TAG}_bench_forward.json | cut -c1-300
echo "== bench drawer" ; timeout 600 python bench.py --workload drawer --no-cpu-baseline 2>/dev/null | tail -1 | tee $OUT/${TAG}_bench_drawer.json | cut -c1-300
echo "== bench cpu64" ; timeout 900 python bench.py --workload cpu64 2>/dev/null | tail -1 | tee $OUT/${TAG}_bench_cpu64.json | cut -c1-300
echo "== bench reference settings (nsample 64, 10000 hyp)" ; timeout 900 python bench.py --nsample 64 --hyp 10000 --no-cpu-baseline --steps 4 --chunks 6 2>/dev/null | tail -1 | tee $OUT/${TAG}_bench_ref_settings.json | cut -c1-300
echo "== timeline" ; timeout 300 python scripts/timeline_probe.py 2>&1 | tail -24 | tee $OUT/${TAG}_timeline.txt | tail -6
