#!/bin/bash
# round 2, call h: new bench.py (workloads, dominant-stage roofline, launch counter, per-batch seeds) + full GPU tests
TAG=r2h; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.log
echo "== bench (driver form)" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/${TAG}_bench.err | tail -1 | tee $OUT/${TAG}_bench.json | cut -c1-3000
tail -5 $OUT/${TAG}_bench.err
echo "== bench forward" ; timeout 600 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | tee $OUT/${TAG}_bench_forward.json | cut -c1-1500
echo "== bench cpu64" ; timeout 900 python bench.py --workload cpu64 --steps 20 2>&1 | tail -1 | tee $OUT/${TAG}_bench_cpu64.json | cut -c1-1500
echo "== bench drawer" ; timeout 600 python bench.py --workload drawer --no-cpu-baseline --steps 10 2>&1 | tail -1 | tee $OUT/${TAG}_bench_drawer.json | cut -c1-1500
echo "== bench mixed (1 GPU, 12500 clouds)" ; timeout 900 python bench.py --workload mixed --clouds 12500 --steps 4 2>&1 | tail -1 | tee $OUT/${TAG}_bench_mixed.json | cut -c1-1500
