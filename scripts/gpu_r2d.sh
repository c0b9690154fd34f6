#!/bin/bash
# round 2, call d: state = lean SA kernels (fused ball query) + ADVICE fixes.  Full GPU tests, smoke, bench, timeline.
TAG=r2d; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/${TAG}_bench.log
echo "== bench ACC1" ; ANCSH_LEAN_ACC1=1 timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_acc1.log
echo "== bench separate ball" ; ANCSH_BALL_FUSED_OFF=1 timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ballsep.log
echo "== timeline" ; timeout 300 python scripts/timeline_probe.py 2>&1 | tail -30 | tee $OUT/${TAG}_timeline.log
