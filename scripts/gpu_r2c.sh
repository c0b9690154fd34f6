#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest ops+network" ; timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -q -x --tb=short 2>&1 | tail -15 | tee $OUT/r2c_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== smoke ACC1" ; ANCSH_LEAN_ACC1=1 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== precision probe (default / ACC1)"; timeout 300 python scripts/precision_probe2.py 2>&1 | grep -A8 "== f16x3" | head -9
ANCSH_LEAN_ACC1=1 timeout 300 python scripts/precision_probe2.py 2>&1 | grep -A8 "== f16x3" | head -9
echo "== bench forward" ; ANCSH_LEAN_TRACE=$OUT/r2c_trace timeout 300 python bench.py --stages forward --no-cpu-baseline --steps 20 2>&1 | tail -1 | grep -o '"stage_ms.*'
echo "== bench forward ACC1" ; ANCSH_LEAN_ACC1=1 timeout 300 python bench.py --stages forward --no-cpu-baseline --steps 20 2>&1 | tail -1 | grep -o '"stage_ms.*'
echo "== bench forward, separate ball" ; ANCSH_BALL_FUSED_OFF=1 timeout 300 python bench.py --stages forward --no-cpu-baseline --steps 20 2>&1 | tail -1 | grep -o '"stage_ms.*'
