#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest ops+network" ; timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -q -x --tb=short 2>&1 | tail -15 | tee $OUT/r2b_pytest.log
echo "== trace (fused ball)" ; ANCSH_LEAN_TRACE=$OUT/r2b_trace timeout 300 python bench.py --stages forward --no-cpu-baseline --steps 3 2>&1 | tail -1 | cut -c1-300
echo "== trace (separate ball)" ; ANCSH_BALL_FUSED_OFF=1 ANCSH_LEAN_TRACE=$OUT/r2b_trace_nofuse timeout 300 python bench.py --stages forward --no-cpu-baseline --steps 3 2>&1 | tail -1 | cut -c1-300
ls -la $OUT/r2b*
