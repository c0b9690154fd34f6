#!/bin/bash
# round-2 quick check of the lean SA kernel: parity, smoke, forward bench with stage times; A/B against the generic kernel
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest network" ; timeout 600 python -m pytest tests/test_network_gpu.py -q -x --tb=short 2>&1 | tail -25 | tee $OUT/r2a_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/r2a_smoke.log
echo "== bench forward (lean)" ; timeout 600 python bench.py --stages forward --no-cpu-baseline --steps 20 2>&1 | tail -1 | tee $OUT/r2a_bench_fwd.log
echo "== bench forward (lean, separate ball query)" ; ANCSH_BALL_FUSED_OFF=1 timeout 600 python bench.py --stages forward --no-cpu-baseline --steps 20 2>&1 | tail -1 | tee $OUT/r2a_bench_fwd_nofuse.log
echo "== bench forward (generic)" ; ANCSH_SA_LEAN_OFF=1 timeout 600 python bench.py --stages forward --no-cpu-baseline --steps 20 2>&1 | tail -1 | tee $OUT/r2a_bench_fwd_generic.log
