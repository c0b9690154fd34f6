#!/bin/bash
# quick GPU iteration: network parity + forward bench (v2 and, for comparison, ANCSH_TC_V1=1), all under timeouts
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest network+pipeline" ; timeout 600 python -m pytest tests/test_network_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== bench forward" ; timeout 600 python bench.py --stages forward --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_fwd.log
if [ "$2" == "v1" ]; then
echo "== bench forward (v1 kernels)" ; ANCSH_TC_V1=1 timeout 600 python bench.py --stages forward --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_fwd_v1.log
fi
if [ "$2" == "full" ] || [ "$3" == "full" ]; then
echo "== bench full" ; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/${TAG}_bench.log
fi
