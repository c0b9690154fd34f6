#!/bin/bash
# round 2, call i: pipelined lean SA kernel (ball query in the MMA shadow) + fused FP (three_nn/interp in the gather, heads in the epilogue)
TAG=r2i; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest ops+network+pipeline" ; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench forward" ; timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms'], d['roofline']['forward_tflops'])"
echo "== bench forward, separate ball" ; ANCSH_BALL_FUSED_OFF=1 timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms'])"
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], d['roofline']['stage_ms'])"
echo "== lean trace" ; ANCSH_LEAN_TRACE=$OUT/${TAG}_trace timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 2 > /dev/null 2>&1; ls $OUT | grep ${TAG}_trace
