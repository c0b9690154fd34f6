#!/bin/bash
# chain kernel: lean MMA / producer loops -- parity, phase trace, stage times
echo "== pytest network/pipeline"; timeout 900 python -m pytest tests/test_network_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -4
ANCSH_CHAIN_TRACE=gpurun_out/r2u_ctrace timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 2 --chunks 2 2>&1 | tail -1 | cut -c1-120
echo "== bench forward" ; timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms'])"
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2u_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], d['roofline']['stage_ms'])"
