#!/bin/bash
# branch-free FPS: bit-exactness + stage time
echo "== pytest ops/network"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -q -x --tb=short 2>&1 | tail -4
echo "== bench forward" ; timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms'])"
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'])"
