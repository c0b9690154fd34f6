#!/bin/bash
echo "== pytest ops/network/pose/pipeline"; timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py tests/test_pose_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -5
echo "== LM trace"; ANCSH_LM_TRACE=1 timeout 300 python bench.py --steps 1 --chunks 2 --no-cpu-baseline 2>&1 | grep "lm trace" | sed -n 6,7p
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2m_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], 'host enqueue first', d['config'].get('host_enqueue_ms_first_step'), d['roofline']['stage_ms'])"
