#!/bin/bash
echo "== pytest pose/pipeline/stream"; timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py tests/test_stream_gpu.py tests/test_pose_f32_reference.py -q -x --tb=short 2>&1 | tail -5
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], 'host enqueue', d['config'].get('host_enqueue_ms_per_step'), {k:v for k,v in d['roofline']['stage_ms'].items() if k.startswith('pose')})"
