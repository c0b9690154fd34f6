#!/bin/bash
# single_refit with 128-thread blocks
run() { echo "== $*"; env "$@" timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stage_ms']; print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], {k:v for k,v in s.items() if k.startswith('pose')})"; }
echo "== pytest pose"; timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py tests/test_pose_f32_reference.py tests/test_stream_gpu.py -q -x --tb=short 2>&1 | tail -3
run A=1
run A=2
