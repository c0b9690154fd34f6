#!/bin/bash
# LM 384-thread blocks (168 registers, some spills) vs 256-thread blocks
run() { echo "== $*"; env "$@" timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stage_ms']; print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], {k:v for k,v in s.items() if k.startswith('pose_joint')})"; }
echo "== pytest pose"; timeout 900 env ANCSH_LM_THREADS=384 python -m pytest tests/test_pose_gpu.py -q -x --tb=short 2>&1 | tail -3
run ANCSH_LM_THREADS=384
run ANCSH_LM_THREADS=384 ANCSH_LM_LANE_PCT=75
run ANCSH_LM_THREADS=256
