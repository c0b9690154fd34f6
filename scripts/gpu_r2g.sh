#!/bin/bash
echo "== pytest pose/pipeline (default mode)"; ANCSH_LM_MODE=0 timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -4
echo "== pytest pose/pipeline demote 3"; ANCSH_LM_MODE=0 ANCSH_LM_DEMOTE=3 timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -4
for cfg in "0 1" "2 1" "3 1" "4 1" "6 1" "3 2" "3 4" "0 4"; do
set -- $cfg
echo "== LM trace demote=$1 serial_blocks=$2"
ANCSH_LM_TRACE=1 ANCSH_LM_MODE=0 ANCSH_LM_DEMOTE=$1 ANCSH_LM_BLOCKS_PER_SM=$2 timeout 300 python bench.py --steps 2 --no-cpu-baseline 2>&1 | grep "lm trace" | sed -n 4,5p
done
for cfg in "0 1" "3 1" "4 1" "3 2"; do
set -- $cfg
echo "== bench demote=$1 serial_blocks=$2"
ANCSH_BALL_FUSED_OFF=1 ANCSH_LM_MODE=0 ANCSH_LM_DEMOTE=$1 ANCSH_LM_BLOCKS_PER_SM=$2 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],2), 'serial', d['config']['serialized_ms_per_step'], 'joint', d['roofline']['stage_ms']['pose_joint_score'], d['config']['joint_lm'])"
done
