#!/bin/bash
# round 2, call e: lane-group LM kernel -- parity tests per mode, phase timings (ANCSH_LM_TRACE), bench per mode
OUT=gpurun_out; mkdir -p $OUT
for mode in 2 1; do
echo "== pytest pose/pipeline/stream mode $mode"; ANCSH_LM_MODE=$mode timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py tests/test_stream_gpu.py -q -x --tb=short 2>&1 | tail -8
done
for cfg in "0 1 4" "0 4 4" "1 1 4" "1 1 2" "1 1 1" "2 1 4" "2 2 4" "2 4 4" "2 1 2"; do
set -- $cfg
echo "== LM trace mode=$1 serial_blocks=$2 group_blocks=$3"
ANCSH_LM_TRACE=1 ANCSH_LM_MODE=$1 ANCSH_LM_BLOCKS_PER_SM=$2 ANCSH_LM_GROUP_BLOCKS_PER_SM=$3 timeout 300 python bench.py --steps 2 --no-cpu-baseline 2>&1 | grep "lm trace" | sed -n 4,5p
done
for cfg in "0 1 4" "1 1 4" "1 1 2" "2 1 4" "2 2 4" "2 1 2"; do
set -- $cfg
echo "== bench mode=$1 serial_blocks=$2 group_blocks=$3"
ANCSH_BALL_FUSED_OFF=1 ANCSH_LM_MODE=$1 ANCSH_LM_BLOCKS_PER_SM=$2 ANCSH_LM_GROUP_BLOCKS_PER_SM=$3 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],2), 'serial', d['config']['serialized_ms_per_step'], 'joint', d['roofline']['stage_ms']['pose_joint_score'], d['config']['joint_lm'])"
done
