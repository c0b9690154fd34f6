#!/bin/bash
# documented run-time switches still work: LM block shapes, separate ball query, generic chain for layer1/2
for E in ANCSH_LM_THREADS=64 ANCSH_LM_THREADS=128 ANCSH_BALL_FUSED_OFF=1 ANCSH_SA_LEAN_OFF=1 ANCSH_FPS_PREFIX_OFF=1; do
echo "== $E"; env $E timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_network_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short -k "not reference_default_hypothesis and not bench_batch" 2>&1 | tail -2
done
