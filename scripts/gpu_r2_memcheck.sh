#!/bin/bash
# compute-sanitizer memcheck over the GPU tests that exercise the kernels changed in round 2
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py tests/test_dataset_gpu.py tests/test_pipeline_gpu.py tests/test_pose_gpu.py tests/test_stream_gpu.py -q -x \
    -k "not reference_default_hypothesis and not bench_batch" 2>&1 | tail -25 | tee $OUT/r02b_memcheck.txt
echo "exit: ${PIPESTATUS[0]}" | tee -a $OUT/r02b_memcheck.txt
