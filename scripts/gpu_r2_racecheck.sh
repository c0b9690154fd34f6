#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the op-level, small network and pose tests; the hazards are
# summarised per pair of source functions (the full log stays on the box)
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 1000 \
    python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py tests/test_pose_gpu.py -q -x \
    -k "not reference_default_hypothesis and not bench_batch and not stress and not wide_dynamic" > /tmp/racecheck_full.txt 2>&1
{
  echo "# compute-sanitizer --tool racecheck --racecheck-report analysis, tests/test_ops_gpu.py test_network_gpu.py test_pose_gpu.py"
  grep -E "passed|failed|RACECHECK SUMMARY" /tmp/racecheck_full.txt
  echo "# hazards by accessing function (Error / and lines):"
  grep -oE "(Write|Read) access at [^+]*" /tmp/racecheck_full.txt | sed -E 's/\(const float.*//' | sort | uniq -c | sort -rn
  echo "# kernels named in the reports:"
  grep -oE "in (net_lean|net_tc2|net_tc|ops|pose|net|metrics)\.cu:[0-9]+" /tmp/racecheck_full.txt | sort | uniq -c | sort -rn | head -40
} | tee $OUT/r02_racecheck.txt
