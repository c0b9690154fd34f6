#!/bin/bash
# FPS block shape sweep (n = 1024, m = 512, 256 clouds) with the branch-free update
for S in 32 64 128 256; do
echo "== shape $S" ; ANCSH_FPS_SHAPE=$S timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 6 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms']['fps1'])"
done
