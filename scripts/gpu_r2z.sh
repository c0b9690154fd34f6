#!/bin/bash
# mixed stream: categories enqueued back to back (run_many_begin / finish); 1 GPU carrying one rank's share of configs[4]
echo "== pytest stream/pipeline"; timeout 900 python -m pytest tests/test_stream_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -3
for S in 4 1; do
echo "== mixed 12500 clouds, $S segment(s)"; timeout 600 python bench.py --workload mixed --clouds 12500 --steps $S --warmup 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['config'].get('seconds_per_rank'), d['config'].get('gather_ms_total'))"
done
