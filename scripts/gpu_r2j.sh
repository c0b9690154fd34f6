#!/bin/bash
# round 2, call j: full GPU suite after the weight-image scaling + new parity tests; bench
TAG=r2j; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1800 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], d['roofline']['stage_ms'])"
