#!/bin/bash
for s in 4 6 8; do
echo "== bench full ANCSH_SLOTS=$s" ; ANCSH_SLOTS=$s timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'])"
done
