#!/bin/bash
for st in 3000 6000; do
echo "== parity stagger $st"; ANCSH_CHAIN_STAGGER_PARITY=1 ANCSH_CHAIN_STAGGER=$st timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stage_ms']; print(round(d['value']), {k:s[k] for k in ('sa3','fp1','fp2','fp3_heads')})"
done
ANCSH_CHAIN_STAGGER_PARITY=1 ANCSH_CHAIN_STAGGER=4000 ANCSH_CHAIN_TRACE=gpurun_out/r2r_ctrace timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 2 --chunks 2 2>&1 | tail -1 | cut -c1-100
