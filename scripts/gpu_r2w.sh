#!/bin/bash
# LM phase durations / item counts with the narrow LM grid
ANCSH_LM_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 1 --warmup 3 --chunks 1 2>&1 | grep "lm trace" | tail -3
ANCSH_LM_TRACE=1 ANCSH_LM_LANE_PCT=100 timeout 300 python bench.py --no-cpu-baseline --steps 1 --warmup 3 --chunks 1 2>&1 | grep "lm trace" | tail -2
