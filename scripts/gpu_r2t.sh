#!/bin/bash
# image-streaming conv2 GEMM (sa3): parity + stage times
echo "== pytest network/pipeline"; timeout 900 python -m pytest tests/test_network_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench forward" ; timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms'])"
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2t_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], d['wall_s_timed_region'] if 'wall_s_timed_region' in d else '', d['roofline']['stage_ms'])"
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:'gemm_img|chain2' -c 40 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --workload forward --no-cpu-baseline --steps 1 --warmup 1 --chunks 1 > /dev/null 2>&1; grep -c gemm_img gpurun_out/r2t_launches.csv
