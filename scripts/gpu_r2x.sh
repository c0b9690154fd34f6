#!/bin/bash
# tiled cloud bias + partition sort fix: parity, stage times; SASS stall tables of the short pose kernels
OUT=gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_network_gpu.py tests/test_pose_gpu.py tests/test_pipeline_gpu.py -q -x --tb=short 2>&1 | tail -4
echo "== bench forward" ; timeout 300 python bench.py --workload forward --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['stage_ms'])"
echo "== bench full" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stage_ms']; print(round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], {k:v for k,v in s.items() if k.startswith('pose')})"
echo "== ncu pose short kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'partition_kernel|single_score_kernel|single_refit_kernel|joint_init_kernel|joint_verify_kernel' -s 10 -c 5 -f -o $OUT/r2x_pose \
    python bench.py --steps 1 --warmup 3 --chunks 1 --no-cpu-baseline > $OUT/r2x_ncu.log 2>&1
tail -1 $OUT/r2x_ncu.log | cut -c1-200
for k in partition single_score single_refit joint_init joint_verify; do
ncu -i $OUT/r2x_pose.ncu-rep --page source --csv -k regex:${k}_kernel > $OUT/r2x_src_$k.csv 2>/dev/null
python scripts/ncu_stalls.py $OUT/r2x_src_$k.csv $k 50 > $OUT/r2x_${k}_stalls.txt 2>&1
rm -f $OUT/r2x_src_$k.csv
done
python scripts/ncu_summary.py $OUT/r2x_pose.ncu-rep $OUT/r2x_kernels_pose_short.txt "bench.py (full), 256 clouds, short pose kernels" > /dev/null 2>&1
rm -f $OUT/r2x_pose.ncu-rep
ls -la $OUT | grep r2x
