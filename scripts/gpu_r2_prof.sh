#!/bin/bash
# round 2 profiles: ncu launch list of a bench step, full captures (source-level) of the dominant kernels, DRAM traffic
TAG=r02; OUT=gpurun_out; mkdir -p $OUT
echo "== ncu launch list (2 device batches per step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --chunks 2 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
tail -2 $OUT/${TAG}_ncu_launches.log | cut -c1-300
echo "== ncu full: forward kernels (batch 256)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'sa_lean_kernel|chain2_kernel|gemm_img_kernel|fps_kernel|three_nn_kernel' -s 22 -c 12 -f -o $OUT/${TAG}_fwd \
    python bench.py --workload forward --steps 1 --warmup 3 --chunks 1 --no-cpu-baseline > $OUT/${TAG}_ncu_fwd.log 2>&1
tail -2 $OUT/${TAG}_ncu_fwd.log | cut -c1-300
# summaries are extracted on the box (gpurun copies back at most 64 MiB): raw metrics of every captured launch, SASS-level
# stall tables of the lean SA kernels and the fused FP chain
ncu -i $OUT/${TAG}_fwd.ncu-rep --page raw --csv > $OUT/${TAG}_fwd_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_fwd.ncu-rep --page source --csv -k regex:'sa_lean_kernel' > $OUT/${TAG}_fwd_src_sa.csv 2>/dev/null
ncu -i $OUT/${TAG}_fwd.ncu-rep --page source --csv -k regex:'chain2_kernel' > $OUT/${TAG}_fwd_src_chain.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/${TAG}_fwd.ncu-rep $OUT/${TAG}_kernels_fwd.txt "bench.py --workload forward, 256 clouds, one forward" > /dev/null 2>&1
python scripts/ncu_stalls.py $OUT/${TAG}_fwd_src_sa.csv sa_lean 45 > $OUT/${TAG}_sa_lean_sass_stalls.txt 2>&1
python scripts/ncu_stalls.py $OUT/${TAG}_fwd_src_chain.csv chain2 45 > $OUT/${TAG}_chain2_sass_stalls.txt 2>&1
rm -f $OUT/${TAG}_fwd.ncu-rep $OUT/${TAG}_fwd_src_sa.csv $OUT/${TAG}_fwd_src_chain.csv
echo "== ncu full: pose kernels (batch 256)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'joint_lm_kernel|joint_refit_kernel|single_score_kernel|joint_init_kernel|joint_verify_kernel|single_refit_kernel' -s 30 -c 8 -f -o $OUT/${TAG}_pose \
    python bench.py --steps 1 --warmup 3 --chunks 1 --no-cpu-baseline > $OUT/${TAG}_ncu_pose.log 2>&1
tail -2 $OUT/${TAG}_ncu_pose.log | cut -c1-300
ncu -i $OUT/${TAG}_pose.ncu-rep --page raw --csv > $OUT/${TAG}_pose_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_pose.ncu-rep --page source --csv -k regex:'joint_lm_kernel' > $OUT/${TAG}_pose_src_lm.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/${TAG}_pose.ncu-rep $OUT/${TAG}_kernels_pose.txt "bench.py (full), 256 clouds, pose stage" > /dev/null 2>&1
python scripts/ncu_stalls.py $OUT/${TAG}_pose_src_lm.csv joint_lm 60 > $OUT/${TAG}_joint_lm_sass_stalls.txt 2>&1
rm -f $OUT/${TAG}_pose.ncu-rep $OUT/${TAG}_pose_src_lm.csv
echo "== ncu dram traffic of the forward kernels at the bench batch"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:'sa_lean_kernel|chain2_kernel|gemm_img_kernel|three_nn_kernel|fps_kernel|cloud_bias' -s 33 -c 11 --csv --log-file $OUT/${TAG}_fwd_traffic.csv \
    python bench.py --workload forward --steps 1 --warmup 3 --chunks 1 --no-cpu-baseline > $OUT/${TAG}_ncu_traffic.log 2>&1
tail -2 $OUT/${TAG}_ncu_traffic.log | cut -c1-300
ls -la $OUT | grep ${TAG}_
