#!/usr/bin/env bash
# GPU session r2q: final-state evidence: launch lists (forward fp16x2, extraction) and --set full captures of the split kernels
set -u
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r2q_launches.csv python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2q_launches_extract.csv python profiles/prof_extract.py > $O/r2q_extract.json 2>/dev/null
timeout 120 python profiles/prof_extract.py | tee $O/r2q_extract.json
for k in incept_a_kernel incept_b_kernel pool3s2_16_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 24 -c 1 -o $O/r2q_full_$k python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pool3s2_16_kernel -s 4 -c 1 -f -o $O/r2q_full_pool3s2_16_kernel python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 1 -o $O/r2q_full_conv3 python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:site_feature_kernel -s 2 -c 1 -o $O/r2q_full_site_feature python profiles/prof_extract.py > /dev/null 2>&1
ls -la $O | grep r2q
