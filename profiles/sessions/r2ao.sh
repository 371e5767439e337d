#!/usr/bin/env bash
# GPU session r2ao: ncu launch list of the fp16x3 plan (one 37,888-sample chunk)
set -u
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2ao_launches_fp16x3.csv python profiles/prof_forward.py --precision fp16x3 --chunks 1 --reps 1 > /dev/null 2>&1
ls -la $O | grep r2ao
