#!/usr/bin/env bash
# GPU session r2c: tests, kernel-B ablations / variants, ncu launch list + full captures of the split kernels, reference arm
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r2c_tests.log
tail -3 $O/r2c_tests.log
P="python profiles/prof_forward.py --precision fp16x2 --chunks 2"
: > $O/r2c_ablate.jsonl
for opt in "" "ablate_b=1" "ablate_b=2" "ablate_b=3" "ablate_a=1" "ablate_a=2" "ablate_a=3"; do
  timeout 120 $P ${opt:+--opt $opt} >> $O/r2c_ablate.jsonl 2>&1
done
DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_b3x6.so timeout 120 $P >> $O/r2c_ablate.jsonl 2>&1
cat $O/r2c_ablate.jsonl | cut -c1-420
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
  --csv --log-file $O/r2c_launches.csv python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:incept_a_kernel -s 24 -c 2 -o $O/r2c_full_a \
  python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:incept_b_kernel -s 24 -c 2 -o $O/r2c_full_b \
  python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 420 python bench.py --impl reference --steps 2 > $O/r2c_ref.json 2>/dev/null
cut -c1-900 $O/r2c_ref.json
ls -la $O | tail -8
