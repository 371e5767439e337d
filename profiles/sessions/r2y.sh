#!/usr/bin/env bash
# GPU session r2y: state after the split fused front end: full GPU test suite, smoke, launch list + front-kernel capture, bench line
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r2y_launches.csv python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:front2_kernel -s 2 -c 1 -f -o $O/r2y_full_front2 python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 > /dev/null 2>&1
timeout 900 python bench.py > $O/r2y_bench.json 2> $O/r2y_bench.err; tail -c 600 $O/r2y_bench.err; head -c 400 $O/r2y_bench.json
ls -la $O | grep r2y
