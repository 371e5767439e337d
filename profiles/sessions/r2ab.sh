#!/usr/bin/env bash
# GPU session r2ab: memcheck of every device path after the front-kernel work; bench line (N=1) of this state
set -u
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize.py fp16x2 fp16 > $O/r2ab_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/r2ab_memcheck.log
timeout 900 python bench.py > $O/r2ab_bench_n1.json 2> $O/r2ab_bench_n1.err; tail -c 300 $O/r2ab_bench_n1.err; head -c 300 $O/r2ab_bench_n1.json; echo
