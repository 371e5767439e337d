#!/usr/bin/env bash
# GPU session r2h: lo8 (8-bit lo parts between the fused split kernels): correctness + timing
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -s 2>&1 | grep -E "fp16x2|passed|failed|Error|error" | head -30 | tee $O/r2h_tests.log
P="python profiles/prof_forward.py --precision fp16x2 --chunks 4"
: > $O/r2h_prof.jsonl
timeout 120 $P >> $O/r2h_prof.jsonl 2>&1
timeout 120 $P >> $O/r2h_prof.jsonl 2>&1
cut -c1-420 $O/r2h_prof.jsonl
