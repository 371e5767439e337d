#!/usr/bin/env bash
# GPU session r2w: where the MMA thread of the split front kernel waits (instrumented variant build, block 0 prints)
set -u
DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_fdiag.so timeout 120 python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 2>&1 | tail -3
timeout 120 python profiles/prof_forward.py --precision fp16x2 --chunks 4 2>&1 | tail -1
timeout 120 python profiles/prof_forward.py --precision fp16 --chunks 4 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -3
