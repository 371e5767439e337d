#!/usr/bin/env bash
# GPU session r2aa: where the MMA threads of the split kernels A and B wait (instrumented variant build, block 0 prints per launch)
set -u
DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_abdiag.so timeout 120 python profiles/prof_forward.py --precision fp16x2 --chunks 1 --reps 1 2>&1 | grep diag | tail -22
