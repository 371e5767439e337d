#!/usr/bin/env bash
# GPU session r2z: front kernel with the pool-validity penalty columns and single-pass split pool staging
set -u
timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 120 python profiles/prof_forward.py --precision fp16x2 --chunks 4 2>&1 | tail -1 | cut -c1-520
done
timeout 120 python profiles/prof_forward.py --precision fp16 --chunks 4 2>&1 | tail -1 | cut -c1-520
timeout 120 python profiles/prof_forward.py --precision bf16 --chunks 4 2>&1 | tail -1 | cut -c1-520
