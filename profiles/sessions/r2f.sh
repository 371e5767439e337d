#!/usr/bin/env bash
# GPU session r2f: new pool kernel (tests + timing), kernel-B build variants
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -4
P="python profiles/prof_forward.py --precision fp16x2 --chunks 4"
: > $O/r2f_variants.jsonl
timeout 120 $P >> $O/r2f_variants.jsonl 2>&1
for v in b3x8s1 b2x12s1; do
  DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_$v.so timeout 120 $P >> $O/r2f_variants.jsonl 2>&1
done
timeout 120 $P >> $O/r2f_variants.jsonl 2>&1
timeout 120 python profiles/prof_forward.py --precision fp16 --chunks 4 >> $O/r2f_variants.jsonl 2>&1
cut -c1-420 $O/r2f_variants.jsonl
