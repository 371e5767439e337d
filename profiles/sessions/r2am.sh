#!/usr/bin/env bash
# GPU session r2am: three-term split mode fp16x3 (weights hi + lo too, un-fused tcgen05 plan): parity, tolerance survey, rate
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "three_term or model_variants" 2>&1 | tail -5
timeout 600 python profiles/tolerance_survey.py 1024 | tee $O/r2am_tolerance_survey.jsonl | cut -c1-200
for p in fp16x3 fp32; do timeout 300 python profiles/prof_forward.py --precision $p --chunks 2 2>&1 | tail -1 | cut -c1-400; done
timeout 120 python profiles/prof_forward.py --precision fp16x2 --chunks 2 --opt fused_blocks=0 --opt fused_front=0 2>&1 | tail -1 | cut -c1-400
