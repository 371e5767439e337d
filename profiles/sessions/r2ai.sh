#!/usr/bin/env bash
# GPU session r2ai: pairwise_block with the loads of its sequential parts batched
set -u
timeout 900 python -m pytest tests/test_gpu_extract.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do timeout 120 python profiles/prof_extract.py | cut -c100-330; done
