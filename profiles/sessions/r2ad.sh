#!/usr/bin/env bash
# GPU session r2ad: site_feature_kernel with the per-site window staged in shared memory: parity tests, timing
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 120 python profiles/prof_extract.py | tee $O/r2ad_extract.json; done
