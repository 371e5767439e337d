#!/usr/bin/env bash
# GPU session r2ag: site_feature_kernel A/B on one box: committed version (old) against the one with warp-cooperative long slices
set -u
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do
echo old; DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_old.so timeout 120 python profiles/prof_extract.py | cut -c100-330
echo new; timeout 120 python profiles/prof_extract.py | cut -c100-330
done
