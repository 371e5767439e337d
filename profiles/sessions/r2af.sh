#!/usr/bin/env bash
# GPU session r2af: site_feature_kernel at 6 / 8 blocks per SM (40 / 32 registers) against the default 4 (63 registers)
set -u
for v in occ6 occ8; do echo $v; DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_$v.so timeout 120 python profiles/prof_extract.py; done
echo default; timeout 120 python profiles/prof_extract.py
