#!/usr/bin/env bash
# GPU session r2ae: --set full capture of the reworked site_feature_kernel (stall samples per instruction)
set -u
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:site_feature_kernel -s 2 -c 1 -f -o $O/r2ae_full_site_feature python profiles/prof_extract.py > /dev/null 2>&1
ls -la $O | grep r2ae
