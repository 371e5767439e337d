#!/usr/bin/env bash
# GPU session r2al: last full check of the committed state: GPU test suite, smoke, extraction profile + launch list
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python profiles/prof_extract.py | tee $O/r2al_extract.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2al_launches_extract.csv python profiles/prof_extract.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:site_feature_kernel -s 2 -c 1 -f -o $O/r2al_full_site_feature python profiles/prof_extract.py > /dev/null 2>&1
ls $O | grep r2al
