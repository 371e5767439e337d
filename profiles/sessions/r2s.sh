#!/usr/bin/env bash
# GPU session r2s: L2 prefetch of the next tile's x boxes in kernel A's split-mode TMA producer (variant build -DDSB_A_SPLIT_PREFETCH=1)
set -u
O=gpurun_out
P="python profiles/prof_forward.py --precision fp16x2 --chunks 4"
: > $O/r2s_prefetch.jsonl
for i in 1 2; do
  timeout 120 $P >> $O/r2s_prefetch.jsonl 2>&1
  DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_apf.so timeout 120 $P >> $O/r2s_prefetch.jsonl 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2s_prefetch.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln); print(d['precision'], d.get('lib'), 'A', d.get('incept_a_ms_per_chunk'), 'B', d.get('incept_b_ms_per_chunk'), 'chunk', round(d['ms_per_chunk'], 2))
PY
DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_apf.so timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -3
