#!/usr/bin/env bash
# GPU session r2t: kernel A dynamic MMA issue order; upper bound of what the single staging buffer costs (racy build: timing only)
set -u
O=gpurun_out
P="python profiles/prof_forward.py --precision fp16x2 --chunks 4"
: > $O/r2t_variants.jsonl
for v in "" adyn racy adynracy; do
  if [ -z "$v" ]; then timeout 120 $P >> $O/r2t_variants.jsonl 2>&1
  else DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_$v.so timeout 120 $P >> $O/r2t_variants.jsonl 2>&1; fi
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2t_variants.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln); print(d['precision'], d.get('lib'), 'A', d.get('incept_a_ms_per_chunk'), 'B', d.get('incept_b_ms_per_chunk'), 'chunk', round(d['ms_per_chunk'], 2))
    else: print(ln[:200])
PY
DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_adyn.so timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -3
