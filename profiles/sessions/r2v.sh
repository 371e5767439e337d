#!/usr/bin/env bash
# GPU session r2v: split-capable fused front end (front2_kernel<true,true>): parity tests, then timing against the un-fused front
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -15
P="python profiles/prof_forward.py --chunks 4"
: > $O/r2v_front.jsonl
timeout 120 $P --precision fp16x2 >> $O/r2v_front.jsonl 2>&1
timeout 120 $P --precision fp16x2 --opt fused_front=0 >> $O/r2v_front.jsonl 2>&1
timeout 120 $P --precision fp16 >> $O/r2v_front.jsonl 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2v_front.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln); print(d['precision'], d['opts'], {k: v for k, v in d.items() if k.endswith('_ms_per_chunk')}, 'chunk', round(d['ms_per_chunk'], 2), round(d['samples_per_s']))
    else: print(ln[:300])
PY
