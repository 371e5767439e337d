#!/usr/bin/env bash
# GPU session r2x: kernel A split mode -- does ring depth matter (4 slots instead of 5), and what do the b1 / lo-part p1 MMAs cost
set -u
O=gpurun_out
P="python profiles/prof_forward.py --precision fp16x2 --chunks 4"
: > $O/r2x_a.jsonl
for opt in "ablate_a=0" "ablate_a=8" "ablate_a=16" "ablate_a=24" "ablate_a=28" "ablate_a=29"; do timeout 120 $P --opt $opt >> $O/r2x_a.jsonl 2>&1; done
DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_a4.so timeout 120 $P >> $O/r2x_a.jsonl 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2x_a.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln); print(d['precision'], d.get('lib'), d['opts'], 'A', d.get('incept_a_ms_per_chunk'), 'B', d.get('incept_b_ms_per_chunk'), 'front', d.get('front_ms_per_chunk'), 'chunk', round(d['ms_per_chunk'], 2))
    else: print(ln[:200])
PY
