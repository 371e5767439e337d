#!/usr/bin/env bash
# GPU session r2r: where kernel A's tile period goes in split mode (epilogue vs MMA/pool chain)
set -u
O=gpurun_out
P="python profiles/prof_forward.py --precision fp16x2 --chunks 2"
: > $O/r2r_ablate_a.jsonl
for opt in "ablate_a=0" "ablate_a=4" "ablate_a=5" "ablate_a=2" "ablate_a=1"; do timeout 120 $P --opt $opt >> $O/r2r_ablate_a.jsonl 2>&1; done
timeout 120 python profiles/prof_forward.py --precision fp16 --chunks 2 --opt ablate_a=0 >> $O/r2r_ablate_a.jsonl 2>&1
timeout 120 python profiles/prof_forward.py --precision fp16 --chunks 2 --opt ablate_a=4 >> $O/r2r_ablate_a.jsonl 2>&1
timeout 120 python profiles/prof_forward.py --precision fp16 --chunks 2 --opt ablate_a=5 >> $O/r2r_ablate_a.jsonl 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2r_ablate_a.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln); print(d['precision'], d['opts'], 'A', d.get('incept_a_ms_per_chunk'), 'chunk', round(d['ms_per_chunk'], 2))
PY
