#!/usr/bin/env bash
# GPU session r2p: kernel B with U tiles decoupled from the accumulator stages (variants), correctness of the default build
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "fused_inception or default_tensor_core or repeated or chunking" 2>&1 | tail -3
P="python profiles/prof_forward.py --precision fp16x2 --chunks 4"
: > $O/r2p_variants.jsonl
timeout 120 $P >> $O/r2p_variants.jsonl 2>&1
for v in u2s10g1 u2s10g2 u2s12g1; do
  DSB_LIB=$PWD/deepsignal_b200/variants/libdsb200_$v.so timeout 120 $P >> $O/r2p_variants.jsonl 2>&1
done
timeout 120 $P >> $O/r2p_variants.jsonl 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2p_variants.jsonl'):
    if ln.startswith('{'):
        d = json.loads(ln); print(round(d['ms_per_chunk'], 2), d.get('incept_a_ms_per_chunk'), d.get('incept_b_ms_per_chunk'))
    else:
        print(ln[:200])
PY
