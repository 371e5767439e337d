#!/usr/bin/env bash
# GPU session r2aj: kernel A split mode with branch 1 on the pooled hi tile only: parity tests, timing, 100k-sample tolerance (bench parity block)
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do timeout 120 python profiles/prof_forward.py --precision fp16x2 --chunks 4 2>&1 | tail -1 | cut -c1-520; done
timeout 900 python bench.py --no-cpu-baseline > $O/r2aj_bench.json 2> $O/r2aj_bench.err; tail -c 200 $O/r2aj_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['e2e']['value'], d['clocks']['sm_mhz'], d['parity'])
PY
