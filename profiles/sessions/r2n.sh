#!/usr/bin/env bash
# GPU session r2n: rehearsal of the driver's round-end sequence (tests, smoke, reference arm, bench) with its flags
set -u
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/r2n_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r2n_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $O/r2n_ref.json 2>/dev/null ) 2>&1 | grep real
cut -c1-400 $O/r2n_ref.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/r2n_bench.json 2> $O/r2n_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/r2n_ref.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ref', r['value'], 'ratio e2e/ref', d['e2e']['value'] / r['value'])
print('feature_file', d['feature_file']['samples_per_s'], 'parity', d['parity']['max_abs_dp'], d['parity']['identical_calls'], d['parity']['meets_tolerance'])
print('extract', d['extract']['roofline']['frac'], d['extract']['per_kernel_ms'])
PY
