#!/usr/bin/env bash
# GPU session r2an: full GPU suite with the fp16x3 mode and the stride-1 pool kernel; rates of fp16x3 / un-fused fp16x2; smoke; bench line
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python profiles/prof_forward.py --precision fp16x3 --chunks 2 2>&1 | tail -1 | cut -c1-400
timeout 900 python bench.py > $O/r2an_bench_n1.json 2> $O/r2an_bench_n1.err; tail -c 200 $O/r2an_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2an_bench_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), d['e2e']['value'], d['clocks']['sm_mhz'], {k:v for k,v in d['parity'].items() if k in ('max_abs_dp','identical_calls','meets_tolerance')})
print({k:{a:b for a,b in v.items() if a in ('max_abs_dp','identical_calls','samples_per_s_through_host_api')} for k,v in d['parity']['other_modes'].items()})
PY
