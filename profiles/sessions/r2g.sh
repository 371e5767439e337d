#!/usr/bin/env bash
# GPU session r2g: full validation of the round-2 state + the remaining evidence (extraction launch list, batch sweep)
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2g_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/r2g_smoke.log
timeout 900 python bench.py --steps 5 > $O/r2g_bench.json 2> $O/r2g_bench.err; tail -c 300 $O/r2g_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/r2g_launches_extract.csv python profiles/prof_extract.py > $O/r2g_extract.json 2>/dev/null
timeout 120 python profiles/prof_extract.py | tee -a $O/r2g_extract.json
timeout 300 python profiles/batch_sweep.py fp16x2 > $O/r2g_sweep.jsonl 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file $O/r2g_sweep_ncu.csv python profiles/batch_sweep.py fp16x2 --once > /dev/null 2>&1
python profiles/batch_sweep.py --merge $O/r2g_sweep_ncu.csv $O/r2g_sweep.jsonl $O/r2g_batch_sweep_fp16x2.jsonl | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'dtype')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['parity']['max_abs_dp'], d['parity']['identical_calls'], d['cpu_baseline']['value'])
PY
