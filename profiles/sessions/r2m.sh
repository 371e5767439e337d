#!/usr/bin/env bash
# GPU session r2m (8 GPUs): scaling rehearsal of bench.py as the driver launches it
set -u
O=gpurun_out
nvidia-smi -L | wc -l
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2m_bench_n8.json 2> $O/r2m_bench_n8.err
tail -c 300 $O/r2m_bench_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_bench_n8.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'dtype')}, 'e2e', d['e2e']['value'], 'extract', d['extract']['sites_per_s'],
      'e2e_reads', d['e2e_reads']['reads_total_per_step'], d['e2e_reads']['reads_per_s'], d['e2e_reads']['sites_per_s'], 'parity', d['parity']['max_abs_dp'], d['clocks'])
PY
