#!/usr/bin/env bash
# GPU session r2e (2 GPUs): torchrun bench at N=2, the --devices CLI test on two real devices, torchrun call_mods
set -u
O=gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "devices" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2e_bench_n2.json 2> $O/r2e_bench_n2.err
tail -c 400 $O/r2e_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2e_bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'dtype')}, d['e2e']['value'], d['extract']['sites_per_s'], d['e2e_reads']['reads_per_s'], d['e2e_reads']['sites_per_s'], d['parity']['max_abs_dp'])
PY
# torchrun call_mods (process sharding): rank 0 writes; compare with the single-process file
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_reads
from deepsignal_b200.extract_features import save_reads
W.save_npz('/tmp/m.ckpt.npz', W.random_checkpoint(17, 'stress'))
save_reads('/tmp/reads.npz', synth_reads(400, seed=45, nb_lo=300, nb_hi=700, cg_frac=0.03))
PY
timeout 300 python -m deepsignal_b200.deepsignal call_mods -i /tmp/reads.npz -m /tmp/m.ckpt.npz -o /tmp/one.tsv --f5_batch_num 1 > /dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m deepsignal_b200.deepsignal call_mods -i /tmp/reads.npz -m /tmp/m.ckpt.npz -o /tmp/two.tsv --f5_batch_num 1 > /dev/null 2> $O/r2e_torchrun_cli.err
cmp /tmp/one.tsv /tmp/two.tsv && echo "torchrun call_mods: byte-identical ($(wc -l < /tmp/one.tsv) lines)" | tee $O/r2e_torchrun_cli.txt
