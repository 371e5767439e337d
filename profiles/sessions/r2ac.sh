#!/usr/bin/env bash
# GPU session r2ac (2 GPUs): the product's --devices 0,1 path on two real GPUs, and the bench line at N=2
set -u
O=gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/r2ac_bench_n2.json 2> $O/r2ac_bench_n2.err; tail -c 300 $O/r2ac_bench_n2.err; head -c 250 $O/r2ac_bench_n2.json; echo
