#!/usr/bin/env bash
# Build libdsb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=${DSB_OUT:-../libdsb200.so}        # DSB_OUT / DSB_BUILD_DIR: variant builds for A/B timing (loaded through DSB_LIB)
BD=${DSB_BUILD_DIR:-build}
NVCC=${NVCC:-nvcc}
FLAGS="${EXTRA_NVCC:-} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p $BD
pids=()
for f in api fold gemm_f32 gemm_tc incept_tc lstm_seq2 front2_tc forward extract sitefreq; do
  if [ ! -f $BD/$f.o ] || [ $f.cu -nt $BD/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer $BD/$f.o)" ] || [ ../../include/dsb200.h -nt $BD/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o $BD/$f.o &
    pids+=($!)
  fi
done
if [ ! -f $BD/tsv.o ] || [ tsv.cpp -nt $BD/tsv.o ] || [ ../../include/dsb200.h -nt $BD/tsv.o ]; then
  g++ -O3 -std=c++17 -fPIC -fvisibility=hidden -c tsv.cpp -o $BD/tsv.o &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $BD/api.o $BD/fold.o $BD/gemm_f32.o $BD/gemm_tc.o $BD/incept_tc.o $BD/lstm_seq2.o $BD/front2_tc.o $BD/forward.o $BD/extract.o $BD/sitefreq.o $BD/tsv.o -cudart static
echo "built $(realpath $OUT)"
