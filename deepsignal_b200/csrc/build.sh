#!/usr/bin/env bash
# Build libdsb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libdsb200.so
NVCC=${NVCC:-nvcc}
FLAGS="${EXTRA_NVCC:-} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p build
pids=()
for f in api fold gemm_f32 gemm_tc incept_tc lstm_seq2 front2_tc forward extract; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer build/$f.o)" ] || [ ../../include/dsb200.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
if [ ! -f build/tsv.o ] || [ tsv.cpp -nt build/tsv.o ] || [ ../../include/dsb200.h -nt build/tsv.o ]; then
  g++ -O3 -std=c++17 -fPIC -fvisibility=hidden -c tsv.cpp -o build/tsv.o &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/api.o build/fold.o build/gemm_f32.o build/gemm_tc.o build/incept_tc.o build/lstm_seq2.o build/front2_tc.o build/forward.o build/extract.o build/tsv.o -cudart static
echo "built $(realpath $OUT)"
