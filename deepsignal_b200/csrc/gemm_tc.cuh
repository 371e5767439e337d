// gemm_tc.cuh — interface of the tcgen05 / TMA bf16 GEMM (sm_100a only).
#pragma once
#include "ops.cuh"

namespace dsb {

struct SegTC {            // one A segment: K blocks read through tensor map `map`
  int map;                // index into TcMaps::a
  int col0;               // first column (elements) in the mapped matrix
  int drow;               // row shift (conv tap - pad)
  int nkb;                // number of BK-wide K blocks
  int wkb;                // first BK-wide block of the weight matrix this segment multiplies (the hi and lo parts of an
                          // activation share their weight blocks in the split-operand mode)
};

struct alignas(64) TcMaps {
  CUtensorMap a[2];
  CUtensorMap b;
};

struct GemmTCParams {
  int nseg = 0;
  SegTC seg[MAX_SEG];
  int BN = 0;             // N tile (multiple of 16, <= 256, divides epi.N)
  int stages = 0;
  int resident = 0;       // 1: the CTA's whole B tile [BN x K] is loaded once and stays in smem (conv weights)
  int wkb_total = 0;      // BK-wide blocks of the weight matrix (= sum of nkb unless segments share weight blocks)
  EpiParams epi;
};

struct TcLaunch {         // everything needed to launch one prebuilt GEMM
  TcMaps maps;
  GemmTCParams p;
  int BK = 64;            // 64 -> SWIZZLE_128B tiles, 32 -> SWIZZLE_64B tiles
  int epi_kind = EPI_CONV;
  bool fp16 = false;       // operand / output element type: false = bf16, true = fp16 (both kind::f16)
  size_t smem = 0;
  double flops_per_row = 0;
};

// Encode a 2-D bf16/fp16 row-major [rows][cols] (leading dimension ld elements) tensor map with a
// [box_rows][box_cols] box; box_cols*2 bytes must equal the swizzle span (128 or 64).
void encode_map_16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                   int box_rows, int box_cols, bool fp16);
void finalize_tc_launch(TcLaunch& L);                       // picks stages / smem
void launch_gemm_tc(const TcLaunch& L, int M, int sm_count, int ablate, cudaStream_t s);

}  // namespace dsb
