// front2_tc.cu — fused CNN front end with conv1 on the tensor cores as well (see front_tc.cuh).
//
// front_tc.cu computes conv1 + pool on CUDA cores and is issue-bound there (two CTAs per tile each spend ~10 k warp
// instructions on the 172 k FMAs of a tile).  Here conv1 is ONE [128 x 48]·[48 x 192] MMA chain per tile:
//   * A1 row i = the 11-sample signal window of pooled row i as 16-bit (hi | lo | hi) parts, x = hi + lo
//   * B1 row (a, c) = conv1 weights of channel c shifted to pool tap a, as (hi | hi | lo) parts, w = hi + lo
//   so acc1[i][a*64 + c] = sum_t (xh*wh + xl*wh + xh*wl) = conv1(2*pos - pl + a)[c] to ~2^-22 relative (the parts are fp16 in
//   both precision modes; |signal| is clamped to the fp16 range) — far below the 16-bit rounding of P1 that follows.  The pool is then lane-local:
//   P1[i][c] = relu(max_a acc1[i][a*64+c] + bias[c]) over the valid taps (+bias and ReLU are monotone, so this equals
//   max_a relu(conv + bias)), 0 for halo rows.
// conv2 / conv3 / epilogue are those of front_tc.cu.  Every buffer is single (the stages of consecutive tiles overlap as
// a pipeline: im2col(k+2) | conv1(k+2) | pool(k+1) | conv2(k+1) | C2 staging(k) | conv3(k) | epilogue(k-1)).
// Not bit-identical to stem_kernel's fp32 FMA chain (see above); P1 differs by one 16-bit ulp for ~1e-4 of the elements.
//
// 704 threads: warp 0 TMA (weights, once), warp 1 MMA issuer, warps 2-5 im2col, 6-9 pool staging, 10-13 C2 staging,
// 14-21 epilogue.
#include <cstdlib>

#include "front_tc.cuh"
#include "tc_common.cuh"

namespace dsb {
namespace {

using namespace tc;

constexpr int G_THREADS = 704;
constexpr uint32_t G_W3_BLK = 128 * 32 * 2;            // [128 x 32] SWIZZLE_64B conv3 weight block; 12 = 3 taps x 4 k-blocks
constexpr uint32_t G_W2_BYTES = 128 * 64 * 2;          // conv2 weights [128 x 64] SWIZZLE_128B
constexpr uint32_t G_W1_BYTES = 192 * 64 * 2;          // conv1 parts [192 x 64] SWIZZLE_128B (K = 48 used)
constexpr uint32_t G_A1_BYTES = 128 * 64 * 2;          // im2col tile [128 x 64] SWIZZLE_128B (K = 48 used)
constexpr uint32_t G_P1_BYTES = 128 * 64 * 2;
constexpr int G_C2_ROWS = 136;
constexpr uint32_t G_C2_KB = G_C2_ROWS * 32 * 2;       // [136 x 32] SWIZZLE_64B k-block of the C2 tile

struct GSmem {
  static constexpr uint32_t w3 = 0;
  static constexpr uint32_t w2 = w3 + 12 * G_W3_BLK;                 // 98304
  static constexpr uint32_t w1 = w2 + G_W2_BYTES;
  static constexpr uint32_t a1 = w1 + G_W1_BYTES;
  static constexpr uint32_t p1 = a1 + G_A1_BYTES;
  static constexpr uint32_t c2 = p1 + G_P1_BYTES;                    // 4 k-blocks
  static constexpr uint32_t stage = c2 + 4 * G_C2_KB;                // 8 epilogue warps x 2 KB
  static constexpr uint32_t bars = stage + 8 * EPI_STAGE_BYTES;
  static constexpr uint32_t nbars = 1 + 18;
  static constexpr uint32_t tmem_slot = (bars + 8 * nbars + 15u) & ~15u;
  static constexpr uint32_t total = tmem_slot + 16;
};

__device__ __forceinline__ uint32_t idesc16(bool f16, int n) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
template <bool F16> __device__ __forceinline__ uint16_t to16(float v) {
  if constexpr (F16) return __half_as_ushort(__float2half_rn(v));
  else return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <bool F16> __device__ __forceinline__ float from16(uint16_t h) {
  if constexpr (F16) return __half2float(__ushort_as_half(h));
  else return __uint_as_float((uint32_t)h << 16);
}

// lo part of a packed pair of fp32 values whose hi parts are already packed in `hi` (split-operand mode, fp16)
__device__ __forceinline__ uint32_t pack_lo(float a, float b, uint32_t hi) {
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  return *reinterpret_cast<const uint32_t*>(&l);
}

// SPLIT (DSB_PREC_FP16X2): P1, C2 and X0 are (hi, lo) fp16 pairs, the weights stay single fp16 and every conv is
// [xh | xl] . [W ; W].  Shared memory has no room for a second P1 / C2 tile, so the hi tile and the lo tile of a stage go
// through the SAME buffer one after the other: the staging warps read the accumulator from TMEM twice (first pass packs
// the hi parts, second pass the lo parts) and the MMA thread interleaves the stages so that neither pass waits on an MMA
// that is issued behind it:   conv3(k-2).hi  conv2(k-1).hi  conv3(k-2).lo[0,1]  conv2(k-1).lo  conv3(k-2).lo[2,3]  conv1(k)
template <bool F16, bool SPLIT>
__global__ void __launch_bounds__(G_THREADS, 1)
front2_kernel(const __grid_constant__ FrontMaps maps, const __grid_constant__ FrontParams p, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbar = base + GSmem::bars;
  // C2_FULL / C2_EMPTY exist per 32-channel k-block of the C2 tile: conv3 starts on k-block 0 while the staging warps
  // still convert k-blocks 1..3, and the next tile's staging re-uses a k-block as soon as its three tap MMAs are done
  enum { A1_FULL = 0, A1_EMPTY, ACC1_FULL, ACC1_EMPTY, P1_FULL, P1_EMPTY, ACC2_FULL, ACC2_EMPTY, ACC3_FULL, ACC3_EMPTY, C2_FULL, C2_EMPTY = C2_FULL + 4 };
  auto bar = [&](int kind) { return base + GSmem::bars + 8u * (1 + kind); };
  const uint32_t tmem_slot = base + GSmem::tmem_slot;

  const int nh = (int)(blockIdx.x & 1);                  // which 128-channel half of conv3 this CTA owns
  const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);
  const int my_tiles = t_first < num_tiles ? (num_tiles - t_first + t_step - 1) / t_step : 0;

  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    mbar_init(bar(A1_FULL), 4);   mbar_init(bar(A1_EMPTY), 1);
    mbar_init(bar(ACC1_FULL), 1); mbar_init(bar(ACC1_EMPTY), 4);
    mbar_init(bar(P1_FULL), 4);   mbar_init(bar(P1_EMPTY), 1);
    mbar_init(bar(ACC2_FULL), 1); mbar_init(bar(ACC2_EMPTY), 4);
    for (int kb = 0; kb < 4; ++kb) { mbar_init(bar(C2_FULL + kb), 4); mbar_init(bar(C2_EMPTY + kb), 1); }
    mbar_init(bar(ACC3_FULL), 1); mbar_init(bar(ACC3_EMPTY), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  // TMEM columns: conv1 [0,192) (tap a, channel c at a*64+c), conv2 [192,320), conv3 [320,448)
  constexpr uint32_t ACC2 = 192, ACC3 = 320;

  if (warp == 0) {
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(wbar, G_W1_BYTES + G_W2_BYTES + 12 * G_W3_BLK);
      tma_load_2d(base + GSmem::w1, &maps.w1, wbar, 0, 0);
      tma_load_2d(base + GSmem::w2, &maps.w2, wbar, 0, 0);
      for (int j = 0; j < 12; ++j) tma_load_2d(base + GSmem::w3 + j * G_W3_BLK, &maps.w3, wbar, j * 32, nh * 128);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && my_tiles > 0) {
      const uint32_t id128 = idesc16(F16, 128), id192 = idesc16(true, 192);     // conv1 parts are fp16 in both modes
      mbar_wait(wbar, 0);
      auto mma1 = [&](int k) {
        const uint32_t ph = (uint32_t)k & 1u;
        mbar_wait(bar(ACC1_EMPTY), ph ^ 1u);
        mbar_wait(bar(A1_FULL), ph);
        tc_fence_after();
        const uint64_t ad = make_kmajor_desc<64>(base + GSmem::a1);
        const uint64_t bd = make_kmajor_desc<64>(base + GSmem::w1);
#pragma unroll
        for (int j = 0; j < 3; ++j) tc_mma_bf16(tmem_base, ad + 2 * j, bd + 2 * j, id192, j != 0 ? 1u : 0u);
        tc_commit(bar(A1_EMPTY));
        tc_commit(bar(ACC1_FULL));
      };
      auto mma2 = [&](int k) {
        const uint32_t ph = (uint32_t)k & 1u;
        mbar_wait(bar(ACC2_EMPTY), ph ^ 1u);
        mbar_wait(bar(P1_FULL), ph);
        tc_fence_after();
        const uint64_t ad = make_kmajor_desc<64>(base + GSmem::p1);
        const uint64_t bd = make_kmajor_desc<64>(base + GSmem::w2);
#pragma unroll
        for (int j = 0; j < 4; ++j) tc_mma_bf16(tmem_base + ACC2, ad + 2 * j, bd + 2 * j, id128, j != 0 ? 1u : 0u);
        tc_commit(bar(P1_EMPTY));
        tc_commit(bar(ACC2_FULL));
      };
      auto mma3 = [&](int k) {
        const uint32_t ph = (uint32_t)k & 1u;
        mbar_wait(bar(ACC3_EMPTY), ph ^ 1u);
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(C2_FULL + kb), ph);
          tc_fence_after();
          for (int tap = 0; tap < 3; ++tap) {
            // output row i, tap t reads C2 row i + t - 1 = tile row i + t (the staging warps store one row down)
            const uint64_t ad = make_kmajor_desc<32>(base + GSmem::c2 + kb * G_C2_KB + (uint32_t)tap * 64u);
            const uint64_t bd = make_kmajor_desc<32>(base + GSmem::w3 + (tap * 4 + kb) * G_W3_BLK);
#pragma unroll
            for (int j = 0; j < 2; ++j) tc_mma_bf16(tmem_base + ACC3, ad + 2 * j, bd + 2 * j, id128, (tap | kb | j) != 0 ? 1u : 0u);
          }
          tc_commit(bar(C2_EMPTY + kb));
        }
        tc_commit(bar(ACC3_FULL));
      };
      // split mode: part 0 = hi tile, part 1 = lo tile of the same buffer (fill number 2k + part of its barrier pair)
      auto mma2s = [&](int k, int part) {
        if (part == 0) mbar_wait(bar(ACC2_EMPTY), ((uint32_t)k & 1u) ^ 1u);
        mbar_wait(bar(P1_FULL), (uint32_t)(2 * k + part) & 1u);
        tc_fence_after();
        const uint64_t ad = make_kmajor_desc<64>(base + GSmem::p1);
        const uint64_t bd = make_kmajor_desc<64>(base + GSmem::w2);
#pragma unroll
        for (int j = 0; j < 4; ++j) tc_mma_bf16(tmem_base + ACC2, ad + 2 * j, bd + 2 * j, id128, (part | j) != 0 ? 1u : 0u);
        tc_commit(bar(P1_EMPTY));
        if (part == 1) tc_commit(bar(ACC2_FULL));
      };
      auto mma3s = [&](int k, int part, int kb0, int kb1) {
        if (part == 0 && kb0 == 0) mbar_wait(bar(ACC3_EMPTY), ((uint32_t)k & 1u) ^ 1u);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar(C2_FULL + kb), (uint32_t)(2 * k + part) & 1u);
          tc_fence_after();
          for (int tap = 0; tap < 3; ++tap) {
            const uint64_t ad = make_kmajor_desc<32>(base + GSmem::c2 + kb * G_C2_KB + (uint32_t)tap * 64u);
            const uint64_t bd = make_kmajor_desc<32>(base + GSmem::w3 + (tap * 4 + kb) * G_W3_BLK);
#pragma unroll
            for (int j = 0; j < 2; ++j) tc_mma_bf16(tmem_base + ACC3, ad + 2 * j, bd + 2 * j, id128, (part | tap | kb | j) != 0 ? 1u : 0u);
          }
          tc_commit(bar(C2_EMPTY + kb));
        }
        if (part == 1 && kb1 == 4) tc_commit(bar(ACC3_FULL));
      };
      for (int it = 0; it < my_tiles + 2; ++it) {
        if constexpr (SPLIT) {
          const bool c3 = it >= 2, c2 = it >= 1 && it - 1 < my_tiles;
          // conv2's lo part (and with it ACC2_FULL of the next tile) goes out in the middle of conv3's lo part: the C2
          // staging of the next tile then overlaps the second half of these MMAs instead of starting after all of them
          if (c3) mma3s(it - 2, 0, 0, 4);
          if (c2) mma2s(it - 1, 0);
          if (c3) mma3s(it - 2, 1, 0, 2);
          if (c2) mma2s(it - 1, 1);
          if (c3) mma3s(it - 2, 1, 2, 4);
          if (it < my_tiles) mma1(it);
        } else {
          if (it < my_tiles) mma1(it);
          if (it >= 1 && it - 1 < my_tiles) mma2(it - 1);
          if (it >= 2) mma3(it - 2);
        }
      }
    }
  } else if (warp < 6) {
    // ---------------- im2col warps: signal window of every tile row as (hi | lo | hi) 16-bit parts ----------------
    const int i = (warp - 2) * 32 + lane;        // tile row
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      const int r = tile * FRONT_TILE_ROWS - 1 + i;
      bool valid = r >= 0 && r < p.M;
      const int s = valid ? r / p.Lp : 0;
      const int pos = r - s * p.Lp;
      valid = valid && pos < p.L;
      uint16_t h[48];
#pragma unroll
      for (int k = 0; k < 48; ++k) h[k] = 0;
      if (valid) {
        const float* sig = p.signals + (size_t)s * p.S;
        const int j0 = 2 * (2 * pos - p.pool_pl) - p.conv_pl;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const int j = j0 + k;
          const float x = (j >= 0 && j < p.S) ? __ldg(sig + j) : 0.f;
          // fp16 parts (also in bf16 mode: 2^-22 instead of 2^-16).  |x| > 65504 becomes inf / NaN on purpose: it reaches the
          // head kernel's non-finite check (DSB_ERR_NONFINITE) instead of being clamped into a silent wrong answer.
          const uint16_t xh = to16<true>(x);
          const uint16_t xl = to16<true>(x - from16<true>(xh));
          h[k] = xh; h[11 + k] = xl; h[22 + k] = xh;
        }
      }
      {
        // columns 33..35 meet a 1.0 in the conv1 operand rows of pool tap 0 / 1 / 2: a tap that lies outside the conv1
        // output (SAME padding of the pool), and every tap of a halo row, gets -65504 added, so the pool's max ignores it
        // and relu(max + bias) of a halo row is 0 without any per-element select in the pool staging warps
        const int q0 = 2 * pos - p.pool_pl;
#pragma unroll
        for (int a = 0; a < 3; ++a) h[33 + a] = (valid && q0 + a >= 0 && q0 + a < p.L1) ? (uint16_t)0 : (uint16_t)0xFBFF;
      }
      mbar_wait(bar(A1_EMPTY), ((uint32_t)kk & 1u) ^ 1u);
      const uint32_t rowp = base + GSmem::a1 + (uint32_t)i * 128u;
#pragma unroll
      for (int pc = 0; pc < 6; ++pc) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = (uint32_t)h[pc * 8 + 2 * j] | ((uint32_t)h[pc * 8 + 2 * j + 1] << 16);
        sts128(rowp + ((((uint32_t)pc) ^ ((uint32_t)i & 7u)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(A1_FULL));
    }
  } else if (warp < 10) {
    // ---------------- pool staging warps: relu(max over valid taps + bias) -> 16-bit P1 tile ----------------
    const int q = warp & 3;
    const int i = q * 32 + lane;
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      const uint32_t ph = (uint32_t)kk & 1u;
      const int r = tile * FRONT_TILE_ROWS - 1 + i;
      bool valid = r >= 0 && r < p.M;
      const int s = valid ? r / p.Lp : 0;
      const int pos = r - s * p.Lp;
      valid = valid && pos < p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      (void)valid;
      mbar_wait(bar(ACC1_FULL), ph);
      const uint32_t rowp = base + GSmem::p1 + (uint32_t)i * 128u;
      constexpr int PARTS = SPLIT ? 2 : 1;
      // split mode: two passes over the accumulator (hi parts, then lo parts into the same tile).  Keeping the 32 lo
      // registers of a row across the wait instead measured slower (spills: the kernel runs at 80 registers per thread).
#pragma unroll 1
      for (int part = 0; part < PARTS; ++part) {
        // fill number kk (single) or 2 kk + part (split) of the P1 tile: wait until the conv2 MMAs of the previous fill have read it
        mbar_wait(bar(P1_EMPTY), ((uint32_t)(PARTS * kk + part) & 1u) ^ 1u);
        tc_fence_after();
#pragma unroll 1
        for (int cg = 0; cg < 4; ++cg) {                    // 16 channels at a time
          // taps outside the conv1 output (and whole halo rows) carry the -65504 penalty of the im2col tile: no selects here
          uint32_t a0[16], a1[16], a2[16];
          tc_ld16(t_row + (uint32_t)(cg * 16), a0);
          tc_ld16(t_row + (uint32_t)(64 + cg * 16), a1);
          tc_ld16(t_row + (uint32_t)(128 + cg * 16), a2);
          tc_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 2 * k + e;
              const float m = fmaxf(__uint_as_float(a0[c]), fmaxf(__uint_as_float(a1[c]), __uint_as_float(a2[c])));
              v[e] = relu_keep_nan(m + p.b1v[cg * 16 + c]);
            }
            pk[k] = pack16<F16>(v[0], v[1]);
            if (SPLIT && part == 1) pk[k] = pack_lo(v[0], v[1], pk[k]);
          }
          sts128(rowp + ((((uint32_t)(2 * cg)) ^ ((uint32_t)i & 7u)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          sts128(rowp + ((((uint32_t)(2 * cg + 1)) ^ ((uint32_t)i & 7u)) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(P1_FULL)); if (part == PARTS - 1) mbar_arrive(bar(ACC1_EMPTY)); }
      }
    }
  } else if (warp < 14) {
    // ---------------- C2 staging warps: relu(conv2 + bias) -> 16-bit A operand of conv3 ----------------
    const int q = warp & 3;
    const int i = q * 32 + lane;
    const uint32_t rr = (uint32_t)(i + 1);                   // tile row (one row down)
    const uint32_t sw = (rr >> 1) & 3u;                      // SWIZZLE_64B: piece c sits at c ^ ((row >> 1) & 3)
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      const uint32_t ph = (uint32_t)kk & 1u;
      const int row = tile * FRONT_TILE_ROWS - 1 + i;
      const bool zero = row < 0 || row >= p.M || (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + ACC2;
      mbar_wait(bar(ACC2_FULL), ph);
      tc_fence_after();
      constexpr int PARTS = SPLIT ? 2 : 1;
      const uint32_t keep = zero ? 0u : 0xffffffffu;
      // fully unrolled: the bias then comes straight from the constant bank as an instruction operand (an indexed
      // p.b2v[kb * 32 + e] costs one LDC per element, and this stage is what paces the kernel in split mode)
#pragma unroll
      for (int part = 0; part < PARTS; ++part) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          uint32_t r2[2][16];
          tc_ld16(t_row + (uint32_t)(kb * 32), r2[0]);
          tc_ld16(t_row + (uint32_t)(kb * 32 + 16), r2[1]);
          tc_wait_ld();
          if (kb == 3 && part == PARTS - 1) {                  // conv2's accumulator is in registers: conv2 of the next tile may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(ACC2_EMPTY));
          }
          // fill number kk (single) or 2 kk + part (split) of this k-block: its previous fill has been read by the conv3 MMAs
          mbar_wait(bar(C2_EMPTY + kb), ((uint32_t)(PARTS * kk + part) & 1u) ^ 1u);
          const uint32_t ub = base + GSmem::c2 + kb * G_C2_KB + rr * 64u;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = c * 8 + 2 * k;
              const float a = relu_keep_nan(__uint_as_float(r2[e >> 4][e & 15]) + p.b2v[kb * 32 + e]);
              const float d = relu_keep_nan(__uint_as_float(r2[(e + 1) >> 4][(e + 1) & 15]) + p.b2v[kb * 32 + e + 1]);
              pk[k] = pack16<F16>(a, d);
              if (SPLIT && part == 1) pk[k] = pack_lo(a, d, pk[k]);
              pk[k] &= keep;                               // halo / out-of-range rows are the zero padding of conv3
            }
            sts128(ub + (((uint32_t)c ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(C2_FULL + kb));
        }
      }
    }
  } else {
    // ---------------- epilogue warps: relu(conv3 + bias) -> X0 ----------------
    const int q = warp & 3;
    const int half = (warp - 14) >> 2;
    const uint32_t stage = base + GSmem::stage + (uint32_t)(warp - 14) * EPI_STAGE_BYTES;
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      const uint32_t ph = (uint32_t)kk & 1u;
      const int i = q * 32 + lane;
      const int row = tile * FRONT_TILE_ROWS - 1 + i;
      const int wrow0 = tile * FRONT_TILE_ROWS - 1 + q * 32;
      const bool row_ok = i >= 1 && i <= FRONT_TILE_ROWS && row < p.M;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const bool halo = row < 0 || (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + ACC3;
      mbar_wait(bar(ACC3_FULL), ph);
      tc_fence_after();
      // both 32-column groups of this warp are read before any arithmetic, so that the conv3 MMAs of the next tile (the
      // accumulator is single-buffered: TMEM holds conv1 192 + conv2 128 + conv3 128 columns) wait for two TMEM loads, not
      // for the first group's stores
      uint32_t racc[2][2][16];
      __syncwarp();
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
        tc_ld16(t_row + (uint32_t)((half + 2 * gi) * 32), racc[gi][0]);
        tc_ld16(t_row + (uint32_t)((half + 2 * gi) * 32 + 16), racc[gi][1]);
      }
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(ACC3_EMPTY));
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
        const int g = half + 2 * gi;
        uint32_t pk[16], pl[16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int col = g * 32 + 16 * h2;
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(racc[gi][h2][k]) + p.b3v[nh * 128 + col + k];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = halo ? 0.f : relu_keep_nan(v[k]);
#pragma unroll
          for (int k = 0; k < 8; ++k) pk[8 * h2 + k] = pack16<F16>(v[2 * k], v[2 * k + 1]);
          if constexpr (SPLIT) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pl[8 * h2 + k] = pack_lo(v[2 * k], v[2 * k + 1], pk[8 * h2 + k]);
          }
        }
        const size_t pitch = (size_t)p.ldX * 2;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.X0) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch + (size_t)(nh * 128 + g * 32) * 2;
        staged_store<4>(stage, pk, lane, ok_mask, dst, pitch);
        if constexpr (SPLIT) staged_store<4>(stage, pl, lane, ok_mask, dst + p.lo_x0 * 2, pitch);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

void launch_front2(const FrontLaunch& L, const float* signals, int n, int sm_count, cudaStream_t s) {
  if (n <= 0) return;
  FrontParams p = L.p;
  p.signals = signals;
  p.M = n * p.Lp;
  const int num_tiles = (p.M + FRONT_TILE_ROWS - 1) / FRONT_TILE_ROWS;
  int grid = sm_count & ~1;
  if (grid > 2 * num_tiles) grid = 2 * num_tiles;
  if (grid < 2) grid = 2;
  const size_t smem = GSmem::total + 1024;
  static bool cfg[3][64] = {{false}};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  if (L.split && !L.fp16) fail(DSB_ERR_STATE, "internal: the split front end is fp16 only");
  auto kern = L.split ? front2_kernel<true, true> : L.fp16 ? front2_kernel<true, false> : front2_kernel<false, false>;
  const int which = L.split ? 2 : L.fp16 ? 1 : 0;
  if (!cfg[which][dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg[which][dev & 63] = true;
  }
  launch_pdl(kern, grid, G_THREADS, smem, s, L.maps, p, num_tiles);
  DSB_CUDA(cudaGetLastError());
}

}  // namespace dsb
