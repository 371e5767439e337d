// lstm_epi.cuh — the LSTM cell update of one 64-gate-column group (16 hidden units) of one TMEM lane quadrant, shared by
// the single-CTA and the CTA-pair sequence kernels.
#pragma once
#include "tc_common.cuh"

namespace dsb {

// t_addr: TMEM address of the group's first column for this warp's lane quadrant; cblk: this lane's first float4 of the
// group's cell-state block (4 float4 per lane, 32 lanes apart); bias: smem bias of the group's first column.
// 32 columns (8 hidden units) at a time keep the live registers < 96.  `release` is called once, right after the last
// tcgen05.ld of the group has completed.
// The cell state is re-read by the same CTA one step later: keep its lines in L2 (evict_last) while the H slots and
// weights stream through.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_c(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_c(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// Loads this lane's four cell-state float4 of a group (call it BEFORE waiting for the accumulator: the L2 latency then
// overlaps the MMA wait instead of sitting in the epilogue's critical path).
__device__ __forceinline__ void lstm_load_c(const float4* cblk, bool c_in, float4 (&c4)[4]) {
  const uint64_t pol = l2_policy_evict_last();
#pragma unroll
  for (int i = 0; i < 4; ++i) c4[i] = c_in ? ld_c(cblk + i * 32, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool F16, bool PRELOADED = false, class Release>
__device__ __forceinline__ void lstm_epi_group(uint32_t t_addr, float4* cblk, bool c_in, bool c_out, const float* bias,
                                               uint32_t (&pk)[8], int ablate, Release&& release, const float4* c4 = nullptr) {
  using namespace tc;
  const uint64_t pol = l2_policy_evict_last();
#pragma unroll
  for (int sub = 0; sub < 2; ++sub) {
    // the cell state is private to this CTA and was written one step (~20 us) ago: L2 hits
    float4 cprev[2];
    if constexpr (PRELOADED) {
      cprev[0] = c4[2 * sub];
      cprev[1] = c4[2 * sub + 1];
    } else if (c_in) {
      cprev[0] = ld_c(cblk + (2 * sub) * 32, pol);
      cprev[1] = ld_c(cblk + (2 * sub + 1) * 32, pol);
    } else {
      cprev[0] = cprev[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t r[2][16];
    __syncwarp();
    tc_ld16(t_addr + (uint32_t)(32 * sub), r[0]);
    tc_ld16(t_addr + (uint32_t)(32 * sub + 16), r[1]);
    tc_wait_ld();
    if (sub == 1) release();
    float hn[8];
#pragma unroll
    for (int i2 = 0; i2 < 2; ++i2) {
      const int i = 2 * sub + i2;
      const float cpv[4] = {cprev[i2].x, cprev[i2].y, cprev[i2].z, cprev[i2].w};
      float cn[4];
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + 16 * i + 4 * j);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + 16 * i + 4 * j + 4);
        const float za[4] = {__uint_as_float(r[i2][4 * j + 0]) + b0.x, __uint_as_float(r[i2][4 * j + 1]) + b0.y,
                             __uint_as_float(r[i2][4 * j + 2]) + b0.z, __uint_as_float(r[i2][4 * j + 3]) + b0.w};
        const float zb[4] = {__uint_as_float(r[i2][4 * j + 4]) + b1.x, __uint_as_float(r[i2][4 * j + 5]) + b1.y,
                             __uint_as_float(r[i2][4 * j + 6]) + b1.z, __uint_as_float(r[i2][4 * j + 7]) + b1.w};
        if (ablate & 4) {
          cn[j] = za[0] + za[1] + cpv[j]; cn[j + 1] = zb[0] + zb[1] + cpv[j + 1];
          hn[4 * i2 + j] = za[2] + za[3]; hn[4 * i2 + j + 1] = zb[2] + zb[3];
        } else {
          lstm_unit_pair_fast(za, zb, cpv[j], cpv[j + 1], cn[j], cn[j + 1], hn[4 * i2 + j], hn[4 * i2 + j + 1]);
        }
      }
      if (c_out) st_c(cblk + i * 32, make_float4(cn[0], cn[1], cn[2], cn[3]), pol);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[4 * sub + i] = pack16<F16>(hn[2 * i], hn[2 * i + 1]);
  }
}

}  // namespace dsb
