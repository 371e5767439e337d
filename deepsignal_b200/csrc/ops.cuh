// ops.cuh — the GEMM-with-epilogue "op" both forward backends execute, and shared device math.
//
// Every heavy layer of the call_mods graph is expressed as
//     D[M, N] = epilogue( sum_s  A_s[row + drow_s, col0_s : col0_s + K_s] · B[:, koff_s : koff_s + K_s]^T )
// where the A "segments" cover 1x1 convs (one segment), k-tap convs (one segment per tap, drow = tap - pad,
// reading the zero halo rows of the padded row layout at sample edges) and LSTM steps (x_t | h_{t-1}).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dsb {

constexpr int MAX_SEG = 15;         // 5 taps x (hi.wh, lo.wh, hi.wl) product terms in the three-term split mode

enum EpiKind { EPI_CONV = 0, EPI_LSTM = 1 };

struct EpiParams {
  int M = 0;              // valid rows
  int N = 0;              // output columns, multiple of 16
  const float* bias = nullptr;
  // EPI_CONV: columns [0, split) -> d0[:, c0 + col], columns [split, N) -> d1[:, c1 + col - split]
  void* d0 = nullptr; int ld0 = 0; int c0 = 0;
  void* d1 = nullptr; int ld1 = 0; int c1 = 0;
  int split = 0;
  uint32_t relu_mask = 0; // bit j: ReLU on columns [16j, 16j+16)
  const void* res = nullptr; int res_ld = 0; int res_c0 = 0;   // residual added before ReLU (d0 columns)
  int Lp = 0, L = 0;      // padded row layout: rows with (row % Lp) >= L are written as zero
  // split-operand mode (DSB_PREC_FP16X2): every activation is stored as hi = fp16(v) and lo = fp16(v - hi) in two
  // matrices of the same shape; lo_off* = element offset from the hi matrix to its lo twin (0 = single 16-bit output)
  long long lo_off0 = 0, lo_off1 = 0, res_lo_off = 0;
  // EPI_LSTM: columns are 4*u+g; c [M][256] fp32 updated in place, h_out [M][256]
  float* c = nullptr;
  void* h_out = nullptr;
};

struct SegF32 {           // fp32 backend: plain pointers
  const float* a; int ld; int col0; int drow; int K;
};

struct GemmF32Params {
  int nseg;
  SegF32 seg[MAX_SEG];
  int rows_valid;         // A rows outside [0, rows_valid) read as zero
  const float* B;         // [N][Ktot] K-major
  int Ktot;
  EpiParams epi;
};

template <class T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <class T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ReLU that keeps NaN (fmaxf(NaN, 0) is 0): an overflowed 16-bit activation (inf, then inf - inf) must reach the head
// kernel's non-finite check instead of being laundered into a finite wrong answer.
__device__ __forceinline__ float relu_keep_nan(float v) {
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));      // one FMNMX.NAN instead of a compare + select
  return r;
}
// max that keeps NaN (fmaxf drops it): used by the pooling kernels for the same reason
__device__ __forceinline__ float max_keep_nan(float a, float b) { return (b > a || b != b) ? b : a; }

// exact-ish activations for the fp32 path, MUFU-based ones for the bf16 path
template <bool FAST> __device__ __forceinline__ float act_tanh(float x) {
  if constexpr (FAST) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    return tanhf(x);
  }
}
template <bool FAST> __device__ __forceinline__ float act_sigmoid(float x) {
  if constexpr (FAST) {
    return fmaf(act_tanh<true>(0.5f * x), 0.5f, 0.5f);
  } else {
    return 1.0f / (1.0f + expf(-x));
  }
}

// LSTMCell step for one hidden unit (reference layers.py:49-50; tf LSTMCell, forget_bias = 1.0)
template <bool FAST>
__device__ __forceinline__ void lstm_unit(float zi, float zj, float zf, float zo, float c_prev, float& c_new, float& h_new) {
  c_new = act_sigmoid<FAST>(zf + 1.0f) * c_prev + act_sigmoid<FAST>(zi) * act_tanh<FAST>(zj);
  h_new = act_sigmoid<FAST>(zo) * act_tanh<FAST>(c_new);
}

// Two hidden units at once for the tensor-core path: the five transcendental evaluations per unit go through
// MUFU tanh.approx.f16x2 (two results per MUFU op; sigmoid(x) = 0.5*tanh(0.5x)+0.5), i.e. 2.5 MUFU ops per unit
// instead of 5.  tanh.approx.f32 has the same 2^-11 relative accuracy, so the accuracy class does not change;
// the MUFU pipe (16 lanes/clk/SM) was the bottleneck of the LSTM epilogue.
__device__ __forceinline__ float2 tanh_pair_f16(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(u) : "r"(u));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ void lstm_unit_pair_fast(const float (&za)[4], const float (&zb)[4], float ca, float cb,
                                                    float& ca_new, float& cb_new, float& ha, float& hb) {
  // z = (i, j, f, o) pre-activations of unit a / unit b (bias already added)
  const float2 a_if = tanh_pair_f16(0.5f * za[0], 0.5f * (za[2] + 1.0f));
  const float2 a_oj = tanh_pair_f16(0.5f * za[3], za[1]);
  const float2 b_if = tanh_pair_f16(0.5f * zb[0], 0.5f * (zb[2] + 1.0f));
  const float2 b_oj = tanh_pair_f16(0.5f * zb[3], zb[1]);
  ca_new = fmaf(fmaf(a_if.y, 0.5f, 0.5f), ca, fmaf(a_if.x, 0.5f, 0.5f) * a_oj.y);
  cb_new = fmaf(fmaf(b_if.y, 0.5f, 0.5f), cb, fmaf(b_if.x, 0.5f, 0.5f) * b_oj.y);
  const float2 tc = tanh_pair_f16(ca_new, cb_new);
  ha = fmaf(a_oj.x, 0.5f, 0.5f) * tc.x;
  hb = fmaf(b_oj.x, 0.5f, 0.5f) * tc.y;
}

// ---- launch helpers implemented in the .cu files -------------------------------------------------
void launch_gemm_f32(const GemmF32Params& p, int epi_kind, cudaStream_t s);

}  // namespace dsb
