// elementwise.cuh — the HBM-bound helpers of the forward: fused stem conv (1x7 s2 + BN + ReLU +
// maxpool 1x3 s2), the two pooling shapes, LSTM layer-0 feature rows and the fused head.
// All kernels work on the padded row layout: sample s, position p -> row s*Lp + p, rows with
// p >= L are zero ("halo" rows shared between neighbouring samples).
#pragma once
#include "ops.cuh"

namespace dsb {

// ---- 8-channel vector helpers (16 B for bf16, 32 B for fp32) -------------------------------------
template <class T> struct Vec8;
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ static Vec8 load(const float* p) {
    Vec8 r;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ static Vec8 load(const __nv_bfloat16* p) {
    Vec8 r;
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <> struct Vec8<__half> {
  float v[8];
  __device__ __forceinline__ static Vec8 load(const __half* p) {
    Vec8 r;
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    }
    return r;
  }
  __device__ __forceinline__ void store(__half* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---- stem: signals [n][S] fp32 -> P1 [n*Lp2][64] --------------------------------------------------
// conv 1x7 stride 2 SAME (1->64) + folded BN + ReLU at L1 positions, then maxpool 1x3 stride 2 SAME
// (reference layers.py:183-191).  One CTA per sample, 256 threads; thread -> (channel, position stripe).
struct StemParams {
  const float* signals; int S;
  const float* w;      // [64][7] folded
  const float* bias;   // [64]
  int L1, L2, Lp2, conv_pl, pool_pl;
  void* out;           // [n*Lp2][64]
  int n;
  long long lo_off = 0;  // split-operand mode: element offset of the lo twin of `out` (0 = single 16-bit output)
};

template <class T>
__global__ void __launch_bounds__(256) stem_kernel(const StemParams p) {
  extern __shared__ float sm[];
  float* xs = sm;                       // S + 16 (zero padded both sides by 8)
  float* ws = xs + p.S + 16;            // 64*7
  float* bs = ws + 64 * 7;              // 64
  const int s = blockIdx.x;
  for (int i = threadIdx.x; i < p.S + 16; i += blockDim.x) {
    const int j = i - 8;
    xs[i] = (j >= 0 && j < p.S) ? p.signals[(size_t)s * p.S + j] : 0.f;
  }
  for (int i = threadIdx.x; i < 64 * 7; i += blockDim.x) ws[i] = p.w[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) bs[i] = p.bias[i];
  __syncthreads();
  T* out = reinterpret_cast<T*>(p.out) + (size_t)s * p.Lp2 * 64;
  const int c = threadIdx.x & 63;
  const int strip = threadIdx.x >> 6, nstrip = blockDim.x >> 6;
  float w[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) w[t] = ws[c * 7 + t];
  const float b = bs[c];
  // each thread walks a strip of pooled positions; conv positions advance by one (signal index by two), so the
  // 7-sample window slides in registers and conv(2p+2) is reused as conv(2(p+1)).
  const int per = (p.Lp2 + nstrip - 1) / nstrip;
  const int p0 = strip * per, p1 = min(p.Lp2, p0 + per);
  if (p0 >= p1) return;
  int q = 2 * p0 - p.pool_pl;                    // first conv position needed
  int x0 = 2 * q - p.conv_pl + 8;                // window start in xs (x0 >= 0 is guaranteed by the 8-sample pad for q >= -3)
  float win[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) win[t] = (x0 + t >= 0 && x0 + t < p.S + 16) ? xs[x0 + t] : 0.f;
  auto conv_here = [&]() {
    float a = b;
#pragma unroll
    for (int t = 0; t < 7; ++t) a = fmaf(w[t], win[t], a);
    return (q >= 0 && q < p.L1) ? relu_keep_nan(a) : 0.f;     // ReLU; invalid cells contribute 0 (<= every valid value)
  };
  auto slide = [&]() {
    ++q;
    x0 += 2;
#pragma unroll
    for (int t = 0; t < 5; ++t) win[t] = win[t + 2];
    win[5] = (x0 + 5 < p.S + 16) ? xs[x0 + 5] : 0.f;
    win[6] = (x0 + 6 < p.S + 16) ? xs[x0 + 6] : 0.f;
  };
  float c0 = conv_here();
  for (int pos = p0; pos < p1; ++pos) {
    slide();
    const float c1 = conv_here();
    slide();
    const float c2 = conv_here();
    const float best = (pos < p.L2) ? fmaxf(c0, fmaxf(c1, c2)) : 0.f;
    const T hi = from_f32<T>(best);
    out[(size_t)pos * 64 + c] = hi;
    if (p.lo_off != 0) out[(size_t)pos * 64 + c + p.lo_off] = from_f32<T>(best - to_f32<T>(hi));
    c0 = c2;
  }
}

// ---- maxpool 1x3 stride 1 SAME over rows (inputs >= 0, halo rows are zero) -------------------------
template <class T>
__global__ void pool3s1_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t M, int C, int ld) {
  const int cv = C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * cv) return;
  const int64_t r = idx / cv;
  const int c = (int)(idx % cv) * 8;
  Vec8<T> a = Vec8<T>::load(x + r * ld + c);
  if (r > 0) {
    const Vec8<T> b = Vec8<T>::load(x + (r - 1) * ld + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i] = max_keep_nan(a.v[i], b.v[i]);
  }
  if (r + 1 < M) {
    const Vec8<T> b = Vec8<T>::load(x + (r + 1) * ld + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i] = max_keep_nan(a.v[i], b.v[i]);
  }
  a.store(y + r * ld + c);
}

// ---- maxpool 1x3 stride 2 SAME between stages, 16-bit activations: one warp per output row, lanes over 16-byte channel
// vectors (coalesced 480-byte row segments), packed half2 / bf162 maxima that keep NaN, 32-bit index math.  SPLIT: a
// value is the pair (hi, lo) with v = hi + lo, |lo| <= ulp(hi)/2; hi is a monotone rounding of v, so ordering pairs by hi
// and then by lo orders them by v exactly: the winner has the largest hi and, among equal hi, the largest lo.  With
// stride 1 the same kernel is the branch-1 pool of the un-fused plan.
template <bool F16> __device__ __forceinline__ uint32_t pmax_nan(uint32_t a, uint32_t b) {
  if constexpr (F16) {
    const __half2 r = __hmax2_nan(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}
// (ah, al) <- max((ah, al), (bh, bl)) elementwise on two packed fp16 pairs
__device__ __forceinline__ void pair_max2(uint32_t& ah, uint32_t& al, uint32_t bh, uint32_t bl) {
  const __half2 A = *reinterpret_cast<const __half2*>(&ah), B = *reinterpret_cast<const __half2*>(&bh);
  const __half2 H = __hmax2_nan(A, B);
  const uint32_t NEG = 0xfc00fc00u;                      // -inf, -inf
  const uint32_t ma = __heq2_mask(A, H), mb = __heq2_mask(B, H);
  const uint32_t la = (al & ma) | (NEG & ~ma), lb = (bl & mb) | (NEG & ~mb);
  const __half2 L = __hmax2_nan(*reinterpret_cast<const __half2*>(&la), *reinterpret_cast<const __half2*>(&lb));
  ah = *reinterpret_cast<const uint32_t*>(&H);
  al = *reinterpret_cast<const uint32_t*>(&L);
}

template <bool SPLIT, bool F16>
__global__ void __launch_bounds__(256) pool3s2_16_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int rows_out, int cv, int ld,
                                                        int Lin, int Lpin, int Lout, int Lpout, int pl, long long lo_in, long long lo_out,
                                                        int stride = 2) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)(gridDim.x * (blockDim.x >> 5));
  for (int row = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); row < rows_out; row += nwarps) {
    const int s = row / Lpout, po = row - s * Lpout;
    for (int c = lane; c < cv; c += 32) {
      uint4 h = make_uint4(0u, 0u, 0u, 0u), l = make_uint4(0u, 0u, 0u, 0u);     // activations are >= 0 (post-ReLU): zero is neutral
      if (po < Lout) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int q = stride * po - pl + k;           // stride 2 between stages; stride 1 = the branch-1 pool of the un-fused plan
          if (q < 0 || q >= Lin) continue;
          const uint16_t* src = x + ((size_t)s * Lpin + q) * ld + c * 8;
          const uint4 vh = __ldg(reinterpret_cast<const uint4*>(src));
          if constexpr (SPLIT) {
            const uint4 vl = __ldg(reinterpret_cast<const uint4*>(src + lo_in));
            pair_max2(h.x, l.x, vh.x, vl.x); pair_max2(h.y, l.y, vh.y, vl.y);
            pair_max2(h.z, l.z, vh.z, vl.z); pair_max2(h.w, l.w, vh.w, vl.w);
          } else {
            h.x = pmax_nan<F16>(h.x, vh.x); h.y = pmax_nan<F16>(h.y, vh.y);
            h.z = pmax_nan<F16>(h.z, vh.z); h.w = pmax_nan<F16>(h.w, vh.w);
          }
        }
      }
      uint16_t* dst = y + (size_t)row * ld + c * 8;
      *reinterpret_cast<uint4*>(dst) = h;
      if constexpr (SPLIT) *reinterpret_cast<uint4*>(dst + lo_out) = l;
    }
  }
}

// ---- maxpool 1x3 stride 2 SAME between stages (reference layers.py:211-213,224-226): generic / fp32 version -------
template <class T>
__global__ void pool3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, int C, int ld,
                               int Lin, int Lpin, int Lout, int Lpout, int pl) {
  const int cv = C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * Lpout * cv) return;
  const int c = (int)(idx % cv) * 8;
  const int64_t ro = idx / cv;
  const int64_t s = ro / Lpout;
  const int po = (int)(ro % Lpout);
  Vec8<T> a;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.v[i] = 0.f;
  if (po < Lout) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int q = 2 * po - pl + k;
      if (q < 0 || q >= Lin) continue;
      const Vec8<T> b = Vec8<T>::load(x + (s * Lpin + q) * ld + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) a.v[i] = max_keep_nan(a.v[i], b.v[i]);
    }
  }
  a.store(y + ro * ld + c);
}

// ---- LSTM layer-0 feature rows: F[pos][n][FW] = onehot(code) | mean | std | len | 0... -------------
template <class T>
__global__ void lstm_feat_kernel(const int32_t* __restrict__ kmer, const float* __restrict__ means,
                                 const float* __restrict__ stds, const float* __restrict__ lens,
                                 T* __restrict__ F, int64_t n, int64_t cap, int klen, int FW, int use_base) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * klen) return;
  const int64_t s = idx / klen;
  const int pos = (int)(idx % klen);
  const int code = use_base ? kmer[idx] : -1;
  Vec8<T> v;
#pragma unroll
  for (int j = 0; j < 5; ++j) v.v[j] = (j == code) ? 1.f : 0.f;
  v.v[5] = means[idx];
  v.v[6] = stds[idx];
  v.v[7] = lens[idx];
  v.store(F + ((int64_t)pos * cap + s) * FW);     // columns 8..FW-1 stay zero (workspace is zeroed at plan build)
}

// ---- head: logits = h_fw·Wr[0:256] + h_bw·Wr[256:512] + sum_{p,c} y[s*Lp+p][c]·Wc[p][c] ------------
// (dense·dense_1 collapsed, avgpool folded into Wc; sigmoid + argmax fused; reference layers.py:250-264,
// model.py:100,108).  One warp per sample.
struct HeadParams {
  const void* y; int Lp, L, ldy;      // final inception output [n*Lp][ldy] (240 live channels) or nullptr
  const void* hfw; const void* hbw;   // [n][256] or nullptr
  const float* Wc;                    // [L*240][2]
  const float* Wr;                    // [512][2]
  float* probs; int32_t* pred; int64_t n;
  long long y_lo_off = 0;             // split-operand mode: element offset of the lo twin of y (0 = single 16-bit values)
  int* nonfinite = nullptr;           // device counter: samples whose logits are not finite (overflowed 16-bit activations)
};

template <class T>
__global__ void __launch_bounds__(256) head_kernel(const HeadParams p) {
  extern __shared__ float hw[];                    // Wc [L*240*2] then Wr [1024]
  const int ncnn = (p.y != nullptr) ? p.L * CH * 2 : 0;
  for (int i = threadIdx.x; i < ncnn; i += blockDim.x) hw[i] = p.Wc[i];
  if (p.hfw != nullptr) for (int i = threadIdx.x; i < 4 * HID; i += blockDim.x) hw[ncnn + i] = p.Wr[i];
  __syncthreads();
  const float* wr = hw + ncnn;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  // A warp takes HS consecutive samples per iteration so that every weight vector read from shared memory is used HS times
  // (the kernel was bound by the shared-memory pipe: 44 KB of weight reads per 11 KB of activations).  The per-sample FMA
  // order is unchanged.
  constexpr int HS = 4;
  for (int64_t s0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * HS; s0 < p.n; s0 += (int64_t)gridDim.x * wpb * HS) {
    float a0[HS], a1[HS];
#pragma unroll
    for (int k = 0; k < HS; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
    const int ns = (int)((p.n - s0) < HS ? (p.n - s0) : HS);
    if (p.y != nullptr) {
      const T* y = reinterpret_cast<const T*>(p.y) + s0 * p.Lp * p.ldy;
      const size_t sstride = (size_t)p.Lp * p.ldy;
      const int nv = p.L * CH / 8;                 // rows 0..L-1 of a sample, 30 vectors of 8 channels per row
      for (int i = lane; i < nv; i += 32) {
        const size_t off = (size_t)(i / (CH / 8)) * p.ldy + (i % (CH / 8)) * 8;
        Vec8<T> v[HS];
#pragma unroll
        for (int k = 0; k < HS; ++k) v[k] = Vec8<T>::load(y + (k < ns ? k : 0) * sstride + off);
        if (p.y_lo_off != 0) {
#pragma unroll
          for (int k = 0; k < HS; ++k) {
            const Vec8<T> lo = Vec8<T>::load(y + (k < ns ? k : 0) * sstride + off + p.y_lo_off);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[k].v[e] += lo.v[e];
          }
        }
        const float4* w = reinterpret_cast<const float4*>(hw + (size_t)i * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = w[q];
#pragma unroll
          for (int k = 0; k < HS; ++k) {
            a0[k] = fmaf(v[k].v[2 * q], ww.x, a0[k]); a1[k] = fmaf(v[k].v[2 * q], ww.y, a1[k]);
            a0[k] = fmaf(v[k].v[2 * q + 1], ww.z, a0[k]); a1[k] = fmaf(v[k].v[2 * q + 1], ww.w, a1[k]);
          }
        }
      }
    }
    if (p.hfw != nullptr) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const T* hsrc = reinterpret_cast<const T*>(half == 0 ? p.hfw : p.hbw) + s0 * HID;
        Vec8<T> v[HS];
#pragma unroll
        for (int k = 0; k < HS; ++k) v[k] = Vec8<T>::load(hsrc + (size_t)(k < ns ? k : 0) * HID + lane * 8);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 ww = reinterpret_cast<const float2*>(wr)[half * HID + lane * 8 + q];
#pragma unroll
          for (int k = 0; k < HS; ++k) { a0[k] = fmaf(v[k].v[q], ww.x, a0[k]); a1[k] = fmaf(v[k].v[q], ww.y, a1[k]); }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < HS; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0[k] += __shfl_xor_sync(0xffffffffu, a0[k], o);
        a1[k] += __shfl_xor_sync(0xffffffffu, a1[k], o);
      }
      if (lane == 0 && k < ns) {
        if (p.nonfinite != nullptr && !(isfinite(a0[k]) && isfinite(a1[k]))) atomicAdd(p.nonfinite, 1);
        const float p0 = 1.0f / (1.0f + expf(-a0[k])), p1 = 1.0f / (1.0f + expf(-a1[k]));
        p.probs[2 * (s0 + k)] = p0;
        p.probs[2 * (s0 + k) + 1] = p1;
        p.pred[s0 + k] = (p1 > p0) ? 1 : 0;     // argmax, ties -> 0
      }
    }
  }
}

}  // namespace dsb
