// gemm_f32.cu — CUDA-core fp32 GEMM with the conv / LSTM epilogues (the DSB_PREC_FP32 path).
//
// 128x64 output tile per CTA, 256 threads, 8x4 outputs per thread, BK = 16, register-staged double
// buffering of the next K block.  This is the parity path (1e-4 abs against the float64 oracle) and
// the on-device reference for the tcgen05 path; it is fp32-FMA bound, not a tensor-core kernel.
#include "ops.cuh"

namespace dsb {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

template <int EPI>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(const GemmF32Params p) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int N = p.epi.N;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader coordinates
  const int a_row[2] = {tid >> 2, (tid + NT) >> 2};
  const int a_kq = tid & 3;
  const int b_n = tid >> 2, b_kq = tid & 3;

  int total_kb = 0;
  for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].K / BK;

  float4 ra[2], rb;
  auto fetch = [&](int kb) {
    int s = 0, kloc = kb, koff = 0;
    while (kloc >= p.seg[s].K / BK) { kloc -= p.seg[s].K / BK; koff += p.seg[s].K; ++s; }
    const SegF32& sg = p.seg[s];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + a_row[h] + sg.drow;
      if (r >= 0 && r < p.rows_valid)
        ra[h] = *reinterpret_cast<const float4*>(sg.a + (size_t)r * sg.ld + sg.col0 + kloc * BK + a_kq * 4);
      else
        ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (n0 + b_n < N)
      rb = *reinterpret_cast<const float4*>(p.B + (size_t)(n0 + b_n) * p.Ktot + koff + kloc * BK + b_kq * 4);
    else
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[buf][a_kq * 4 + 0][a_row[h]] = ra[h].x;
      As[buf][a_kq * 4 + 1][a_row[h]] = ra[h].y;
      As[buf][a_kq * 4 + 2][a_row[h]] = ra[h].z;
      As[buf][a_kq * 4 + 3][a_row[h]] = ra[h].w;
    }
    Bs[buf][b_kq * 4 + 0][b_n] = rb.x;
    Bs[buf][b_kq * 4 + 1][b_n] = rb.y;
    Bs[buf][b_kq * 4 + 2][b_n] = rb.z;
    Bs[buf][b_kq * 4 + 3][b_n] = rb.w;
  };

  fetch(0);
  stash(0);
  __syncthreads();
  for (int kb = 0; kb < total_kb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < total_kb) fetch(kb + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < total_kb) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue -------------------------------------------------------------------------------
  const EpiParams& e = p.epi;
  const int col = n0 + tx * 4;
  if (col >= N) return;
  const float4 bias = *reinterpret_cast<const float4*>(e.bias + col);
  if constexpr (EPI == EPI_CONV) {
    const bool relu = (e.relu_mask >> (col >> 4)) & 1u;
    const bool first = col < e.split;
    float* dst = first ? reinterpret_cast<float*>(e.d0) : reinterpret_cast<float*>(e.d1);
    const int ld = first ? e.ld0 : e.ld1;
    const int dc = first ? e.c0 + col : e.c1 + col - e.split;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + ty * 8 + i;
      if (r >= e.M) break;
      float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
      if (e.res != nullptr && first) {
        const float4 rs = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.res) + (size_t)r * e.res_ld + e.res_c0 + col);
        v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w;
      }
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (e.Lp > 0 && (r % e.Lp) >= e.L) v = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + (size_t)r * ld + dc) = v;
    }
  } else {
    const int u = col >> 2;
    float* hout = reinterpret_cast<float*>(e.h_out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + ty * 8 + i;
      if (r >= e.M) break;
      float cn, hn;
      lstm_unit<false>(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w,
                       e.c[(size_t)r * HID + u], cn, hn);
      e.c[(size_t)r * HID + u] = cn;
      hout[(size_t)r * HID + u] = hn;
    }
  }
}

}  // namespace

void launch_gemm_f32(const GemmF32Params& p, int epi_kind, cudaStream_t s) {
  dim3 grid((p.epi.M + BM - 1) / BM, (p.epi.N + BN - 1) / BN);
  if (epi_kind == EPI_CONV)
    gemm_f32_kernel<EPI_CONV><<<grid, NT, 0, s>>>(p);
  else
    gemm_f32_kernel<EPI_LSTM><<<grid, NT, 0, s>>>(p);
  DSB_CUDA(cudaGetLastError());
}

}  // namespace dsb
