// Experiment: may one kernel mix tcgen05.mma.cta_group::2 (issued by the leader for both CTAs) with cta_group::1 MMAs
// that every CTA of the pair issues for itself, on TMEM allocated with cta_group::2?  (ptxas accepts it.)  Needed by the
// planned single-kernel inception block: TMA-fed GEMMs share their weights over the pair, thread-fed GEMMs stay CTA-local.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I deepsignal_b200/csrc profiles/experiments/mix_cta_group.cu -o /tmp/mix && /tmp/mix
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace dsb::tc;

constexpr int M = 128, K = 32, N = 64;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// A: [2][128][32], B: [2][64][32] (full local B of each CTA), Bh: [2][32][32] (this CTA's half of the pair's B)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mix_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* Bh,
                                                                            float* D1, float* D2, int order) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 8192, bh_base = base + 12288, bar1 = base + 14336, bar2 = base + 14344, slot = base + 14352;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_rank();
  auto fill = [&](const __nv_bfloat16* src, int rows, uint32_t off) {
    for (int i = tid; i < rows * 4; i += 128) {
      const int r = i / 4, c = i % 4;
      const uint4 v = *reinterpret_cast<const uint4*>(src + r * K + c * 8);
      *reinterpret_cast<uint4*>(gen + off + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;     // SWIZZLE_64B
    }
  };
  fill(A + rank * M * K, M, 0);
  fill(B + rank * N * K, N, 8192);
  fill(Bh + rank * (N / 2) * K, N / 2, 12288);
  if (tid == 0) { mbar_init(bar1, 1); mbar_init(bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
  if (tid == 0) {
    const uint32_t id1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t id2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t ad = make_kmajor_desc<32>(a_base), bd = make_kmajor_desc<32>(b_base), bhd = make_kmajor_desc<32>(bh_base);
    auto pair = [&]() {
      if (rank == 0) {
        for (int k = 0; k < 2; ++k)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(ad + 2 * k), "l"(bhd + 2 * k), "r"(id2), "r"((uint32_t)(k != 0)) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar2), "h"((uint16_t)3) : "memory");
      }
    };
    auto local = [&]() {
      for (int k = 0; k < 2; ++k) tc_mma_bf16(tmem + 64, ad + 2 * k, bd + 2 * k, id1, k != 0 ? 1u : 0u);
      tc_commit(bar1);
    };
    if (order == 0) { pair(); local(); } else { local(); pair(); }
  }
  mbar_wait(bar1, 0);
  mbar_wait(bar2, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int part = 0; part < 2; ++part) {
    float* out = (part == 0 ? D2 : D1) + (size_t)rank * M * N;
    for (int g = 0; g < 4; ++g) {
      uint32_t r[16];
      tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + part * 64 + g * 16, r);
      tc_wait_ld();
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + g * 16 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(2 * M * K), hB(2 * N * K), hBh(2 * (N / 2) * K);
  std::vector<float> fA(hA.size()), fB(hB.size()), fBh(hBh.size());
  srand(3);
  auto rnd = [&](std::vector<__nv_bfloat16>& h, std::vector<float>& f, int div) {
    for (size_t i = 0; i < h.size(); ++i) { float v = (rand() % 17 - 8) / (float)div; h[i] = __float2bfloat16(v); f[i] = __bfloat162float(h[i]); }
  };
  rnd(hA, fA, 8); rnd(hB, fB, 4); rnd(hBh, fBh, 4);
  __nv_bfloat16 *dA, *dB, *dBh; float *dD1, *dD2;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dBh, hBh.size() * 2);
  cudaMalloc(&dD1, 2 * M * N * 4); cudaMalloc(&dD2, 2 * M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBh, hBh.data(), hBh.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> h1(2 * M * N), h2(2 * M * N);
  for (int order = 0; order < 2; ++order) {
    cudaMemset(dD1, 0, h1.size() * 4); cudaMemset(dD2, 0, h2.size() * 4);
    mix_kernel<<<2, 128, 32768>>>(dA, dB, dBh, dD1, dD2, order);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h1.data(), dD1, h1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), dD2, h2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int c = 0; c < 2; ++c)
      for (int i = 0; i < M; ++i)
        for (int n = 0; n < N; ++n) {
          double r1 = 0, r2 = 0;
          for (int k = 0; k < K; ++k) {
            r1 += (double)fA[(c * M + i) * K + k] * fB[(c * N + n) * K + k];
            r2 += (double)fA[(c * M + i) * K + k] * fBh[((n / 32) * 32 + n % 32) * K + k];      // pair B = [half of CTA 0 ; half of CTA 1]
          }
          e1 = fmax(e1, fabs(r1 - h1[(c * M + i) * N + n]));
          e2 = fmax(e2, fabs(r2 - h2[(c * M + i) * N + n]));
        }
    printf("order %d (%s first): %s  max err local cta_group::1 MMA %.3g, pair cta_group::2 MMA %.3g\n", order, order == 0 ? "pair" : "local",
           cudaGetErrorString(e), e1, e2);
  }
  return 0;
}
