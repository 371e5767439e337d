// Experiment: can a K-major SWIZZLE_64B UMMA A-operand descriptor start at a row that is not a multiple of 8
// (start address = tile base + s*64 bytes), with or without the descriptor's base_offset field?
// If yes, k-tap convolutions can read ONE staged activation tile at s = tap offsets instead of one TMA box per tap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I deepsignal_b200/csrc profiles/experiments/umma_shift.cu -o /tmp/umma_shift && /tmp/umma_shift
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace dsb::tc;

constexpr int ROWS = 144, K = 32, N = 32, M = 128;

__global__ void __launch_bounds__(128) shift_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 16384, bar = base + 20480, slot = base + 20496;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x;
  // fill A (ROWS x 32) and B (32 x 32) with the SWIZZLE_64B pattern TMA would produce (tile base 1024-aligned)
  for (int i = tid; i < ROWS * 4; i += 128) {
    const int r = i / 4, c = i % 4;
    const uint4 v = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    *reinterpret_cast<uint4*>(gen + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;
  }
  for (int i = tid; i < N * 4; i += 128) {
    const int r = i / 4, c = i % 4;
    const uint4 v = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
    *reinterpret_cast<uint4*>(gen + 16384 + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;
  }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_addr = a_base + shift * 64;
    uint64_t ad = make_kmajor_desc<32>(a_addr);
    if (mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;           // base_offset = (start >> 7) & 7
    if (mode == 2) ad |= (uint64_t)(shift & 7) << 49;                     // base_offset = row shift
    const uint64_t bd = make_kmajor_desc<32>(b_base);
    for (int k = 0; k < 2; ++k) tc_mma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, k != 0 ? 1u : 0u);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  uint32_t r[2][16];
  tc_ld16(tmem + ((uint32_t)(warp * 32) << 16), r[0]);
  tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 16, r[1]);
  tc_wait_ld();
  for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + j] = __uint_as_float(r[j / 16][j % 16]);
  tc_fence_before();
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> hD(M * N);
  for (int mode = 0; mode < 3; ++mode) {
    for (int shift = 0; shift <= 9; ++shift) {
      cudaMemset(dD, 0, M * N * 4);
      shift_kernel<<<1, 128, 32768>>>(dA, dB, dD, shift, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
      int bad = 0, bad_rows = 0;
      for (int m = 0; m < M; ++m) {
        int rb = 0;
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < K; ++k) ref += fA[(m + shift) * K + k] * fB[n * K + k];
          if (fabsf(ref - hD[m * N + n]) > 1e-3f) { ++bad; rb = 1; }
        }
        bad_rows += rb;
      }
      printf("mode %d (base_offset %s) shift %d rows: %s (%d wrong values, %d wrong rows)\n", mode,
             mode == 0 ? "0" : mode == 1 ? "(addr>>7)&7" : "shift", shift, bad ? "MISMATCH" : "ok", bad, bad_rows);
    }
  }
  return 0;
}
