// Experiment: why are my cta_group::2 kernels slower than their single-CTA versions?  A bare TMA -> tcgen05.mma ring
// (no epilogue work: one warp hands the accumulator straight back) in four variants, all 148 SMs busy:
//   0  single CTA, M=128 N=256, 48 KB per stage, 4 stages                       (gemm_tc.cu / lstm_seq.cu)
//   1  CTA pair,  M=256 N=256, 32 KB per stage per CTA, 6 stages, both CTAs' TMA bytes counted on the LEADER's barrier
//      (.cta_group::2 loads + remote arrive)                                     (gemm_tc2.cu / lstm_seq2.cu)
//   2  CTA pair, every CTA's loads complete on its OWN barrier; a relay thread in the peer forwards one remote arrive
//      per stage to a second barrier of the leader
//   3  as 1 with N=128 (half the cross-SM B-operand traffic per MMA)
//   4  as 1 with test_wait polling;  5-9 raw MMA / commit rates;  10-13 rings without MMAs;
//   15 / 16  as 13 / 1 with the producer's remote arrive demoted to .relaxed.cluster  <- the fix:
//   RESULT (B200): 0: 518 cycles per k-block, 1: 1340, 2: 1338, 4: 1334, 5/7/8/9: 514, 10: 1310, 11: 427, 12: 426, 13: 1319,
//   15: 378, 16: see profiles/r1_ablation.md.  `mbarrier.arrive.release.cluster` issued by the TMA-producer thread drains that
//   thread's outstanding bulk loads, which collapses the ring to depth 1; the arrive carries no data, .relaxed is enough.
// Prints cycles per 64-wide k-block (ideal: 4 MMAs x 128 clk = 512).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I deepsignal_b200/csrc profiles/experiments/pair_ring.cu -o /tmp/pair_ring -lcuda && /tmp/pair_ring
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace dsb::tc;

static void encode(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

struct alignas(64) Maps { CUtensorMap a, b; };

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
__device__ __forceinline__ void arrive_cluster(uint32_t a) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void arrive_cluster_relaxed(uint32_t a) { asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void tma_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2sm_local(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

constexpr int NKB = 8;

// non-blocking poll (mbarrier.test_wait) instead of the potentially-suspending try_wait
__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
#define WAIT(bar, par) do { if (MODE == 4) spin_wait(bar, par); else mbar_wait(bar, par); } while (0)

// ---------------------------------------------------------------- variant 0: single CTA
template <int RAW>
__global__ void __launch_bounds__(128, 1) ring1(const __grid_constant__ Maps maps, int tiles, int m_tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int ST = 4;
  constexpr uint32_t SB = 49152;
  const uint32_t bars = base + ST * SB;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (ST + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * ST + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * ST + 2 + a); };
  const uint32_t done = bars + 8u * (2 * ST + 4);
  const uint32_t slot = bars + 8u * (2 * ST + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    for (int s = 0; s < ST; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const int row0 = ((blockIdx.x + t * 7) % m_tiles) * 128, n0 = (t & 3) * 256;
      for (int kb = 0; kb < NKB; ++kb) {
        if (RAW == 1 && (t > 0 || kb >= ST)) continue;          // RAW 1: fill the ring once, then the MMAs re-read it; RAW 2: ring without MMAs
        mbar_wait(empty(stage), ph ^ 1u);
        mbar_expect_tx(full(stage), SB);
        tma_load_2d(base + stage * SB, &maps.a, full(stage), kb * 64, row0);
        tma_load_2d(base + stage * SB + 16384, &maps.b, full(stage), kb * 64, n0);
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0, acc = 0; uint32_t ph = 0, aph = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(tempty(acc), aph ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < NKB; ++kb) {
        if (RAW != 1 || (t == 0 && kb < ST)) mbar_wait(full(stage), ph);
        tc_fence_after();
        const uint64_t ad = make_kmajor_desc<64>(base + stage * SB), bd = make_kmajor_desc<64>(base + stage * SB + 16384);
        if (RAW != 2) for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem + acc * 256, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        if (RAW != 1) tc_commit(empty(stage));
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
      tc_commit(tfull(acc));
      acc ^= 1; if (!acc) aph ^= 1u;
    }
    tc_commit(done);
    mbar_wait(done, 0);
    out[blockIdx.x] = clock64() - t0;
  } else if (warp == 2 && lane == 0) {
    int acc = 0; uint32_t aph = 0;
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(tfull(acc), aph);
      mbar_arrive(tempty(acc));
      acc ^= 1; if (!acc) aph ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

// variant 12: the single-CTA ring (no MMAs), merely LAUNCHED as a cluster of two
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) ring1c(const __grid_constant__ Maps maps, int tiles, int m_tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int ST = 4;
  constexpr uint32_t SB = 49152;
  const uint32_t bars = base + ST * SB;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (ST + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const int row0 = ((blockIdx.x + t * 7) % m_tiles) * 128, n0 = (t & 3) * 256;
      for (int kb = 0; kb < NKB; ++kb) {
        mbar_wait(empty(stage), ph ^ 1u);
        mbar_expect_tx(full(stage), SB);
        tma_load_2d(base + stage * SB, &maps.a, full(stage), kb * 64, row0);
        tma_load_2d(base + stage * SB + 16384, &maps.b, full(stage), kb * 64, n0);
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t)
      for (int kb = 0; kb < NKB; ++kb) {
        mbar_wait(full(stage), ph);
        mbar_arrive(empty(stage));
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  cluster_sync_all();
}

// ---------------------------------------------------------------- variants 1..3: CTA pair
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) ring2(const __grid_constant__ Maps maps, int tiles, int m_tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int ST = 6;
  constexpr int BN = MODE == 3 ? 128 : 256;              // MODE 4 = MODE 1 with test_wait polling              // MMA N; every CTA stages BN/2 weight rows
  constexpr uint32_t BB = (BN / 2) * 128;
  constexpr uint32_t SB = 16384 + BB;
  const uint32_t bars = base + ST * 32768;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (ST + s); };
  auto pfull = [&](int s) { return bars + 8u * (2 * ST + s); };          // MODE 2: "the peer's stage is full" (leader side)
  auto tfull = [&](int a) { return bars + 8u * (3 * ST + a); };
  auto tempty = [&](int a) { return bars + 8u * (3 * ST + 2 + a); };
  const uint32_t done = bars + 8u * (3 * ST + 4);
  const uint32_t slot = bars + 8u * (3 * ST + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    for (int s = 0; s < ST; ++s) { mbar_init(full(s), MODE == 2 ? 1 : 2); mbar_init(empty(s), 1); mbar_init(pfull(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
  const int cid = blockIdx.x >> 1;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const int row0 = (((cid + t * 7) * 2 + (int)rank) % m_tiles) * 128, n0 = (t & 3) * 256 + (int)rank * (BN / 2);
      for (int kb = 0; kb < NKB; ++kb) {
        if (MODE >= 5 && MODE != 10 && MODE != 13 && MODE != 15 && MODE != 16 && (t > 0 || kb >= ST)) continue;
        WAIT(empty(stage), ph ^ 1u);
        const uint32_t dst = base + stage * 32768;
        if (MODE == 2) {
          mbar_expect_tx(full(stage), SB);
          tma_load_2d(dst, &maps.a, full(stage), kb * 64, row0);
          tma_load_2d(dst + 16384, &maps.b, full(stage), kb * 64, n0);
        } else {
          const uint32_t lfull = map_to_cta(full(stage), 0);
          if (leader) mbar_expect_tx(full(stage), 2 * SB); else if (MODE == 15 || MODE == 16) arrive_cluster_relaxed(lfull); else arrive_cluster(lfull);
          tma_2sm(dst, &maps.a, lfull, kb * 64, row0);
          tma_2sm(dst + 16384, &maps.b, lfull, kb * 64, n0);
        }
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    int stage = 0, acc = 0; uint32_t ph = 0, aph = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      WAIT(tempty(acc), aph ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < NKB; ++kb) {
        if (MODE < 5 || MODE == 10 || MODE == 13 || MODE == 15 || MODE == 16 || (t == 0 && kb < ST)) WAIT(full(stage), ph);
        if (MODE == 2) WAIT(pfull(stage), ph);
        tc_fence_after();
        const uint64_t ad = make_kmajor_desc<64>(base + stage * 32768), bd = make_kmajor_desc<64>(base + stage * 32768 + 16384);
        if (MODE != 10 && MODE != 13 && MODE != 15) for (int k = 0; k < 4; ++k) mma_2sm(tmem + acc * 256, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        if (MODE < 5 || MODE == 10 || MODE == 16) commit_2sm(empty(stage), 3);
        if (MODE == 13) { mbar_arrive(empty(stage)); arrive_cluster(map_to_cta(empty(stage), 1)); }
        if (MODE == 15) { mbar_arrive(empty(stage)); arrive_cluster_relaxed(map_to_cta(empty(stage), 1)); }
        if (MODE == 7) commit_2sm(pfull(stage), 3);          // dummy barrier, never waited on
        if (MODE == 8) commit_2sm(pfull(stage), 1);
        if (MODE == 9) commit_2sm_local(pfull(stage));
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
      commit_2sm(tfull(acc), 3);
      acc ^= 1; if (!acc) aph ^= 1u;
    }
    commit_2sm(done, 1);
    WAIT(done, 0);
    out[cid] = clock64() - t0;
  } else if (warp == 2 && lane == 0) {
    int acc = 0; uint32_t aph = 0;
    const uint32_t lt[2] = {map_to_cta(tempty(0), 0), map_to_cta(tempty(1), 0)};
    for (int t = 0; t < tiles; ++t) {
      WAIT(tfull(acc), aph);
      arrive_cluster(lt[acc]);
      acc ^= 1; if (!acc) aph ^= 1u;
    }
  } else if (MODE == 2 && warp == 3 && lane == 0 && !leader) {
    // relay: my stage is full -> one remote arrive on the leader's pfull
    int stage = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t)
      for (int kb = 0; kb < NKB; ++kb) {
        WAIT(full(stage), ph);
        arrive_cluster(map_to_cta(pfull(stage), 0));
        if (++stage == ST) { stage = 0; ph ^= 1u; }
      }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main() {
  cuInit(0);
  const int m_tiles = 296, rows = m_tiles * 128, K = 512, N = 1024, tiles = 400;
  std::vector<__nv_bfloat16> hA((size_t)rows * K), hB((size_t)N * K);
  srand(1);
  for (auto& v : hA) v = __float2bfloat16((rand() % 255 - 127) / 64.0f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 255 - 127) / 256.0f);
  __nv_bfloat16 *dA, *dB;
  long long* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 148 * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  Maps m1, m2, m3;
  encode(&m1.a, dA, rows, K, 128, 64); encode(&m1.b, dB, N, K, 256, 64);
  encode(&m2.a, dA, rows, K, 128, 64); encode(&m2.b, dB, N, K, 128, 64);
  encode(&m3.a, dA, rows, K, 128, 64); encode(&m3.b, dB, N, K, 64, 64);
  const size_t smem = 1024 + 6 * 32768 + 512;
  cudaFuncSetAttribute(ring1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring1c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring1<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ring2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<long long> h(148);
  for (int rep = 0; rep < 1; ++rep)
    for (int mode = 0; mode < 17; ++mode) {
      if ((mode >= 2 && mode <= 9) || (mode >= 10 && mode <= 14)) continue;
      cudaMemset(dOut, 0, 148 * 8);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) ring1<0><<<148, 128, smem>>>(m1, tiles, m_tiles, dOut);
      else if (mode == 6) ring1<1><<<148, 128, smem>>>(m1, tiles, m_tiles, dOut);
      else if (mode == 5) ring2<5><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 12) ring1c<<<148, 128, smem>>>(m1, tiles, m_tiles, dOut);
      else if (mode == 16) ring2<16><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 15) ring2<15><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 13) ring2<13><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 10) ring2<10><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 11) ring1<2><<<148, 128, smem>>>(m1, tiles, m_tiles, dOut);
      else if (mode == 7) ring2<7><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 8) ring2<8><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 9) ring2<9><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 1) ring2<1><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 2) ring2<2><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      else if (mode == 3) ring2<3><<<148, 128, smem>>>(m3, tiles, m_tiles, dOut);
      else ring2<4><<<148, 128, smem>>>(m2, tiles, m_tiles, dOut);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h.data(), dOut, 148 * 8, cudaMemcpyDeviceToHost);
      const int nctr = (mode == 0 || mode == 6 || mode == 11 || mode == 12) ? 148 : 74;
      double avg = 0; long long mx = 0;
      for (int i = 0; i < nctr; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
      avg /= nctr;
      const double rows_per_kb = (mode == 0 || mode == 6) ? 128.0 : 256.0;
      printf("mode %d: %s  %.3f ms  cycles per k-block: avg %.0f max %.0f  (per 128 rows per SM: %.0f)  TFLOP/s %.0f\n", mode, cudaGetErrorString(err), ms,
             avg / (tiles * NKB), (double)mx / (tiles * NKB), avg / (tiles * NKB) / (rows_per_kb / 128.0) * (mode == 0 ? 1.0 : 2.0) / (mode == 3 ? 0.5 : 1.0),
             (double)((mode == 0 || mode == 6) ? 148 : 74) * tiles * NKB * rows_per_kb * (mode == 3 ? 128.0 : 256.0) * 64 * 2 / (ms * 1e-3) / 1e12);
    }
  return 0;
}
