// gemm_tc2.cu — the LSTM step GEMM on CTA pairs (tcgen05.mma cta_group::2): one 256x256 output tile per pair of
// SMs, each CTA stages its own 128 activation rows and HALF of the weight tile, so the weight bytes every SM has
// to ingest per output row are halved (the single-CTA kernel in gemm_tc.cu is bound by operand ingest, see
// profiles/r1_ablation.md).  Same segments / epilogue / cell-state layout as gemm_tc_kernel<64, EPI_LSTM>.
//
// Protocol (as in CUTLASS' 2-SM kernels): cluster (2,1,1); both CTAs allocate TMEM with cta_group::2; both issue
// their TMA loads with .cta_group::2 so the bytes are counted on the LEADER's (rank 0) "full" barrier; the leader
// alone issues the MMAs and commits with .multicast::cluster to the "empty" / "accumulator full" barriers of both
// CTAs; the epilogue warps of both CTAs arrive on the leader's "accumulator empty" barrier.
#include <cstdlib>

#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace dsb {
namespace {

using namespace tc;

constexpr int T2_THREADS = 320;
constexpr int T2_STAGES = 6;
constexpr uint32_t T2_A_BYTES = 128 * 64 * 2, T2_B_BYTES = 128 * 64 * 2, T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // .relaxed: a .release.cluster arrive from the TMA-producer thread drains its outstanding bulk loads (profiles/experiments/pair_ring.cu)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

template <bool F16> __device__ __forceinline__ uint32_t pack16(float a, float b) {
  if constexpr (F16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}

// maps.a[0..1]: A segment matrices (box 128 x 64); maps.b: weights [1024][K] with a 128-row box.
template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
gemm_tc2_lstm_kernel(const __grid_constant__ TcMaps maps, const GemmTCParams p, const int m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const uint32_t bar_base = smem_base + T2_STAGES * T2_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (T2_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * T2_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * T2_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * T2_STAGES + 4);
  float* sbias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));
  const uint32_t stage_base = (tmem_slot + 16u + (uint32_t)GATES * 4u + 15u) & ~15u;
  int total_kb = 0;
  for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;
  const int n_tiles_n = GATES / 256;                                  // 4
  const int m_pairs = (m_tiles + 1) / 2;
  const int num_tiles = m_pairs * n_tiles_n;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < GATES; i += T2_THREADS) sbias[i] = p.epi.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait_prior_grid();

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int row0 = ((tile / n_tiles_n) * 2 + (int)rank) * 128;
        const int nrow0 = (tile % n_tiles_n) * 256 + (int)rank * 128;
        int kb_global = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const SegTC sg = p.seg[s];
          for (int j = 0; j < sg.nkb; ++j) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t lfull = map_to_cta(full_bar(stage), 0);        // the leader's barrier counts both CTAs' bytes
            if (leader) mbar_expect_tx(full_bar(stage), 2 * T2_STAGE_BYTES);
            else mbar_arrive_cluster(lfull);
            const uint32_t a_dst = smem_base + stage * T2_STAGE_BYTES;
            tma_load_2d_2sm(a_dst, &maps.a[sg.map], lfull, sg.col0 + j * 64, row0 + sg.drow);
            tma_load_2d_2sm(a_dst + T2_A_BYTES, &maps.b, lfull, kb_global * 64, nrow0);
            ++kb_global;
            if (++stage == T2_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    if (lane == 0 && leader) {
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * T2_STAGE_BYTES;
          const uint64_t adesc = make_kmajor_desc<64>(a_addr);
          const uint64_t bdesc = make_kmajor_desc<64>(a_addr + T2_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit_2sm(empty_bar(stage), 3);
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2sm(tfull_bar(acc), 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ---------------- epilogue warps (both CTAs): LSTM cell update, as gemm_tc_kernel<64, EPI_LSTM> ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const EpiParams& e = p.epi;
    const uint32_t stage = stage_base + (uint32_t)(warp - 2) * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int tm = (tile / n_tiles_n) * 2 + (int)rank;             // this CTA's 128-row tile
      const int n0 = (tile % n_tiles_n) * 256;
      const int row = tm * 128 + q * 32 + lane;
      const int wrow0 = tm * 128 + q * 32;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      const bool row_ok = row < e.M;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      if (tm < m_tiles) {
        for (int c = half * 64; c < 256; c += 128) {
          uint32_t r[4][16];
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) tc_ld16(t_row + (uint32_t)(c + 16 * i), r[i]);
          const int col = n0 + c;
          const int u0 = col >> 2;
          float4* cblk = reinterpret_cast<float4*>(e.c) + ((((size_t)tm * (GATES / 64) + (size_t)(col >> 6)) * 4 + q) * 4) * 32 + lane;
          float4 cprev[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) cprev[i] = cblk[i * 32];
          tc_wait_ld();
          float hn[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float cpv[4] = {cprev[i].x, cprev[i].y, cprev[i].z, cprev[i].w};
            float cn[4];
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + col + 16 * i + 4 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + col + 16 * i + 4 * j + 4);
              const float za[4] = {__uint_as_float(r[i][4 * j + 0]) + b0.x, __uint_as_float(r[i][4 * j + 1]) + b0.y,
                                   __uint_as_float(r[i][4 * j + 2]) + b0.z, __uint_as_float(r[i][4 * j + 3]) + b0.w};
              const float zb[4] = {__uint_as_float(r[i][4 * j + 4]) + b1.x, __uint_as_float(r[i][4 * j + 5]) + b1.y,
                                   __uint_as_float(r[i][4 * j + 6]) + b1.z, __uint_as_float(r[i][4 * j + 7]) + b1.w};
              lstm_unit_pair_fast(za, zb, cpv[j], cpv[j + 1], cn[j], cn[j + 1], hn[4 * i + j], hn[4 * i + j + 1]);
            }
            cblk[i * 32] = make_float4(cn[0], cn[1], cn[2], cn[3]);
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack16<F16>(hn[2 * i], hn[2 * i + 1]);
          uint8_t* hdst = reinterpret_cast<uint8_t*>(e.h_out) + ((ptrdiff_t)wrow0 * HID + u0) * 2;
          staged_store<2>(stage, pk, lane, ok_mask, hdst, (size_t)HID * 2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(acc), 0));   // the leader's MMA issuer waits for both CTAs
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // no CTA leaves while its partner may still address its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// L: a finalized single-CTA LSTM launch (segments, epilogue); b_half: weights map with a 128-row box.
void launch_gemm_tc2_lstm(const TcLaunch& L, const CUtensorMap& b_half, int M, int sm_count, cudaStream_t s) {
  const int m_tiles = (M + 127) / 128;
  if (m_tiles <= 0) return;
  TcLaunch Lr = L;
  Lr.p.epi.M = M;
  Lr.maps.b = b_half;
  const size_t smem = 1024 + (size_t)T2_STAGES * T2_STAGE_BYTES + 8 * (2 * T2_STAGES + 4) + 32 + GATES * 4 + 16 + 8 * 2048;
  const int num_tiles = ((m_tiles + 1) / 2) * (GATES / 256);
  int grid = 2 * std::min(num_tiles, sm_count / 2);
  auto kern = L.fp16 ? gemm_tc2_lstm_kernel<true> : gemm_tc2_lstm_kernel<false>;
  static bool configured[2][64] = {{false}};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  if (!configured[L.fp16 ? 1 : 0][dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[L.fp16 ? 1 : 0][dev & 63] = true;
  }
  launch_pdl(kern, grid, T2_THREADS, smem, s, Lr.maps, Lr.p, m_tiles);
}

}  // namespace dsb
