// lstm_seq2.cu — the time-persistent BiLSTM kernel of lstm_seq.cu on CTA pairs (tcgen05.mma cta_group::2).
//
// lstm_seq_kernel is bound by L2 -> SM operand ingest (48 KB per 128x256x64 MMA block, ~17 TB/s over the chip).  Here a
// cluster of two CTAs works on two neighbouring 128-sample tiles of the same direction in lock step: every CTA stages its
// own activation rows and HALF of the 256-row weight block, the leader issues M=256 MMAs that read both halves, so a
// CTA ingests 32 KB per block instead of 48 KB.  Everything else (walk order, cell state in L2, h via TMA store, two
// pair-items interleaved per cluster, 16 epilogue warps) is lstm_seq.cu; results are bit-identical to it.
//
// Protocol (gemm_tc2.cu): both CTAs allocate TMEM with cta_group::2 and issue their TMA loads with .cta_group::2 so the
// bytes are counted on the LEADER's "full" barrier; the leader alone issues the MMAs and commits with
// .multicast::cluster to the "empty" / "accumulator full" barriers of both CTAs; the epilogue warps of both CTAs arrive
// on the leader's "accumulator empty" barrier.  The h_t hand-off (epilogue -> own TMA producer) stays CTA-local.
#include <cstdio>
#include <cstdlib>

#include "lstm_seq.cuh"
#include "tc_common.cuh"
#include "lstm_epi.cuh"

namespace dsb {
namespace {

using namespace tc;

constexpr int P_EW = 16;                                  // epilogue warps
constexpr int P_THREADS = 64 + 32 * P_EW + 32;     // + one warp that issues the h_t TMA stores
constexpr int P_STAGES = 5;
constexpr uint32_t P_A_BYTES = 128 * 64 * 2, P_B_BYTES = 128 * 64 * 2, P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr uint32_t P_HTILE_BYTES = 128 * 64 * 2;
constexpr int P_CHUNKS = GATES / 256;
constexpr size_t P_SMEM = 1024 + (size_t)P_STAGES * P_STAGE_BYTES + 2 * P_HTILE_BYTES + 256 + 2 * GATES * sizeof(float);

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
// Remote arrives are .relaxed: they carry no data (TMA bytes are tracked by complete_tx, TMEM reads are complete before the
// arrive is issued).  A .release.cluster arrive issued by the TMA-producer thread drains that thread's outstanding bulk
// loads and collapses the ring to depth 1 -- 1340 instead of 512 cycles per k-block (profiles/experiments/pair_ring.cu).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
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

// (cluster pair-of-items, layer, t, slot) walk shared by the roles.  A pair-item is (direction, two neighbouring M tiles).
template <class Body>
__device__ __forceinline__ void pair_walk(const LstmSeqParams& p, Body&& body) {
  const int m_tiles = (p.n + 127) / 128;
  const int m_pairs = (m_tiles + 1) / 2;
  const int pitems = 2 * m_pairs;
  const int cpairs = (pitems + 1) / 2;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  for (int cp = cluster_id; cp < cpairs; cp += n_clusters)
    for (int l = 0; l < LSTM_SEQ_LAYERS; ++l)
      for (int t = 0; t < p.KL; ++t)
        for (int s = 0; s < 2; ++s) {
          const int pitem = 2 * cp + s;
          if (pitem >= pitems) continue;
          const int d = pitem / m_pairs, mp = pitem - d * m_pairs;
          body(cp, l, t, s, d, mp);
        }
}

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
lstm_seq2_kernel(const __grid_constant__ LstmSeqMaps maps, const LstmSeqParams p, const int ablate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const uint32_t ring_base = smem_base;
  const uint32_t htile = ring_base + P_STAGES * P_STAGE_BYTES;
  const uint32_t bar_base = htile + 2 * P_HTILE_BYTES;            // two h tiles: one being stored while the next is written
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (P_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * P_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * P_STAGES + 2 + a); };
  auto hready_bar = [&](int s) { return bar_base + 8u * (2 * P_STAGES + 4 + s); };
  auto hfull_bar = [&](int b) { return bar_base + 8u * (2 * P_STAGES + 6 + b); };
  auto hfree_bar = [&](int b) { return bar_base + 8u * (2 * P_STAGES + 8 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * P_STAGES + 10);
  float* sbias = reinterpret_cast<float*>(smem_raw + ((bar_base + 256u) - smem_u32(smem_raw)));     // [2 slots][1024]

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * P_EW); mbar_init(hready_bar(a), 1);
      mbar_init(hfull_bar(a), P_EW); mbar_init(hfree_bar(a), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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

  const int cap = (int)p.cap;
  const int m_tiles = (p.n + 127) / 128;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t steps_done[2] = {0u, 0u};
      pair_walk(p, [&](int, int l, int t, int s, int d, int mp) {
        int mt = 2 * mp + (int)rank;
        if (mt >= m_tiles) mt = m_tiles - 1;                  // odd tail: the partner-less CTA re-reads a valid tile, stores nothing
        const int row0 = mt * 128;
        const int pos = (d == 0) ? t : p.KL - 1 - t;
        if (steps_done[s] > 0) mbar_wait(hready_bar(s), (steps_done[s] - 1u) & 1u);
        ++steps_done[s];
        const CUtensorMap* wmap = &maps.w2[d][l];
        for (int ch = 0; ch < P_CHUNKS; ++ch) {
          auto load = [&](const CUtensorMap* amap, int col, int row, int wkb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t lfull = map_to_cta(full_bar(stage), 0);      // the leader's barrier counts both CTAs' bytes
            if (leader) mbar_expect_tx(full_bar(stage), 2 * P_STAGE_BYTES);
            else mbar_arrive_cluster(lfull);
            const uint32_t a_dst = ring_base + stage * P_STAGE_BYTES;
            tma_load_2d_2sm(a_dst, amap, lfull, col, row);
            tma_load_2d_2sm(a_dst + P_A_BYTES, wmap, lfull, wkb * 64, ch * 256 + (int)rank * 128);
            if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
          };
          if (l == 0) {
            if (t > 0)
              for (int j = 0; j < 4; ++j) load(&maps.h[d][0], j * 64, t * cap + row0, j);
            load(&maps.f, 0, pos * cap + row0, 4);
          } else {
            for (int j = 0; j < 4; ++j) load(&maps.h[d][l - 1], j * 64, (t + 1) * cap + row0, j);
            if (t > 0)
              for (int j = 0; j < 4; ++j) load(&maps.h[d][l], j * 64, t * cap + row0, 4 + j);
          }
        }
      });
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && leader) {
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      pair_walk(p, [&](int, int l, int t, int, int, int) {
        const int nkb = (l == 0) ? (t > 0 ? 5 : 1) : (t > 0 ? 8 : 4);
        for (int ch = 0; ch < P_CHUNKS; ++ch) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = ring_base + stage * P_STAGE_BYTES;
            const uint64_t adesc = make_kmajor_desc<64>(a_addr);
            const uint64_t bdesc = make_kmajor_desc<64>(a_addr + P_A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit_2sm(empty_bar(stage), 3);
            if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit_2sm(tfull_bar(acc), 3);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      });
    }
  } else if (warp < 2 + P_EW) {
    // ================= epilogue warps (both CTAs) =================
    const int q = warp & 3;
    const int cs = (warp - 2) >> 2;           // 64-column group of every 256-column chunk
    const int et = threadIdx.x - 64;
    constexpr int ET = P_EW * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    int bias_cp = -1, bias_layer = -1;
    uint32_t hcount = 0;                      // chunks processed by this CTA (selects the h tile buffer)
    const uint32_t ltempty[2] = {map_to_cta(tempty_bar(0), 0), map_to_cta(tempty_bar(1), 0)};
    pair_walk(p, [&](int cp, int l, int t, int s, int d, int mp) {
      if (cp != bias_cp || l != bias_layer) {
        bias_cp = cp;
        bias_layer = l;
        named_bar_sync(1, ET);
        const int m_pairs = (m_tiles + 1) / 2;
        for (int ss = 0; ss < 2; ++ss) {
          const int pitem = 2 * cp + ss;
          if (pitem >= 2 * m_pairs) continue;
          const float* b = p.bias[pitem / m_pairs][l];
          for (int i = et; i < GATES; i += ET) sbias[ss * GATES + i] = b[i];
        }
        named_bar_sync(1, ET);
      }
      const float* sb = sbias + s * GATES;
      const int mt = 2 * mp + (int)rank;
      const bool mt_ok = mt < m_tiles;
      const bool c_in = t > 0 && mt_ok && !(ablate & 1), c_out = t + 1 < p.KL && mt_ok && !(ablate & 1);
      for (int ch = 0; ch < P_CHUNKS; ++ch) {
        const int col = ch * 256 + cs * 64;
        float4* cblk = reinterpret_cast<float4*>(p.c[d]) + ((((size_t)(mt_ok ? mt : 0) * (GATES / 64) + (size_t)(col >> 6)) * 4 + q) * 4) * 32 + lane;
        float4 c4[4];
        lstm_load_c(cblk, c_in, c4);                               // in flight while this warp waits for the accumulator
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
        uint32_t pk[8];
        lstm_epi_group<F16, true>(t_row + (uint32_t)(cs * 64), cblk, c_in, c_out, sb + col, pk, ablate, [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(ltempty[acc]);       // the leader's MMA issuer waits for both CTAs
        }, c4);
        // h_t of this chunk: each warp drops its [32 x 16] piece into the swizzled [128 x 64] tile of buffer hb and moves on;
        // the store warp ships the tile with one TMA store once all 16 pieces are there (no CTA-wide barrier on this path)
        {
          const int hb = (int)(hcount & 1u);
          mbar_wait(hfree_bar(hb), ((hcount >> 1) & 1u) ^ 1u);
          const uint32_t rr = (uint32_t)(q * 32 + lane);
          const uint32_t rowp = htile + (uint32_t)hb * P_HTILE_BYTES + rr * 128u;
          sts128(rowp + (((uint32_t)(2 * cs) ^ (rr & 7u)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          sts128(rowp + (((uint32_t)(2 * cs + 1) ^ (rr & 7u)) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(hfull_bar(hb));
          ++hcount;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    });
  } else {
    // ================= h_t store warp (both CTAs) =================
    if (lane == 0) {
      uint32_t hcount = 0;
      pair_walk(p, [&](int, int l, int t, int s, int d, int mp) {
        const int mt = 2 * mp + (int)rank;
        const bool mt_ok = mt < m_tiles;
        for (int ch = 0; ch < P_CHUNKS; ++ch) {
          const int hb = (int)(hcount & 1u);
          mbar_wait(hfull_bar(hb), (hcount >> 1) & 1u);
          if (mt_ok && !(ablate & 2)) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&maps.h[d][l])), "r"(htile + (uint32_t)hb * P_HTILE_BYTES), "r"(ch * 64), "r"((t + 1) * cap + mt * 128)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the tile has been read: the epilogue may rewrite it
          mbar_arrive(hfree_bar(hb));
          if (ch == P_CHUNKS - 1) {
            // publish h_t: stores complete (async proxy) -> this CTA's TMA producer may load them for the next step
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            mbar_arrive(hready_bar(s));
          }
          ++hcount;
        }
      });
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

void launch_lstm_seq2(const LstmSeqLaunch& L, int n, int sm_count, int ablate, cudaStream_t s) {
  if (n <= 0) return;
  const int m_tiles = (n + 127) / 128;
  const int m_pairs = (m_tiles + 1) / 2;
  const int cpairs = m_pairs;                     // 2 * m_pairs pair-items, two per cluster pass
  int clusters = sm_count / 2;
  if (clusters > cpairs) clusters = cpairs;
  if (clusters < 1) clusters = 1;
  LstmSeqParams p = L.p;
  p.n = n;
  static bool configured[64][2] = {{false}};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  auto kern = L.fp16 ? lstm_seq2_kernel<true> : lstm_seq2_kernel<false>;
  if (!configured[dev & 63][L.fp16 ? 1 : 0]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM));
    configured[dev & 63][L.fp16 ? 1 : 0] = true;
  }
  static int max_clusters[64] = {0};
  if (max_clusters[dev & 63] == 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (unsigned)(sm_count / 2));
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = P_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    DSB_CUDA(cudaOccupancyMaxActiveClusters(&nc, kern, &cfg));
    max_clusters[dev & 63] = nc > 0 ? nc : 1;
  }
  if (clusters > max_clusters[dev & 63]) clusters = max_clusters[dev & 63];
  launch_pdl(kern, 2 * clusters, P_THREADS, P_SMEM, s, L.maps, p, ablate);
}

}  // namespace dsb
