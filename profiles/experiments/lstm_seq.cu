// lstm_seq.cu — the whole 3-layer bidirectional LSTM of one chunk as ONE persistent tcgen05 kernel.
//
// The reference runs tf.nn.bidirectional_dynamic_rnn over 17 bases with three stacked LSTMCells per direction
// (layers.py:20-76).  A work item is (direction, 128-sample tile): every item is independent of every other item for
// all layers and time steps, so one CTA walks its items through layer -> time -> four 256-gate-column chunks and the
// step GEMM [x_t | h_{t-1}] · W is the same TMA -> tcgen05.mma -> TMEM pipeline as gemm_tc.cu, with
//   * the cell state of an item private to its CTA and re-read every ~10 us: it never leaves L2 (the per-step
//     launches streamed 67 MB of fp32 cell state through HBM per step), and needs no load at t = 0;
//   * h_t staged as a swizzled [128 x 64] shared-memory tile per chunk and written to H[dir][layer] slot t+1 with one TMA
//     store (whole 128-byte lines); the same CTA's TMA producer reads it back for step t+1 once the store group has
//     completed (bulk wait_group -> mbarrier);
//   * two items interleaved per CTA so the h_t round trip of one item hides behind the MMAs of the other;
//   * 102 launches per chunk -> 1.
// Segment order and K layout are those of the per-step path, so results are bit-identical to it.
//
// 576 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue (four per TMEM lane quadrant, one 64-column
// group of every 256-column chunk each; the accumulator stage is released as soon as it is in registers).
#include <cstdlib>

#include "lstm_seq.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include "lstm_epi.cuh"

namespace dsb {
namespace {

using namespace tc;

constexpr int SEQ_EPI_WARPS = 16;     // epilogue warps (8 or 16)
constexpr uint32_t SEQ_HTILE_BYTES = 128 * 64 * 2;   // h_t of one chunk (64 hidden units) staged for the TMA store
constexpr int SEQ_BM = 128;
constexpr int SEQ_BN = 256;
constexpr int SEQ_BK = 64;
constexpr int SEQ_STAGES = 4;
constexpr uint32_t SEQ_A_BYTES = SEQ_BM * SEQ_BK * 2;       // 16 KB
constexpr uint32_t SEQ_B_BYTES = SEQ_BN * SEQ_BK * 2;       // 32 KB
constexpr uint32_t SEQ_STAGE_BYTES = SEQ_A_BYTES + SEQ_B_BYTES;
constexpr int SEQ_CHUNKS = GATES / SEQ_BN;                  // 4
constexpr size_t seq_smem(int ew) { return 1024 + (size_t)SEQ_STAGES * SEQ_STAGE_BYTES + 256 + 2 * GATES * sizeof(float) + SEQ_HTILE_BYTES + 0 * (size_t)ew; }


// Walks the (pair, layer, t, slot, chunk) sequence shared by the three roles.
template <class Body>
__device__ __forceinline__ void seq_walk(const LstmSeqParams& p, Body&& body) {
  const int m_tiles = (p.n + SEQ_BM - 1) / SEQ_BM;
  const int items = 2 * m_tiles;
  const int pairs = (items + 1) / 2;
  for (int pair = blockIdx.x; pair < pairs; pair += gridDim.x)
    for (int l = 0; l < LSTM_SEQ_LAYERS; ++l)
      for (int t = 0; t < p.KL; ++t)
        for (int s = 0; s < 2; ++s) {
          const int item = 2 * pair + s;
          if (item >= items) continue;
          const int d = item / m_tiles, mt = item - d * m_tiles;
          body(pair, l, t, s, d, mt);
        }
}

template <bool F16, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
lstm_seq_kernel(const __grid_constant__ LstmSeqMaps maps, const LstmSeqParams p, const int ablate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring_base = smem_base;
  const uint32_t htile = ring_base + SEQ_STAGES * SEQ_STAGE_BYTES;      // [128 rows x 64 units] 16-bit, SWIZZLE_128B: TMA-store source
  const uint32_t bar_base = htile + SEQ_HTILE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SEQ_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SEQ_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SEQ_STAGES + 2 + a); };
  auto hready_bar = [&](int s) { return bar_base + 8u * (2 * SEQ_STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SEQ_STAGES + 6);
  const uint32_t bias_off = (bar_base + 256u) - smem_u32(smem_raw);
  float* sbias = reinterpret_cast<float*>(smem_raw + bias_off);                  // [2 slots][1024]

  if (threadIdx.x == 0) {
    for (int s = 0; s < SEQ_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EW); mbar_init(hready_bar(a), 1); }
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

  const int cap = (int)p.cap;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t steps_done[2] = {0u, 0u};          // completed-step counters per item slot (hready phases)
      int cur_pair = -1;
      seq_walk(p, [&](int pair, int l, int t, int s, int d, int mt) {
        if (pair != cur_pair) { cur_pair = pair; }
        const int row0 = mt * SEQ_BM;
        const int pos = (d == 0) ? t : p.KL - 1 - t;         // bw consumes the k-mer reversed
        // h_{t-1} of this item (and everything it wrote earlier) must be globally visible before any load of this step
        if (steps_done[s] > 0) mbar_wait(hready_bar(s), (steps_done[s] - 1u) & 1u);
        ++steps_done[s];
        const CUtensorMap* wmap = &maps.w[d][l];
        for (int ch = 0; ch < SEQ_CHUNKS; ++ch) {
          auto load = [&](const CUtensorMap* amap, int col, int row, int wkb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), SEQ_STAGE_BYTES);
            const uint32_t a_dst = ring_base + stage * SEQ_STAGE_BYTES;
            tma_load_2d(a_dst, amap, full_bar(stage), col, row);
            tma_load_2d(a_dst + SEQ_A_BYTES, wmap, full_bar(stage), wkb * SEQ_BK, ch * SEQ_BN);
            if (++stage == SEQ_STAGES) { stage = 0; phase ^= 1u; }
          };
          if (l == 0) {
            if (t > 0)
              for (int j = 0; j < 4; ++j) load(&maps.h[d][0], j * SEQ_BK, t * cap + row0, j);
            load(&maps.f, 0, pos * cap + row0, 4);
          } else {
            for (int j = 0; j < 4; ++j) load(&maps.h[d][l - 1], j * SEQ_BK, (t + 1) * cap + row0, j);
            if (t > 0)
              for (int j = 0; j < 4; ++j) load(&maps.h[d][l], j * SEQ_BK, t * cap + row0, 4 + j);
          }
        }
      });
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(SEQ_BN >> 3) << 17) | ((uint32_t)(SEQ_BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      seq_walk(p, [&](int, int l, int t, int, int, int) {
        const int nkb = (l == 0) ? (t > 0 ? 5 : 1) : (t > 0 ? 8 : 4);
        for (int ch = 0; ch < SEQ_CHUNKS; ++ch) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SEQ_BN);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = ring_base + stage * SEQ_STAGE_BYTES;
            const uint64_t adesc = make_kmajor_desc<SEQ_BK>(a_addr);
            const uint64_t bdesc = make_kmajor_desc<SEQ_BK>(a_addr + SEQ_A_BYTES);
#pragma unroll
            for (int k = 0; k < SEQ_BK / 16; ++k)
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == SEQ_STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull_bar(acc));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      });
    }
  } else {
    // ================= epilogue warps (EW = 8 or 16; EW/4 per TMEM lane quadrant) =================
    const int q = warp & 3;                   // TMEM lane quadrant this warp may access
    const int cs = (warp - 2) >> 2;           // column set: 64-column groups cs, cs + EW/4, ... of every chunk
    const int et = threadIdx.x - 64;
    constexpr int ET = EW * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    int bias_pair = -1, bias_layer = -1;
    static_assert(EW == 16, "one 64-column group per warp and chunk");
    bool store_pending = false;               // (elected thread) a TMA store may still be reading the h tile
    seq_walk(p, [&](int pair, int l, int t, int s, int d, int mt) {
      if (pair != bias_pair || l != bias_layer) {
        // new (pair, layer): both slots' biases into shared memory
        bias_pair = pair;
        bias_layer = l;
        named_bar_sync(1, ET);
        const int m_tiles = (p.n + SEQ_BM - 1) / SEQ_BM;
        for (int ss = 0; ss < 2; ++ss) {
          const int item = 2 * pair + ss;
          if (item >= 2 * m_tiles) continue;
          const float* b = p.bias[item / m_tiles][l];
          for (int i = et; i < GATES; i += ET) sbias[ss * GATES + i] = b[i];
        }
        named_bar_sync(1, ET);
      }
      const float* sb = sbias + s * GATES;
      const bool c_in = t > 0 && !(ablate & 1), c_out = t + 1 < p.KL && !(ablate & 1);
      for (int ch = 0; ch < SEQ_CHUNKS; ++ch) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * SEQ_BN);
        constexpr int GROUPS = 4 / (EW / 4);          // 64-column groups per warp and chunk
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
          const int c = (cs + g * (EW / 4)) * 64;
          const int col = ch * SEQ_BN + c;
          float4* cblk = reinterpret_cast<float4*>(p.c[d]) + ((((size_t)mt * (GATES / 64) + (size_t)(col >> 6)) * 4 + q) * 4) * 32 + lane;
          uint32_t pk[8];
          lstm_epi_group<F16>(t_row + (uint32_t)c, cblk, c_in, c_out, sb + col, pk, ablate, [&]() {
            // the accumulator stage is in registers now: hand it back to the MMA warp before doing the math
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          });
          // h_t of this chunk: [128 x 64] tile in shared memory (128-byte rows, 16-byte pieces XOR-swizzled by row & 7 =
          // the SWIZZLE_128B pattern of the H tensor map) -> one TMA store of whole 128-byte lines
          if (et == 0 && store_pending && !(ablate & 16)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          named_bar_sync(2, ET);
          {
            const uint32_t rr = (uint32_t)(q * 32 + lane);
            const uint32_t rowp = htile + rr * 128u;
            sts128(rowp + (((uint32_t)(2 * cs) ^ (rr & 7u)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
            sts128(rowp + (((uint32_t)(2 * cs + 1) ^ (rr & 7u)) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
          }
          fence_proxy_async();
          named_bar_sync(2, ET);
          if (et == 0) {
            if (!(ablate & 2)) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&maps.h[d][l])), "r"(htile), "r"(ch * 64), "r"((t + 1) * cap + mt * SEQ_BM)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
            if (ch == SEQ_CHUNKS - 1) {
              // publish h_t: stores complete (async proxy) -> the TMA loads of the next step may read them
              if (!(ablate & 8)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
              mbar_arrive(hready_bar(s));
              store_pending = false;
            }
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    });
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

void launch_lstm_seq(const LstmSeqLaunch& L, int n, int sm_count, cudaStream_t s) {
  if (n <= 0) return;
  const int m_tiles = (n + SEQ_BM - 1) / SEQ_BM;
  const int pairs = m_tiles;                       // 2 * m_tiles items, two per pair
  const int grid = pairs < sm_count ? pairs : sm_count;
  LstmSeqParams p = L.p;
  p.n = n;
  static const int ablate = getenv("DSB_ABLATE_S") ? atoi(getenv("DSB_ABLATE_S")) : 0;   // timing experiments only
  static bool configured[64][2] = {{false}};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  constexpr int EW = SEQ_EPI_WARPS;
  const size_t smem = seq_smem(EW);
  auto kern = L.fp16 ? lstm_seq_kernel<true, EW> : lstm_seq_kernel<false, EW>;
  if (!configured[dev & 63][L.fp16 ? 1 : 0]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev & 63][L.fp16 ? 1 : 0] = true;
  }
  launch_pdl(kern, grid, 64 + 32 * EW, smem, s, L.maps, p, ablate);
}

}  // namespace dsb
