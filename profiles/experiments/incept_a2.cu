// incept_a2.cu — inception kernel A (incept_tc.cu) on CTA pairs: tcgen05.mma cta_group::2, M = 256 (each CTA its own
// 128-row tile), every 1x1 weight matrix split over the two CTAs (p1: 96 + 96 of 192 rows, b1: 24 + 24 of 48).  The
// weights take 60 KB instead of 120 KB per SM, the freed space goes into a deeper x ring (8 instead of 5 slots), and the
// tensor cores read half of the weight operand from shared memory per output row.  Same arithmetic, bit-identical.
// It is also the first half of the planned single-kernel inception block, which needs exactly
// this weight split.  Protocol as lstm_seq2.cu; x tiles complete on the CTA's own barrier (its pool warps wait on it)
// and the peer relays "slot full" to the leader with one remote arrive.
#include <cstdlib>

#include "incept_tc.cuh"
#include "tc_common.cuh"

namespace dsb {
namespace {

using namespace tc;

template <bool F16> __device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
  if constexpr (F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}
__device__ __forceinline__ uint32_t idesc2(bool f16, int n) {        // M = 256 (cta_group::2)
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
// relaxed: carries no data and may be issued by a thread with bulk loads in flight (profiles/experiments/pair_ring.cu)
__device__ __forceinline__ void arrive_cluster_relaxed(uint32_t a) { asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
// release: publishes this thread's (fenced) shared-memory writes to the MMA the leader issues next
__device__ __forceinline__ void arrive_cluster_release(uint32_t a) {
#ifdef A2_RELAXED
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
#else
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
#endif
}
__device__ __forceinline__ void tma_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

constexpr int A_THREADS = 448;       // warp 0 TMA, warp 1 MMA (leader only), warps 2-5 pool, warps 6-13 epilogue
constexpr int A2_SLOTS = 8;
constexpr uint32_t A_SLOT_BYTES = 128 * 64 * 2;        // one [128 x 64] SWIZZLE_128B tile
constexpr uint32_t A2_WP1_BLK = 96 * 64 * 2;           // this CTA's half of a p1 k-block
constexpr uint32_t A2_WB1_BLK = 24 * 64 * 2;           // this CTA's half of a b1 k-block
constexpr int A_NKB = 4;
constexpr uint32_t A_ACC_STRIDE = 256;

struct A2Smem {
  static constexpr uint32_t wp1 = 0;
  static constexpr uint32_t wb1 = wp1 + A_NKB * A2_WP1_BLK;
  static constexpr uint32_t ring = (wb1 + A_NKB * A2_WB1_BLK + 1023u) & ~1023u;
  static constexpr uint32_t stage = ring + A2_SLOTS * A_SLOT_BYTES;          // 8 epilogue warps x 2 KB
  static constexpr uint32_t bars = stage + 8 * EPI_STAGE_BYTES;
  // barriers: full_x[S] p1_done[S] full_xp[S] empty[S] wbar tfull[2] tempty[2] pfull_x[S]
  static constexpr uint32_t nbars = 5 * A2_SLOTS + 5;
  static constexpr uint32_t tmem_slot = (bars + 8 * nbars + 15u) & ~15u;
  static constexpr uint32_t total = tmem_slot + 16;
};

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(A_THREADS, 1)
incept_a2_kernel(const __grid_constant__ InceptAMaps maps, const __grid_constant__ InceptAParams p, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto full_x = [&](int s) { return base + A2Smem::bars + 8u * s; };
  auto p1_done = [&](int s) { return base + A2Smem::bars + 8u * (A2_SLOTS + s); };
  auto full_xp = [&](int s) { return base + A2Smem::bars + 8u * (2 * A2_SLOTS + s); };
  auto empty = [&](int s) { return base + A2Smem::bars + 8u * (3 * A2_SLOTS + s); };
  const uint32_t wbar = base + A2Smem::bars + 8u * (4 * A2_SLOTS);
  auto tfull = [&](int a) { return base + A2Smem::bars + 8u * (4 * A2_SLOTS + 1 + a); };
  auto tempty = [&](int a) { return base + A2Smem::bars + 8u * (4 * A2_SLOTS + 3 + a); };
  auto pfull_x = [&](int s) { return base + A2Smem::bars + 8u * (4 * A2_SLOTS + 5 + s); };     // leader: the peer's x slot is full
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int cluster_id = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const int tile_pairs = (num_tiles + 1) / 2;
  const uint32_t tmem_slot = base + A2Smem::tmem_slot;

  if (threadIdx.x == 0) {
    for (int s = 0; s < A2_SLOTS; ++s) { mbar_init(full_x(s), 1); mbar_init(p1_done(s), 1); mbar_init(full_xp(s), 8); mbar_init(empty(s), 1); mbar_init(pfull_x(s), 1); }
    mbar_init(wbar, 2);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // both CTAs' barriers exist before any remote arrive / multicast
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait_prior_grid();

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0 && cluster_id < tile_pairs) {
      // this CTA's half of the weights (p1 rows rank*96.., b1 rows rank*24..); both CTAs' bytes are counted on the leader's wbar
      const uint32_t lwbar = map_to_cta(wbar, 0);
      if (leader) mbar_expect_tx(wbar, 2 * A_NKB * (A2_WP1_BLK + A2_WB1_BLK)); else arrive_cluster_relaxed(lwbar);
      for (int j = 0; j < A_NKB; ++j) {
        tma_2sm(base + A2Smem::wp1 + j * A2_WP1_BLK, &maps.wp1, lwbar, j * 64, (int)rank * 96);
        tma_2sm(base + A2Smem::wb1 + j * A2_WB1_BLK, &maps.wb1, lwbar, j * 64, (int)rank * 24);
      }
      int slot = 0;
      uint32_t phase = 0;
      for (int tp = cluster_id; tp < tile_pairs; tp += n_clusters) {
        const int tile = 2 * tp + (int)rank;
        const int row0 = tile * INCEPT_A_TILE_ROWS - 1;
        for (int j = 0; j < A_NKB; ++j) {
          mbar_wait(empty(slot), phase ^ 1u);
          mbar_expect_tx(full_x(slot), A_SLOT_BYTES);          // x tiles complete on the CTA's own barrier (its pool warps wait on it)
          tma_load_2d(base + A2Smem::ring + slot * A_SLOT_BYTES, &maps.x, full_x(slot), j * 64, row0);
          if (++slot == A2_SLOTS) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && leader && cluster_id < tile_pairs) {
      const uint32_t id_p1 = idesc2(F16, 192), id_b1 = idesc2(F16, 48);
      mbar_wait(wbar, 0);
      int slot = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tp = cluster_id; tp < tile_pairs; tp += n_clusters) {
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_b1 = tmem_base + acc * A_ACC_STRIDE, d_p1 = d_b1 + 48;
        // software-pipelined issue order p1(0) p1(1) b1(0) p1(2) b1(1) p1(3) b1(2) b1(3): a slot is released one
        // step after its p1 MMAs, so the TMA ring keeps most of its slots in flight.
        int sp = slot, sb = slot;
        uint32_t php = phase, phb = phase;
        for (int step = 0; step <= A_NKB; ++step) {
          if (step < A_NKB) {                              // p1: x tile straight from TMA
            mbar_wait(full_x(sp), php);
            mbar_wait(pfull_x(sp), php);
            tc_fence_after();
            const uint64_t ad = make_kmajor_desc<64>(base + A2Smem::ring + sp * A_SLOT_BYTES);
            const uint64_t bd = make_kmajor_desc<64>(base + A2Smem::wp1 + step * A2_WP1_BLK);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_2sm(d_p1, ad + 2 * k, bd + 2 * k, id_p1, (step | k) != 0 ? 1u : 0u);
            commit_2sm(p1_done(sp), 3);
            if (++sp == A2_SLOTS) { sp = 0; php ^= 1u; }
          }
          if (step >= 1) {                                 // b1: the same slot after the pool warps rewrote it
            const int j = step - 1;
            mbar_wait(full_xp(sb), phb);
            tc_fence_after();
            const uint64_t ad = make_kmajor_desc<64>(base + A2Smem::ring + sb * A_SLOT_BYTES);
            const uint64_t bd = make_kmajor_desc<64>(base + A2Smem::wb1 + j * A2_WB1_BLK);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_2sm(d_b1, ad + 2 * k, bd + 2 * k, id_b1, (j | k) != 0 ? 1u : 0u);
            commit_2sm(empty(sb), 3);
            if (++sb == A2_SLOTS) { sb = 0; phb ^= 1u; }
          }
        }
        int s = sp;
        uint32_t ph = php;
        slot = s;
        phase = ph;
        commit_2sm(tfull(acc), 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 6) {
    // ---------------- pool warps: maxpool 1x3 over rows, in place ----------------
    // thread -> (16-byte channel chunk c, group of 8 rows): ten row loads give eight pooled rows.
    const int pt = threadIdx.x - 64;
    const uint32_t c = (uint32_t)(pt & 7);
    const int r0 = (pt >> 3) * 8;
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t lfull_xp0 = map_to_cta(full_xp(0), 0), lpfull0 = map_to_cta(pfull_x(0), 0);
    for (int tp = cluster_id; tp < tile_pairs; tp += n_clusters) {
      for (int j = 0; j < A_NKB; ++j) {
        const uint32_t tb = base + A2Smem::ring + slot * A_SLOT_BYTES;
        mbar_wait(full_x(slot), phase);
        if (!leader && pt == 0) arrive_cluster_release(lpfull0 + 8u * slot);      // relay: the peer's x slot is full
        if (p.ablate & 1) { mbar_wait(p1_done(slot), phase); __syncwarp(); if (lane == 0) arrive_cluster_release(lfull_xp0 + 8u * slot); if (++slot == A2_SLOTS) { slot = 0; phase ^= 1u; } continue; }
        uint4 in[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const int rr = r0 - 1 + k;
          // logical 16-byte chunk c of row rr sits at physical chunk c ^ (rr & 7)  (SWIZZLE_128B)
          if (rr >= 0 && rr < 128) in[k] = lds128(tb + (uint32_t)rr * 128u + ((c ^ (uint32_t)(rr & 7)) << 4));
          else in[k] = make_uint4(0u, 0u, 0u, 0u);          // activations are >= 0: zero is neutral for max
        }
        uint4 out[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          out[k].x = max16x2<F16>(max16x2<F16>(in[k].x, in[k + 1].x), in[k + 2].x);
          out[k].y = max16x2<F16>(max16x2<F16>(in[k].y, in[k + 1].y), in[k + 2].y);
          out[k].z = max16x2<F16>(max16x2<F16>(in[k].z, in[k + 1].z), in[k + 2].z);
          out[k].w = max16x2<F16>(max16x2<F16>(in[k].w, in[k + 1].w), in[k + 2].w);
        }
        mbar_wait(p1_done(slot), phase);                 // the p1 MMAs have finished reading this slot
        named_bar_sync(1, 128);                          // every pool thread has finished reading its neighbours
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = r0 + k;
          sts128(tb + (uint32_t)rr * 128u + ((c ^ (uint32_t)(rr & 7)) << 4), out[k]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) arrive_cluster_release(lfull_xp0 + 8u * slot);        // the leader's MMA issuer waits for both CTAs' pool warps
        if (++slot == A2_SLOTS) { slot = 0; phase ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const uint32_t stage = base + A2Smem::stage + (uint32_t)(warp - 6) * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t ltempty0 = map_to_cta(tempty(0), 0);
    for (int tp = cluster_id; tp < tile_pairs; tp += n_clusters) {
      const int tile = 2 * tp + (int)rank;
      const int i = q * 32 + lane;                       // tile row
      const int row = tile * INCEPT_A_TILE_ROWS - 1 + i;
      const int wrow0 = tile * INCEPT_A_TILE_ROWS - 1 + q * 32;   // global row of this warp's lane 0
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * A_ACC_STRIDE;
      const bool row_ok = i >= 1 && i <= INCEPT_A_TILE_ROWS && row < p.M && !(p.ablate & 2);
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const bool halo = (row % p.Lp) >= p.L;
      // 32-column groups: g0..g2 -> Y[:, 0:96) (b1 | b2), g3/g4/g5 -> planes P3/P4/P5, g6 / g7 -> stem plane (g7 = stem
      // channels 16..47, overlapping g6 by 16 so that every group is 32 wide).
      // g0, g1 hold b1, whose first / last tile row lacks a pool neighbour: row-masked staged stores.  Everything from g2 on is a
      // 1x1 conv of x and is valid on all 128 tile rows (the rows shared with the neighbouring tile get identical values),
      // so those groups leave as one TMA store per warp: no shared-memory read-back by the threads.
      for (int g = half; g < 8; g += 2) {
        const int c = g < 7 ? g * 32 : 208;
        uint32_t rr[2][16];
        __syncwarp();
        tc_ld16(t_row + (uint32_t)c, rr[0]);
        tc_ld16(t_row + (uint32_t)(c + 16), rr[1]);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int col = c + 16 * h2;
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(rr[h2][k]) + p.biasv[col + k];
          if (g < 6) {                                    // the residual stem (g6, g7) keeps its sign
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
          }
          if (halo) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) pk[8 * h2 + k] = pack16<F16>(v[2 * k], v[2 * k + 1]);
        }
        // the staging tile may still be the source of this warp's previous TMA store
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (g < 2 || wrow0 < 0) {                         // (TMA stores reject negative coordinates: first warp of tile 0)
          uint8_t* dst;
          size_t pitch;
          if (g < 3) { pitch = (size_t)p.ldY * 2; dst = reinterpret_cast<uint8_t*>(p.Y) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch + g * 64; }
          else if (g < 6) { pitch = 64; dst = reinterpret_cast<uint8_t*>(g == 3 ? p.P3 : g == 4 ? p.P4 : p.P5) + (ptrdiff_t)wrow0 * 64; }
          else { pitch = 96; dst = reinterpret_cast<uint8_t*>(p.PS) + (ptrdiff_t)wrow0 * 96 + (g == 7 ? 32 : 0); }
          staged_store<4>(stage, pk, lane, ok_mask, dst, pitch);
        } else {
          const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
            sts128(stage + (uint32_t)lane * 64u + (((uint32_t)pc ^ sw) << 4), make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.ablate & 2)) {
            const CUtensorMap* mp = g == 2 ? &maps.ys : g == 3 ? &maps.p3s : g == 4 ? &maps.p4s : g == 5 ? &maps.p5s : &maps.pss;
            const int col = g == 2 ? 64 : g == 7 ? 16 : 0;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(mp)), "r"(stage), "r"(col), "r"(wrow0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_cluster_relaxed(ltempty0 + 8u * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory stays valid until every store has read it
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // no CTA leaves while its partner may still address its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <class K> void set_smem_once2(K kern, bool* flags) {
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  if (!flags[dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    flags[dev & 63] = true;
  }
}

}  // namespace

// L.maps.wp1 / wb1 must be encoded with 96- / 24-row boxes (forward.cu does that when the pair kernel is selected).
void launch_incept_a2(const InceptALaunch& L, int M, int sm_count, cudaStream_t s) {
  if (M <= 0) return;
  const int num_tiles = (M + INCEPT_A_TILE_ROWS - 1) / INCEPT_A_TILE_ROWS;
  const int tile_pairs = (num_tiles + 1) / 2;
  int clusters = sm_count / 2;
  if (clusters > tile_pairs) clusters = tile_pairs;
  InceptAParams p = L.p;
  p.M = M;
  if (const char* ev = getenv("DSB_ABLATE_A")) p.ablate = atoi(ev);
  const size_t smem = A2Smem::total + 1024;
  static bool cfg[2][64] = {{false}};
  if (L.fp16) {
    set_smem_once2(incept_a2_kernel<true>, cfg[1]);
    launch_pdl(incept_a2_kernel<true>, 2 * clusters, A_THREADS, smem, s, L.maps2, p, num_tiles);
  } else {
    set_smem_once2(incept_a2_kernel<false>, cfg[0]);
    launch_pdl(incept_a2_kernel<false>, 2 * clusters, A_THREADS, smem, s, L.maps2, p, num_tiles);
  }
  DSB_CUDA(cudaGetLastError());
}

}  // namespace dsb
