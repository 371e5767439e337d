// incept_f.cu — one inception block (reference layers.py:87-139) as ONE kernel: incept_a_kernel + incept_b_kernel fused so
// that the block-internal planes (b3a, b4a, b5a, residual stem: 576 of the 1536 HBM bytes per row) never leave the SM.
// Opt-in (DSB_AB_FUSED=1): functionally complete and bit-identical to A + B, but not yet faster (12.9 ms against 8.1 ms per
// 37,888-sample chunk): the five-stage chain of a tile is not overlapped across tiles yet (see the end of the MMA issuer).
//
// A CTA pair works on two neighbouring 128-row tiles (124 output rows each: b4 needs +-2 plane rows).  Only the p1 GEMM
// (x from TMA, 192 output channels) runs as tcgen05.mma cta_group::2 with its weights split over the pair; every GEMM
// whose A operand is produced by threads (pooled x, planes, U) is a cta_group::1 MMA the CTA issues for itself with its
// own full copy of those (small) weights, so no thread-produced operand needs a cross-CTA hand-off
// (profiles/experiments/mix_cta_group.cu, profiles/r1_ablation.md).
//
// Per tile and CTA:   TMA x k-blocks -> p1 (pair) -> 4 pool warps max-pool the slot in place -> b1 (local)
//                     converters: TMEM -> b2 to Y, planes P3/P4/P5 to 16-bit smem tiles (two rows down, halo rows zero)
//                     b5b / b3 / b4 taps (local, row-shifted descriptors) -> converters: relu(b5b) -> U tile -> b5c (local)
//                     epilogue: b1, b3, b4, relu(b5c + stem) -> Y
// TMEM, 256 columns per tile (two tiles in flight): [b1 0..47 | - | b2 64..111 | P3 112..143 | P4 144..175 | P5 176..207 |
// stem 208..255]; after the converters drained them: b5b -> 48..111 (b5c re-uses 48..95), b3 -> 112..159, b4 -> 160..207.
// 704 threads: warp 0 TMA, warp 1 MMA, warps 2-5 pool, 6-13 converters, 14-21 epilogue.
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
__device__ __forceinline__ uint32_t idesc(bool f16, int m, int n) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
__device__ __forceinline__ void arrive_cluster_relaxed(uint32_t a) { asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void tma_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

constexpr int F_THREADS = 704;
constexpr int F_TILE_OUT = 124;                        // output rows per 128-row tile (tile rows 2..125)
constexpr int XS = 3;                                  // x ring slots ([128 x 64] SWIZZLE_128B each)
constexpr uint32_t X_SLOT = 128 * 64 * 2;
constexpr int NKB = 4;
constexpr uint32_t WP1H_BLK = 96 * 64 * 2;             // this CTA's half of a p1 k-block
constexpr uint32_t WB1_BLK = 48 * 64 * 2;
constexpr uint32_t W48_BLK = 48 * 32 * 2, W64_BLK = 64 * 32 * 2;
constexpr uint32_t PLANE_BYTES = 136 * 32 * 2;         // [136 x 32] SWIZZLE_64B
constexpr uint32_t U_KB = 128 * 32 * 2;

struct FS {
  static constexpr uint32_t wp1 = 0;
  static constexpr uint32_t wb1 = wp1 + NKB * WP1H_BLK;             // 49152
  static constexpr uint32_t w3 = wb1 + NKB * WB1_BLK;               // 73728
  static constexpr uint32_t w4 = w3 + 3 * W48_BLK;
  static constexpr uint32_t w5b = w4 + 5 * W48_BLK;
  static constexpr uint32_t w5c = w5b + 3 * W64_BLK;
  static constexpr uint32_t ring = (w5c + 2 * W48_BLK + 1023u) & ~1023u;
  static constexpr uint32_t planes = ring + XS * X_SLOT;            // P5, P3, P4
  static constexpr uint32_t u = planes + 3 * PLANE_BYTES;           // 2 k-blocks; doubles as the converters' b2 staging (8 x 2 KB)
  static constexpr uint32_t stage = u + 2 * U_KB;                   // 8 epilogue warps x 2 KB
  static constexpr uint32_t bars = stage + 8 * EPI_STAGE_BYTES;
  static constexpr uint32_t nbars = 40;
  static constexpr uint32_t tmem_slot = (bars + 8 * nbars + 15u) & ~15u;
  static constexpr uint32_t total = tmem_slot + 16;
};

struct alignas(64) FMaps { CUtensorMap x, wp1h, wb1, w3, w4, w5b, w5c; };
struct FParams {
  int M, Lp, L;
  void* Y; int ldY;
  float ba[240];        // b1 48 | b2 48 | b3a 32 | b4a 32 | b5a 32 | stem 48   (accumulator order of kernel A)
  float bb[208];        // b3 48 | b4 48 | b5c 48 | b5b 64                       (kernel B)
};

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1)
incept_f_kernel(const __grid_constant__ FMaps maps, const __grid_constant__ FParams p, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int cluster_id = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const int tile_pairs = (num_tiles + 1) / 2;
  const int my_tiles = cluster_id < tile_pairs ? (tile_pairs - cluster_id + n_clusters - 1) / n_clusters : 0;
  auto bar = [&](int i) { return base + FS::bars + 8u * i; };
  // barrier indices
  enum { FULL_X = 0 /*3, leader*/, P1_DONE = 3 /*3*/, FULL_XP = 6 /*3*/, EMPTY = 9 /*3*/, WPAIR = 12 /*leader*/, WLOCAL = 13,
         P1_ALL = 14 /*2*/, PL_FULL = 16 /*2*/, PL_EMPTY = 18, B5B_FULL = 19 /*2*/, U_FULL = 21, U_EMPTY = 22,
         TFULL = 23 /*2*/, TEMPTY = 25 /*2, local*/, TEMPTY_PAIR = 27 /*2, leader*/ };
  const uint32_t tmem_slot = base + FS::tmem_slot;

  if (threadIdx.x == 0) {
    for (int s = 0; s < XS; ++s) { mbar_init(bar(FULL_X + s), 2); mbar_init(bar(P1_DONE + s), 1); mbar_init(bar(FULL_XP + s), 4); mbar_init(bar(EMPTY + s), 1); }
    mbar_init(bar(WPAIR), 2); mbar_init(bar(WLOCAL), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(P1_ALL + a), 1); mbar_init(bar(PL_FULL + a), 8); mbar_init(bar(B5B_FULL + a), 1);
      mbar_init(bar(TFULL + a), 1); mbar_init(bar(TEMPTY + a), 8); mbar_init(bar(TEMPTY_PAIR + a), 16);
    }
    mbar_init(bar(PL_EMPTY), 1); mbar_init(bar(U_FULL), 8); mbar_init(bar(U_EMPTY), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait_prior_grid();

  auto tile_of = [&](int k) { return 2 * (cluster_id + k * n_clusters) + (int)rank; };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0 && my_tiles > 0) {
      const uint32_t lwp = map_to_cta(bar(WPAIR), 0);
      if (leader) mbar_expect_tx(bar(WPAIR), 2 * NKB * WP1H_BLK); else arrive_cluster_relaxed(lwp);
      for (int j = 0; j < NKB; ++j) tma_2sm(base + FS::wp1 + j * WP1H_BLK, &maps.wp1h, lwp, j * 64, (int)rank * 96);
      mbar_expect_tx(bar(WLOCAL), NKB * WB1_BLK + 8 * W48_BLK + 3 * W64_BLK + 2 * W48_BLK);
      for (int j = 0; j < NKB; ++j) tma_load_2d(base + FS::wb1 + j * WB1_BLK, &maps.wb1, bar(WLOCAL), j * 64, 0);
      for (int t = 0; t < 3; ++t) tma_load_2d(base + FS::w3 + t * W48_BLK, &maps.w3, bar(WLOCAL), t * 32, 0);
      for (int t = 0; t < 5; ++t) tma_load_2d(base + FS::w4 + t * W48_BLK, &maps.w4, bar(WLOCAL), t * 32, 0);
      for (int t = 0; t < 3; ++t) tma_load_2d(base + FS::w5b + t * W64_BLK, &maps.w5b, bar(WLOCAL), t * 32, 0);
      for (int t = 0; t < 2; ++t) tma_load_2d(base + FS::w5c + t * W48_BLK, &maps.w5c, bar(WLOCAL), t * 32, 0);
      int slot = 0;
      uint32_t ph = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int row0 = tile_of(k) * F_TILE_OUT - 2;
        for (int j = 0; j < NKB; ++j) {
          mbar_wait(bar(EMPTY + slot), ph ^ 1u);
          const uint32_t lfull = map_to_cta(bar(FULL_X + slot), 0);
          if (leader) mbar_expect_tx(bar(FULL_X + slot), 2 * X_SLOT); else arrive_cluster_relaxed(lfull);
          tma_2sm(base + FS::ring + slot * X_SLOT, &maps.x, lfull, j * 64, row0);
          if (++slot == XS) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: p1 for the pair (leader only), everything else for this CTA ----------------
    if (lane == 0 && my_tiles > 0) {
      const uint32_t id_p1 = idesc(F16, 256, 192), id_b1 = idesc(F16, 128, 48), id48 = idesc(F16, 128, 48), id64 = idesc(F16, 128, 64);
      if (leader) mbar_wait(bar(WPAIR), 0);
      mbar_wait(bar(WLOCAL), 0);
      int sp = 0, sb = 0;                       // slot cursors of p1 / b1
      uint32_t php = 0, phb = 0;
      auto a_stage = [&](int k) {               // x -> p1 (pair) -> pooled x -> b1 (local)
        const int a = k & 1;
        const uint32_t aph = (uint32_t)(k >> 1) & 1u;
        const uint32_t d0 = tmem_base + (uint32_t)a * 256u;
        if (leader) mbar_wait(bar(TEMPTY_PAIR + a), aph ^ 1u);
        mbar_wait(bar(TEMPTY + a), aph ^ 1u);
        tc_fence_after();
        auto p1 = [&](int j) {
          if (!leader) return;
          mbar_wait(bar(FULL_X + sp), php);
          tc_fence_after();
          const uint64_t ad = make_kmajor_desc<64>(base + FS::ring + sp * X_SLOT);
          const uint64_t bd = make_kmajor_desc<64>(base + FS::wp1 + j * WP1H_BLK);
#pragma unroll
          for (int q = 0; q < 4; ++q) mma_2sm(d0 + 64, ad + 2 * q, bd + 2 * q, id_p1, (j | q) != 0 ? 1u : 0u);
          commit_2sm(bar(P1_DONE + sp), 3);
          if (j == NKB - 1) commit_2sm(bar(P1_ALL + a), 3);
          if (++sp == XS) { sp = 0; php ^= 1u; }
        };
        auto b1 = [&](int j) {
          mbar_wait(bar(FULL_XP + sb), phb);
          tc_fence_after();
          const uint64_t ad = make_kmajor_desc<64>(base + FS::ring + sb * X_SLOT);
          const uint64_t bd = make_kmajor_desc<64>(base + FS::wb1 + j * WB1_BLK);
#pragma unroll
          for (int q = 0; q < 4; ++q) tc_mma_bf16(d0, ad + 2 * q, bd + 2 * q, id_b1, (j | q) != 0 ? 1u : 0u);
          tc_commit(bar(EMPTY + sb));
          if (++sb == XS) { sb = 0; phb ^= 1u; }
        };
        // three x slots for four k-blocks: p1(3) re-uses the slot b1(0) frees
        p1(0); p1(1); b1(0); p1(2); b1(1); p1(3); b1(2); b1(3);
      };
      auto b_stage = [&](int k) {               // planes -> taps -> U -> b5c
        const int a = k & 1;
        const uint32_t aph = (uint32_t)(k >> 1) & 1u;
        const uint32_t d0 = tmem_base + (uint32_t)a * 256u;
        auto taps = [&](int ntaps, int first_row, uint32_t plane, uint32_t w_off, uint32_t w_blk, uint32_t d, uint32_t id) {
          for (int t = 0; t < ntaps; ++t) {
            const uint64_t ad = make_kmajor_desc<32>(base + FS::planes + plane * PLANE_BYTES + (uint32_t)(first_row + t) * 64u);
            const uint64_t bd = make_kmajor_desc<32>(base + w_off + t * w_blk);
#pragma unroll
            for (int q = 0; q < 2; ++q) tc_mma_bf16(d, ad + 2 * q, bd + 2 * q, id, (t | q) != 0 ? 1u : 0u);
          }
        };
        mbar_wait(bar(PL_FULL + a), aph);
        tc_fence_after();
        // planes sit two rows down in their tiles: output row i, tap offset d reads tile row i + 2 + d
        taps(3, 1, 0, FS::w5b, W64_BLK, d0 + 48, id64);          // b5b: 1x3 on P5 -> columns 48..111
        tc_commit(bar(B5B_FULL + a));
        taps(3, 1, 1, FS::w3, W48_BLK, d0 + 112, id48);          // b3: 1x3 on P3 -> 112..159
        taps(5, 0, 2, FS::w4, W48_BLK, d0 + 160, id48);          // b4: 1x5 on P4 -> 160..207
        tc_commit(bar(PL_EMPTY));
        mbar_wait(bar(U_FULL), (uint32_t)k & 1u);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ad = make_kmajor_desc<32>(base + FS::u + kb * U_KB);
          const uint64_t bd = make_kmajor_desc<32>(base + FS::w5c + kb * W48_BLK);
#pragma unroll
          for (int q = 0; q < 2; ++q) tc_mma_bf16(d0 + 48, ad + 2 * q, bd + 2 * q, id48, (kb | q) != 0 ? 1u : 0u);
        }
        tc_commit(bar(U_EMPTY));
        tc_commit(bar(TFULL + a));
      };
      // Tile after tile.  Issuing the A stage of tile k+1 before the B stage of tile k was measured slower (17.0 vs 12.9 ms
      // per chunk): with two TMEM sets the A stage of tile k+1 has to wait for the whole epilogue of tile k-1, which only
      // starts when b5c(k-1) has been issued at the end of the previous iteration.  See DESIGN.md (appendix) for what a
      // faster schedule needs.
      for (int k = 0; k < my_tiles; ++k) {
        a_stage(k);
        b_stage(k);
      }
    }
  } else if (warp < 6) {
    // ---------------- pool warps: maxpool 1x3 over rows, in place, once p1 has read the slot ----------------
    const int pt = threadIdx.x - 64;
    const uint32_t c = (uint32_t)(pt & 7);
    const int r0 = (pt >> 3) * 8;
    int slot = 0;
    uint32_t ph = 0;
    for (int k = 0; k < my_tiles; ++k) {
      for (int j = 0; j < NKB; ++j) {
        const uint32_t tb = base + FS::ring + slot * X_SLOT;
        mbar_wait(bar(P1_DONE + slot), ph);               // implies the tile has landed in both CTAs
        uint4 in[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) {
          const int rr = r0 - 1 + q;
          if (rr >= 0 && rr < 128) in[q] = lds128(tb + (uint32_t)rr * 128u + ((c ^ (uint32_t)(rr & 7)) << 4));
          else in[q] = make_uint4(0u, 0u, 0u, 0u);
        }
        uint4 out[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          out[q].x = max16x2<F16>(max16x2<F16>(in[q].x, in[q + 1].x), in[q + 2].x);
          out[q].y = max16x2<F16>(max16x2<F16>(in[q].y, in[q + 1].y), in[q + 2].y);
          out[q].z = max16x2<F16>(max16x2<F16>(in[q].z, in[q + 1].z), in[q + 2].z);
          out[q].w = max16x2<F16>(max16x2<F16>(in[q].w, in[q + 1].w), in[q + 2].w);
        }
        named_bar_sync(1, 128);                          // every pool thread has finished reading its neighbours
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int rr = r0 + q;
          sts128(tb + (uint32_t)rr * 128u + ((c ^ (uint32_t)(rr & 7)) << 4), out[q]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(FULL_XP + slot));
        if (++slot == XS) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 14) {
    // ---------------- converters: b2 -> Y, planes -> smem, later relu(b5b) -> U ----------------
    const int q = warp & 3;
    const int h = (warp - 6) >> 2;
    const int i = q * 32 + lane;
    const uint32_t cstage = base + FS::u + (uint32_t)(warp - 6) * EPI_STAGE_BYTES;      // b2 staging inside the (idle) U region
    for (int k = 0; k < my_tiles; ++k) {
      const int a = k & 1;
      const uint32_t aph = (uint32_t)(k >> 1) & 1u, kph = (uint32_t)k & 1u;
      const int tile = tile_of(k);
      const int row = tile * F_TILE_OUT - 2 + i;
      const int wrow0 = tile * F_TILE_OUT - 2 + q * 32;
      const bool zero = row < 0 || row >= p.M || (row % p.Lp) >= p.L;
      const bool row_ok = i >= 2 && i < 2 + F_TILE_OUT && row < p.M;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * 256u;
      mbar_wait(bar(P1_ALL + a), aph);
      mbar_wait(bar(PL_EMPTY), kph ^ 1u);               // the previous tile's tap MMAs have read the plane tiles
      mbar_wait(bar(U_EMPTY), kph ^ 1u);                // ... and its b5c the U tile (this warp's b2 staging lives there)
      tc_fence_after();
      const uint32_t rr = (uint32_t)(i + 2);             // plane tile row (two rows down)
      const uint32_t sw = (rr >> 1) & 3u;
      auto plane = [&](int col0, int bias0, uint32_t tile_idx) {       // 32 accumulator columns -> one plane
        uint32_t r2[2][16];
        tc_ld16(t_row + (uint32_t)col0, r2[0]);
        tc_ld16(t_row + (uint32_t)(col0 + 16), r2[1]);
        tc_wait_ld();
        const uint32_t ub = base + FS::planes + tile_idx * PLANE_BYTES + rr * 64u;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t pk[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const int e = c4 * 8 + 2 * e2;
            float x0 = fmaxf(__uint_as_float(r2[e >> 4][e & 15]) + p.ba[bias0 + e], 0.f);
            float x1 = fmaxf(__uint_as_float(r2[(e + 1) >> 4][(e + 1) & 15]) + p.ba[bias0 + e + 1], 0.f);
            if (zero) { x0 = 0.f; x1 = 0.f; }
            pk[e2] = pack16<F16>(x0, x1);
          }
          sts128(ub + (((uint32_t)c4 ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      };
      if (h == 0) {
        // b2: accumulator columns 64..111 -> Y[:, 48:96)
        for (int part = 0; part < 2; ++part) {
          const int c0 = part * 32;
          const bool two = part == 0;
          uint32_t r2[2][16];
          tc_ld16(t_row + 64u + (uint32_t)c0, r2[0]);
          if (two) tc_ld16(t_row + 64u + (uint32_t)(c0 + 16), r2[1]);
          tc_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (h2 == 1 && !two) break;
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2) {
              const int e = 16 * h2 + 2 * e2;
              float x0 = fmaxf(__uint_as_float(r2[h2][2 * e2]) + p.ba[48 + c0 + e], 0.f);
              float x1 = fmaxf(__uint_as_float(r2[h2][2 * e2 + 1]) + p.ba[48 + c0 + e + 1], 0.f);
              if (zero) { x0 = 0.f; x1 = 0.f; }
              pk[8 * h2 + e2] = pack16<F16>(x0, x1);
            }
          }
          const size_t pitch = (size_t)p.ldY * 2;
          uint8_t* dst = reinterpret_cast<uint8_t*>(p.Y) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch + (size_t)(48 + c0) * 2;
          if (two) staged_store<4>(cstage, pk, lane, ok_mask, dst, pitch);
          else staged_store<2>(cstage, pk, lane, ok_mask, dst, pitch);
        }
        plane(112, 96, 1);                                // P3 (b3a)
      } else {
        plane(144, 128, 2);                               // P4 (b4a)
        plane(176, 160, 0);                               // P5 (b5a)
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PL_FULL + a));
      // ---- U = relu(b5b + bias): accumulator columns 48..111, 32 channels per converter half ----
      mbar_wait(bar(B5B_FULL + a), aph);
      tc_fence_after();
      {
        uint32_t r2[2][16];
        tc_ld16(t_row + 48u + (uint32_t)(h * 32), r2[0]);
        tc_ld16(t_row + 48u + (uint32_t)(h * 32 + 16), r2[1]);
        tc_wait_ld();
        const uint32_t usw = (uint32_t)((i >> 1) & 3);
        const uint32_t ub = base + FS::u + (uint32_t)h * U_KB + (uint32_t)i * 64u;
        // all b2 staging reads of this warp are complete (staged_store ends with __syncwarp), but OTHER warps may still be
        // reading their staging slices, which overlap this U k-block: wait for all converters
        named_bar_sync(2, 256);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t pk[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const int e = c4 * 8 + 2 * e2;
            float x0 = fmaxf(__uint_as_float(r2[e >> 4][e & 15]) + p.bb[144 + h * 32 + e], 0.f);
            float x1 = fmaxf(__uint_as_float(r2[(e + 1) >> 4][(e + 1) & 15]) + p.bb[144 + h * 32 + e + 1], 0.f);
            if (zero) { x0 = 0.f; x1 = 0.f; }
            pk[e2] = pack16<F16>(x0, x1);
          }
          sts128(ub + (((uint32_t)c4 ^ usw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(U_FULL));
    }
  } else {
    // ---------------- epilogue: b1 | b3 | b4 | relu(b5c + stem) -> Y ----------------
    const int q = warp & 3;
    const int h = (warp - 14) >> 2;
    const uint32_t stage = base + FS::stage + (uint32_t)(warp - 14) * EPI_STAGE_BYTES;
    const uint32_t ltp0 = map_to_cta(bar(TEMPTY_PAIR), 0);
    for (int k = 0; k < my_tiles; ++k) {
      const int a = k & 1;
      const uint32_t aph = (uint32_t)(k >> 1) & 1u;
      const int tile = tile_of(k);
      const int i = q * 32 + lane;
      const int row = tile * F_TILE_OUT - 2 + i;
      const int wrow0 = tile * F_TILE_OUT - 2 + q * 32;
      const bool row_ok = i >= 2 && i < 2 + F_TILE_OUT && row < p.M;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const bool halo = row < 0 || (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * 256u;
      mbar_wait(bar(TFULL + a), aph);
      tc_fence_after();
      // work units: (accumulator column, width, Y column, bias index in ba (>=0) or bb (-1-idx), stem column or -1)
      //  half 0: b1[0..32)  b3b4[0..32)  b3b4[64..96)  b5[32..48)      half 1: b1[32..48)  b3b4[32..64)  b5[0..32)
      for (int u = 0; u < 4; ++u) {
        int col, width, ycol, bidx, scol;
        bool useb;
        if (h == 0) {
          if (u == 0) { col = 0; width = 32; ycol = 0; bidx = 0; useb = false; scol = -1; }
          else if (u == 1) { col = 112; width = 32; ycol = 96; bidx = 0; useb = true; scol = -1; }
          else if (u == 2) { col = 176; width = 32; ycol = 160; bidx = 64; useb = true; scol = -1; }
          else { col = 80; width = 16; ycol = 224; bidx = 128; useb = true; scol = 240; }
        } else {
          if (u == 0) { col = 32; width = 16; ycol = 32; bidx = 32; useb = false; scol = -1; }
          else if (u == 1) { col = 144; width = 32; ycol = 128; bidx = 32; useb = true; scol = -1; }
          else if (u == 2) { col = 48; width = 32; ycol = 192; bidx = 96; useb = true; scol = 208; }
          else break;
        }
        const bool two = width == 32;
        uint32_t r2[2][16], rs[2][16];
        __syncwarp();
        tc_ld16(t_row + (uint32_t)col, r2[0]);
        if (two) tc_ld16(t_row + (uint32_t)(col + 16), r2[1]);
        if (scol >= 0) {
          tc_ld16(t_row + (uint32_t)scol, rs[0]);
          if (two) tc_ld16(t_row + (uint32_t)(scol + 16), rs[1]);
        }
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          if (h2 == 1 && !two) break;
#pragma unroll
          for (int e2 = 0; e2 < 8; ++e2) {
            float v[2];
#pragma unroll
            for (int z = 0; z < 2; ++z) {
              const int e = 16 * h2 + 2 * e2 + z;
              float x = __uint_as_float(r2[h2][2 * e2 + z]) + (useb ? p.bb[bidx + e] : p.ba[bidx + e]);
              if (scol >= 0) {
                // residual stem exactly as kernel A stores it (bias added, rounded to 16 bit, no ReLU)
                const float s = __uint_as_float(rs[h2][2 * e2 + z]) + p.ba[192 + (scol - 208) + e];
                const float2 sr = unpack16<F16>(pack16<F16>(s, 0.f));
                x += halo ? 0.f : sr.x;
              }
              v[z] = halo ? 0.f : fmaxf(x, 0.f);
            }
            pk[8 * h2 + e2] = pack16<F16>(v[0], v[1]);
          }
        }
        const size_t pitch = (size_t)p.ldY * 2;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.Y) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch + (size_t)ycol * 2;
        if (two) staged_store<4>(stage, pk, lane, ok_mask, dst, pitch);
        else staged_store<2>(stage, pk, lane, ok_mask, dst, pitch);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(TEMPTY + a)); arrive_cluster_relaxed(ltp0 + 8u * a); }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// Host side: the launch description is built from the two existing ones (same weights, same buffers).
struct InceptFLaunchImpl { FMaps maps; FParams p; bool fp16; };

void* make_incept_f(const InceptALaunch& A, const InceptBLaunch& B, const CUtensorMap& wp1_half, const CUtensorMap& wb1_full) {
  auto* L = new InceptFLaunchImpl();
  L->maps.x = A.maps.x; L->maps.wp1h = wp1_half; L->maps.wb1 = wb1_full;
  L->maps.w3 = B.maps.w3; L->maps.w4 = B.maps.w4; L->maps.w5b = B.maps.w5b; L->maps.w5c = B.maps.w5c;
  L->p.M = 0; L->p.Lp = A.p.Lp; L->p.L = A.p.L; L->p.Y = A.p.Y; L->p.ldY = A.p.ldY;
  for (int i = 0; i < 240; ++i) L->p.ba[i] = A.p.biasv[i];
  for (int i = 0; i < 208; ++i) L->p.bb[i] = B.p.biasv[i];
  L->fp16 = A.fp16;
  return L;
}
void free_incept_f(void* h) { delete static_cast<InceptFLaunchImpl*>(h); }

void launch_incept_f(const void* h, int M, int sm_count, cudaStream_t s) {
  if (M <= 0) return;
  const auto* L = static_cast<const InceptFLaunchImpl*>(h);
  const int num_tiles = (M + F_TILE_OUT - 1) / F_TILE_OUT;
  const int tile_pairs = (num_tiles + 1) / 2;
  int clusters = sm_count / 2;
  if (clusters > tile_pairs) clusters = tile_pairs;
  FParams p = L->p;
  p.M = M;
  const size_t smem = FS::total + 1024;
  static bool cfg[2][64] = {{false}};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  auto kern = L->fp16 ? incept_f_kernel<true> : incept_f_kernel<false>;
  if (!cfg[L->fp16 ? 1 : 0][dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg[L->fp16 ? 1 : 0][dev & 63] = true;
  }
  launch_pdl(kern, 2 * clusters, F_THREADS, smem, s, L->maps, p, num_tiles);
  DSB_CUDA(cudaGetLastError());
}

}  // namespace dsb
