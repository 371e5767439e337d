// incept_tc.cu — the two fused inception-block kernels (see incept_tc.cuh).  Both are persistent,
// warp-specialised tcgen05 kernels with every conv weight of their half of the block resident in shared
// memory, a TMA-fed ring of activation tiles, and double-buffered accumulators in TMEM.
#include <cstdlib>

#include "incept_tc.cuh"
#include "tc_common.cuh"

namespace dsb {
namespace {

using namespace tc;

template <bool F16> __device__ __forceinline__ uint32_t pack16(float a, float b) {
  if constexpr (F16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}
template <bool F16> __device__ __forceinline__ float2 unpack16(uint32_t w) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <bool F16> __device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
  if constexpr (F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}
// split-operand pool: elementwise max of three (hi, lo) pairs, two elements per 32-bit word.  hi is a monotone rounding of
// the value, so the winner has the largest hi and, among equal hi, the largest lo.
__device__ __forceinline__ void max3_pair(uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, uint32_t ch, uint32_t cl,
                                          uint32_t& oh, uint32_t& ol) {
  const __half2 A = *reinterpret_cast<const __half2*>(&ah), B = *reinterpret_cast<const __half2*>(&bh), C = *reinterpret_cast<const __half2*>(&ch);
  const __half2 H = __hmax2(__hmax2(A, B), C);
  const uint32_t NEG = 0xfc00fc00u;                      // -inf, -inf
  const uint32_t ma = __heq2_mask(A, H), mb = __heq2_mask(B, H), mc = __heq2_mask(C, H);
  const uint32_t la = (al & ma) | (NEG & ~ma), lb = (bl & mb) | (NEG & ~mb), lc = (cl & mc) | (NEG & ~mc);
  const __half2 L = __hmax2(__hmax2(*reinterpret_cast<const __half2*>(&la), *reinterpret_cast<const __half2*>(&lb)), *reinterpret_cast<const __half2*>(&lc));
  oh = *reinterpret_cast<const uint32_t*>(&H);
  ol = *reinterpret_cast<const uint32_t*>(&L);
}
// lo part of a packed pair of fp32 values whose hi parts are already packed in `hi`
__device__ __forceinline__ uint32_t pack_lo(float a, float b, uint32_t hi) {
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  return *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t idesc_f16(bool f16, int n) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// ============================================================================================
// Kernel A
// ============================================================================================
constexpr int A_THREADS = 448;       // warp 0 TMA, warp 1 MMA, warps 2-5 pool, warps 6-13 epilogue
#ifndef DSB_A_SLOTS
#define DSB_A_SLOTS 5
#endif
constexpr int A_SLOTS = DSB_A_SLOTS;
constexpr uint32_t A_SLOT_BYTES = 128 * 64 * 2;        // one [128 x 64] SWIZZLE_128B tile
constexpr uint32_t A_WP1_BLK = 192 * 64 * 2;           // 24576
constexpr uint32_t A_WB1_BLK = 48 * 64 * 2;            // 6144
constexpr int A_NKB = 4;
constexpr uint32_t A_ACC_STRIDE = 256;                 // TMEM columns per accumulator stage (240 used):
                                                       // b1 [0,48) b2 [48,96) b3a [96,128) b4a [128,160) b5a [160,192) stem [192,240)

struct ASmem {
  static constexpr uint32_t wp1 = 0;
  static constexpr uint32_t wb1 = wp1 + A_NKB * A_WP1_BLK;
  static constexpr uint32_t ring = wb1 + A_NKB * A_WB1_BLK;
  static constexpr uint32_t stage = ring + A_SLOTS * A_SLOT_BYTES;          // 8 epilogue warps x 2 KB
  static constexpr uint32_t bars = stage + 8 * EPI_STAGE_BYTES;
  // barriers: full_x[S] p1_done[S] full_xp[S] empty[S] wbar tfull[2] tempty[2]
  static constexpr uint32_t nbars = 4 * A_SLOTS + 5;
  static constexpr uint32_t tmem_slot = (bars + 8 * nbars + 15u) & ~15u;
  static constexpr uint32_t total = tmem_slot + 16;
};

template <bool F16, bool SPLIT>
__global__ void __launch_bounds__(A_THREADS, 1)
incept_a_kernel(const __grid_constant__ InceptAMaps maps, const __grid_constant__ InceptAParams p, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto full_x = [&](int s) { return base + ASmem::bars + 8u * s; };
  auto p1_done = [&](int s) { return base + ASmem::bars + 8u * (A_SLOTS + s); };
  auto full_xp = [&](int s) { return base + ASmem::bars + 8u * (2 * A_SLOTS + s); };
  auto empty = [&](int s) { return base + ASmem::bars + 8u * (3 * A_SLOTS + s); };
  const uint32_t wbar = base + ASmem::bars + 8u * (4 * A_SLOTS);
  auto tfull = [&](int a) { return base + ASmem::bars + 8u * (4 * A_SLOTS + 1 + a); };
  auto tempty = [&](int a) { return base + ASmem::bars + 8u * (4 * A_SLOTS + 3 + a); };
  const uint32_t tmem_slot = base + ASmem::tmem_slot;
  constexpr int PARTS = SPLIT ? 2 : 1;                   // split mode: every k-block arrives as a hi slot and a lo slot

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_SLOTS; ++s) { mbar_init(full_x(s), 1); mbar_init(p1_done(s), 1); mbar_init(full_xp(s), 4); mbar_init(empty(s), 1); }
    mbar_init(wbar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); }
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

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0 && (int)blockIdx.x < num_tiles) {
      mbar_expect_tx(wbar, A_NKB * (A_WP1_BLK + A_WB1_BLK));
      for (int j = 0; j < A_NKB; ++j) {
        tma_load_2d(base + ASmem::wp1 + j * A_WP1_BLK, &maps.wp1, wbar, j * 64, 0);
        tma_load_2d(base + ASmem::wb1 + j * A_WB1_BLK, &maps.wb1, wbar, j * 64, 0);
      }
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * INCEPT_A_TILE_ROWS - 1;
        for (int j = 0; j < A_NKB; ++j) {
#pragma unroll
          for (int part = 0; part < PARTS; ++part) {
            mbar_wait(empty(slot), phase ^ 1u);
            mbar_expect_tx(full_x(slot), A_SLOT_BYTES);
            tma_load_2d(base + ASmem::ring + slot * A_SLOT_BYTES, part == 0 ? &maps.x : &maps.xl, full_x(slot), j * 64, row0);
            if (++slot == A_SLOTS) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && (int)blockIdx.x < num_tiles) {
      const uint32_t id_p1 = idesc_f16(F16, 192), id_b1 = idesc_f16(F16, 48);
      mbar_wait(wbar, 0);
      int slot = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_b1 = tmem_base + acc * A_ACC_STRIDE, d_p1 = d_b1 + 48;
        // software-pipelined issue order p1(0) p1(1) b1(0) p1(2) b1(1) p1(3) b1(2) b1(3): a slot is released one
        // step after its p1 MMAs, so the TMA ring keeps most of its slots in flight.
        int sp = slot, sb = slot;
        uint32_t php = phase, phb = phase;
        for (int step = 0; step <= A_NKB; ++step) {
          if (step < A_NKB) {                              // p1: x tile straight from TMA (hi part, then lo part: same weights)
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              mbar_wait(full_x(sp), php);
              tc_fence_after();
              const uint64_t ad = make_kmajor_desc<64>(base + ASmem::ring + sp * A_SLOT_BYTES);
              const uint64_t bd = make_kmajor_desc<64>(base + ASmem::wp1 + step * A_WP1_BLK);
              if (!((p.ablate & 16) && part == 1)) {         // ablate 16 (timing only): no lo-part p1 MMAs
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_bf16(d_p1, ad + 2 * k, bd + 2 * k, id_p1, (step | part | k) != 0 ? 1u : 0u);
              }
              tc_commit(p1_done(sp));
              if (++sp == A_SLOTS) { sp = 0; php ^= 1u; }
            }
          }
          if (step >= 1) {                                 // b1: the same slot(s) after the pool warps rewrote them
            const int j = step - 1;
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              mbar_wait(full_xp(sb), phb);
              tc_fence_after();
              const uint64_t ad = make_kmajor_desc<64>(base + ASmem::ring + sb * A_SLOT_BYTES);
              const uint64_t bd = make_kmajor_desc<64>(base + ASmem::wb1 + j * A_WB1_BLK);
              if (!(p.ablate & 8)) {                         // ablate 8 (timing only): no b1 MMAs
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_bf16(d_b1, ad + 2 * k, bd + 2 * k, id_b1, (j | part | k) != 0 ? 1u : 0u);
              }
              tc_commit(empty(sb));
              if (++sb == A_SLOTS) { sb = 0; phb ^= 1u; }
            }
          }
        }
        int s = sp;
        uint32_t ph = php;
        slot = s;
        phase = ph;
        tc_commit(tfull(acc));
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int j = 0; j < A_NKB; ++j) {
        if constexpr (SPLIT) {
          // the k-block is a (hi, lo) pair of slots: pool the pairs (max by hi, then by lo), rewrite both tiles in place
          const int s2 = (slot + 1 == A_SLOTS) ? 0 : slot + 1;
          const uint32_t ph2 = (slot + 1 == A_SLOTS) ? (phase ^ 1u) : phase;
          const uint32_t th = base + ASmem::ring + slot * A_SLOT_BYTES, tl = base + ASmem::ring + s2 * A_SLOT_BYTES;
          mbar_wait(full_x(slot), phase);
          mbar_wait(full_x(s2), ph2);
          uint4 oh[8], ol[8];
          if (!(p.ablate & 1)) {
            auto ld = [&](uint32_t tb, int rr) {
              return (rr >= 0 && rr < 128) ? lds128(tb + (uint32_t)rr * 128u + ((c ^ (uint32_t)(rr & 7)) << 4)) : make_uint4(0u, 0u, 0u, 0u);
            };
            uint4 ph_ = ld(th, r0 - 1), pl_ = ld(tl, r0 - 1), ch_ = ld(th, r0), cl_ = ld(tl, r0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint4 nh = ld(th, r0 + k + 1), nl = ld(tl, r0 + k + 1);
              max3_pair(ph_.x, pl_.x, ch_.x, cl_.x, nh.x, nl.x, oh[k].x, ol[k].x);
              max3_pair(ph_.y, pl_.y, ch_.y, cl_.y, nh.y, nl.y, oh[k].y, ol[k].y);
              max3_pair(ph_.z, pl_.z, ch_.z, cl_.z, nh.z, nl.z, oh[k].z, ol[k].z);
              max3_pair(ph_.w, pl_.w, ch_.w, cl_.w, nh.w, nl.w, oh[k].w, ol[k].w);
              ph_ = ch_; pl_ = cl_; ch_ = nh; cl_ = nl;
            }
          }
          mbar_wait(p1_done(slot), phase);               // the p1 MMAs have finished reading both slots
          mbar_wait(p1_done(s2), ph2);
          named_bar_sync(1, 128);                        // every pool thread has finished reading its neighbours
          if (!(p.ablate & 1)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int rr = r0 + k;
              const uint32_t off = (uint32_t)rr * 128u + ((c ^ (uint32_t)(rr & 7)) << 4);
              sts128(th + off, oh[k]);
              sts128(tl + off, ol[k]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(full_xp(slot)); mbar_arrive(full_xp(s2)); }
          slot += 2;
          if (slot >= A_SLOTS) { slot -= A_SLOTS; phase ^= 1u; }
          continue;
        }
        const uint32_t tb = base + ASmem::ring + slot * A_SLOT_BYTES;
        mbar_wait(full_x(slot), phase);
        if (p.ablate & 1) { mbar_wait(p1_done(slot), phase); __syncwarp(); if (lane == 0) mbar_arrive(full_xp(slot)); if (++slot == A_SLOTS) { slot = 0; phase ^= 1u; } continue; }
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
        if (lane == 0) mbar_arrive(full_xp(slot));
        if (++slot == A_SLOTS) { slot = 0; phase ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const uint32_t stage = base + ASmem::stage + (uint32_t)(warp - 6) * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
        if (p.ablate & 4) break;                           // timing experiment: the epilogue only releases the accumulator
        const int c = g < 7 ? g * 32 : 208;
        uint32_t rr[2][16];
        __syncwarp();
        tc_ld16(t_row + (uint32_t)c, rr[0]);
        tc_ld16(t_row + (uint32_t)(c + 16), rr[1]);
        tc_wait_ld();
        uint32_t pk[16], pl[16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int col = c + 16 * h2;
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(rr[h2][k]) + p.biasv[col + k];
          if (g < 6) {                                    // the residual stem (g6, g7) keeps its sign
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = relu_keep_nan(v[k]);
          }
          if (halo) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) pk[8 * h2 + k] = pack16<F16>(v[2 * k], v[2 * k + 1]);
          if constexpr (SPLIT) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pl[8 * h2 + k] = pack_lo(v[2 * k], v[2 * k + 1], pk[8 * h2 + k]);
          }
        }
#pragma unroll
        for (int part = 0; part < PARTS; ++part) {
          const uint32_t (&pp)[16] = part == 0 ? pk : pl;
          // the staging tile may still be the source of this warp's previous TMA store
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          if (g < 2 || wrow0 < 0) {                       // (TMA stores reject negative coordinates: first warp of tile 0)
            uint8_t* dst;
            size_t pitch;
            if (g < 3) { pitch = (size_t)p.ldY * 2; dst = reinterpret_cast<uint8_t*>(p.Y) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch + g * 64 + (part ? p.lo_y * 2 : 0); }
            else if (g < 6) {
              pitch = 64;
              dst = reinterpret_cast<uint8_t*>(g == 3 ? p.P3 : g == 4 ? p.P4 : p.P5) + (ptrdiff_t)wrow0 * 64 + (part ? (g == 3 ? p.lo_p3 : g == 4 ? p.lo_p4 : p.lo_p5) * 2 : 0);
            } else { pitch = 96; dst = reinterpret_cast<uint8_t*>(p.PS) + (ptrdiff_t)wrow0 * 96 + (g == 7 ? 32 : 0) + (part ? p.lo_ps * 2 : 0); }
            staged_store<4>(stage, pp, lane, ok_mask, dst, pitch);
          } else {
            const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
            for (int pc = 0; pc < 4; ++pc)
              sts128(stage + (uint32_t)lane * 64u + (((uint32_t)pc ^ sw) << 4), make_uint4(pp[4 * pc], pp[4 * pc + 1], pp[4 * pc + 2], pp[4 * pc + 3]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(p.ablate & 2)) {
              const CUtensorMap* mp = part == 0 ? (g == 2 ? &maps.ys : g == 3 ? &maps.p3s : g == 4 ? &maps.p4s : g == 5 ? &maps.p5s : &maps.pss)
                                                : (g == 2 ? &maps.ysl : g == 3 ? &maps.p3sl : g == 4 ? &maps.p4sl : g == 5 ? &maps.p5sl : &maps.pssl);
              const int col = g == 2 ? 64 : g == 7 ? 16 : 0;
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(mp)), "r"(stage), "r"(col), "r"(wrow0) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory stays valid until every store has read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ============================================================================================
// Kernel B
// ============================================================================================
constexpr int B_THREADS = 448;       // warp 0 TMA, warp 1 MMA, warps 2-5 branch-5 staging, warps 6-13 epilogue
constexpr int B_TILE_ROWS = 136;                       // rows r0-2 .. r0+133 of one plane (taps reach r0-2 .. r0+129)
constexpr uint32_t B_SLOT_BYTES = B_TILE_ROWS * 32 * 2; // one [136 x 32] SWIZZLE_64B tile = 8704 B (17 x 512)
constexpr uint32_t B_U_BYTES = 128 * 32 * 2;            // one [128 x 32] k-block of the staged relu(b5b) tile
constexpr uint32_t B_W48_BLK = 48 * 32 * 2;            // 3072
constexpr uint32_t B_W64_BLK = 64 * 32 * 2;            // 4096
constexpr uint32_t B_ACC_STRIDE = 160;                 // columns: b3 [0,48) b4 [48,96) b5b [96,160), b5c re-uses [96,144)

// Split mode doubles every activation tile (hi + lo parts; the weights are shared): the U tiles take 96 KB for three
// accumulator stages, which leaves an 8-slot ring and one staging buffer per epilogue warp.
// Split-mode geometry of kernel B (build-time switches for A/B timing; profiles/r2_ablation.md): three accumulator stages
// (so the tap GEMMs of tile i+2 do not wait for the epilogue of tile i), an 8-slot ring and one staging buffer per epilogue
// warp measured 6.1 ms per 37,888-sample chunk against 6.9 ms for 2 stages / 9 slots / 2 buffers and 7.4 ms for 2 / 10 / 1.
#ifndef DSB_B_SPLIT_SLOTS
#define DSB_B_SPLIT_SLOTS 8
#endif
#ifndef DSB_B_SPLIT_ACC
#define DSB_B_SPLIT_ACC 3
#endif
#ifndef DSB_B_SPLIT_STG
#define DSB_B_SPLIT_STG 1          // staging buffers per epilogue warp in split mode: 2 = hi and lo side by side, 1 = one after the other
#endif
#ifndef DSB_B_SPLIT_USTG
#define DSB_B_SPLIT_USTG 3         // U tiles (hi + lo, 32 KB each) in split mode; a U tile is free once its b5c MMAs have completed
#endif
template <bool SPLIT>
struct BSmemT {
  static constexpr int PARTS = SPLIT ? 2 : 1;
  static constexpr int SLOTS = SPLIT ? DSB_B_SPLIT_SLOTS : 10;   // ring of plane tiles, 3 * PARTS per row tile
  static constexpr int ACC_STAGES = SPLIT ? DSB_B_SPLIT_ACC : 3;
  static constexpr int U_STAGES = SPLIT ? DSB_B_SPLIT_USTG : 3;
  static constexpr uint32_t w3 = 0;
  static constexpr uint32_t w4 = w3 + 3 * B_W48_BLK;
  static constexpr uint32_t w5b = w4 + 5 * B_W48_BLK;
  static constexpr uint32_t w5c = w5b + 3 * B_W64_BLK;
  static constexpr uint32_t u = w5c + 2 * B_W48_BLK;          // 43008: per stage PARTS x two [128 x 32] blocks
  static constexpr uint32_t ring = u + 2 * PARTS * U_STAGES * B_U_BYTES;
  static constexpr int STG = SPLIT ? DSB_B_SPLIT_STG : 2;                   // staging buffers per epilogue warp
  static constexpr uint32_t stage = ring + SLOTS * B_SLOT_BYTES;            // 8 epilogue warps x STG x 2 KB (TMA-store sources)
  static constexpr uint32_t bars = stage + 8 * STG * EPI_STAGE_BYTES;
  // barriers: full[S] empty[S] wbar b5b_full[A] u_full[A] tfull[A] tempty[A] u_empty[U]   (u_full indexed by U stage)
  static constexpr uint32_t nbars = 2 * SLOTS + 1 + 4 * ACC_STAGES + U_STAGES;
  static constexpr uint32_t tmem_slot = (bars + 8 * nbars + 15u) & ~15u;
  static constexpr uint32_t total = tmem_slot + 16;
  static_assert(total + 1024 <= 227 * 1024, "kernel B shared memory exceeds 227 KB");
  static_assert(ACC_STAGES * B_ACC_STRIDE <= 512, "kernel B accumulator stages exceed TMEM");
};

template <bool F16, bool SPLIT>
__global__ void __launch_bounds__(B_THREADS, 1)
incept_b_kernel(const __grid_constant__ InceptBMaps maps, const __grid_constant__ InceptBParams p, const int num_tiles) {
  using BSmem = BSmemT<SPLIT>;
  constexpr int PARTS = BSmem::PARTS, B_SLOTS = BSmem::SLOTS, B_ACC_STAGES = BSmem::ACC_STAGES, U_STAGES = BSmem::U_STAGES;
  static_assert(U_STAGES <= B_ACC_STAGES, "at most one U tile per accumulator stage");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto full = [&](int s) { return base + BSmem::bars + 8u * s; };
  auto empty = [&](int s) { return base + BSmem::bars + 8u * (B_SLOTS + s); };
  const uint32_t wbar = base + BSmem::bars + 8u * (2 * B_SLOTS);
  auto b5b_full = [&](int a) { return base + BSmem::bars + 8u * (2 * B_SLOTS + 1 + a); };
  auto u_full = [&](int a) { return base + BSmem::bars + 8u * (2 * B_SLOTS + 1 + B_ACC_STAGES + a); };
  auto tfull = [&](int a) { return base + BSmem::bars + 8u * (2 * B_SLOTS + 1 + 2 * B_ACC_STAGES + a); };
  auto tempty = [&](int a) { return base + BSmem::bars + 8u * (2 * B_SLOTS + 1 + 3 * B_ACC_STAGES + a); };
  auto u_empty = [&](int a) { return base + BSmem::bars + 8u * (2 * B_SLOTS + 1 + 4 * B_ACC_STAGES + a); };
  const uint32_t tmem_slot = base + BSmem::tmem_slot;
  // U tile of accumulator stage a, part (0 hi / 1 lo), k-block kb
  auto u_blk = [&](int a, int part, int kb) { return base + BSmem::u + (uint32_t)((a * PARTS + part) * 2 + kb) * B_U_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < B_SLOTS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(wbar, 1);
    for (int a = 0; a < B_ACC_STAGES; ++a) { mbar_init(b5b_full(a), 1); mbar_init(u_full(a), 4); mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); }
    for (int a = 0; a < U_STAGES; ++a) mbar_init(u_empty(a), 1);
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

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0 && (int)blockIdx.x < num_tiles) {
      mbar_expect_tx(wbar, 3 * B_W48_BLK + 5 * B_W48_BLK + 3 * B_W64_BLK + 2 * B_W48_BLK);
      for (int t = 0; t < 3; ++t) tma_load_2d(base + BSmem::w3 + t * B_W48_BLK, &maps.w3, wbar, t * 32, 0);
      for (int t = 0; t < 5; ++t) tma_load_2d(base + BSmem::w4 + t * B_W48_BLK, &maps.w4, wbar, t * 32, 0);
      for (int t = 0; t < 3; ++t) tma_load_2d(base + BSmem::w5b + t * B_W64_BLK, &maps.w5b, wbar, t * 32, 0);
      for (int t = 0; t < 2; ++t) tma_load_2d(base + BSmem::w5c + t * B_W48_BLK, &maps.w5c, wbar, t * 32, 0);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * 128;
        // one [136 x 32] tile per plane (and part), rows r0-2 ..., in the order of consumption P5, P3, P4.  The taps of the
        // 1x3 / 1x5 convs are realised by the MMA issuer with row-shifted operand descriptors on this single tile
        // (the UMMA swizzle is a function of the absolute shared-memory address: profiles/experiments/umma_shift.cu).
        for (int l = 0; l < 3; ++l) {
#pragma unroll
          for (int part = 0; part < PARTS; ++part) {
            const CUtensorMap* mp = part == 0 ? (l == 0 ? &maps.t5 : l == 1 ? &maps.t3 : &maps.t4)
                                              : (l == 0 ? &maps.t5l : l == 1 ? &maps.t3l : &maps.t4l);
            mbar_wait(empty(slot), phase ^ 1u);
            mbar_expect_tx(full(slot), B_SLOT_BYTES);
            tma_load_2d(base + BSmem::ring + slot * B_SLOT_BYTES, mp, full(slot), 0, row0 - 2);
            if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && (int)blockIdx.x < num_tiles) {
      const uint32_t id48 = idesc_f16(F16, 48), id64 = idesc_f16(F16, 64);
      mbar_wait(wbar, 0);
      int slot = 0;
      uint32_t phase = 0;
      // all taps of one branch read the same plane tile at row offsets first_row .. first_row + ntaps - 1; in split
      // mode the lo tile follows the hi tile and multiplies the same weight blocks
      auto run = [&](int ntaps, int first_row, uint32_t w_off, uint32_t w_blk, uint32_t d, uint32_t idesc) {
#pragma unroll
        for (int part = 0; part < PARTS; ++part) {
          mbar_wait(full(slot), phase);
          tc_fence_after();
          const uint32_t tile = base + BSmem::ring + slot * B_SLOT_BYTES;
          for (int t = 0; t < ntaps; ++t) {
            const uint64_t ad = make_kmajor_desc<32>(tile + (uint32_t)(first_row + t) * 64u);
            const uint64_t bd = make_kmajor_desc<32>(base + w_off + t * w_blk);
#pragma unroll
            for (int k = 0; k < 2; ++k) tc_mma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (part | t | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty(slot));
          if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
        }
      };
      // tap GEMMs of one tile into accumulator stage `a`: branch 5 first so its staging starts early
      auto front = [&](int a, uint32_t a_phase) {
        mbar_wait(tempty(a), a_phase ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + a * B_ACC_STRIDE;
        run(3, 1, BSmem::w5b, B_W64_BLK, d0 + 96, id64);     // 1x3: rows r0-1 .. r0+1  (tile row 0 = r0-2)
        tc_commit(b5b_full(a));
        run(3, 1, BSmem::w3, B_W48_BLK, d0, id48);
        run(5, 0, BSmem::w4, B_W48_BLK, d0 + 48, id48);        // 1x5: rows r0-2 .. r0+2
      };
      // software pipeline across tiles: the tap GEMMs of tile i+1 are issued before waiting for the staged
      // relu(b5b) tile of tile i, so the TMEM -> smem round trip of branch 5 is off the critical path.
      int acc = 0, k = 0;
      uint32_t acc_phase = 0;
      if ((int)blockIdx.x < num_tiles) front(0, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
        const int nacc = (acc + 1 == B_ACC_STAGES) ? 0 : acc + 1;
        const uint32_t nphase = (nacc == 0) ? (acc_phase ^ 1u) : acc_phase;
        if (tile + (int)gridDim.x < num_tiles) front(nacc, nphase);
        const uint32_t d0 = tmem_base + acc * B_ACC_STRIDE;
        const int ub = k % U_STAGES;                      // U tile of this CTA's k-th row tile
        mbar_wait(u_full(ub), (uint32_t)(k / U_STAGES) & 1u);   // relu(b5b) tile is in shared memory
        tc_fence_after();
#pragma unroll
        for (int part = 0; part < PARTS; ++part) {
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t ad = make_kmajor_desc<32>(u_blk(ub, part, kb));
            const uint64_t bd = make_kmajor_desc<32>(base + BSmem::w5c + kb * B_W48_BLK);
#pragma unroll
            for (int k = 0; k < 2; ++k) tc_mma_bf16(d0 + 96, ad + 2 * k, bd + 2 * k, id48, (part | kb | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(u_empty(ub));                           // the U tile may be rewritten once these MMAs have read it
        tc_commit(tfull(acc));
        acc = nacc;
        acc_phase = nphase;
      }
    }
  } else if (warp < 6) {
    // ---------------- branch-5 staging warps: relu(b5b + bias) -> 16-bit A operand of the 1x1 that follows ----------------
    const int q = warp & 3;
    int acc = 0, k = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
      const int i = q * 32 + lane;
      const int row = tile * 128 + i;
      const bool zero = row >= p.M || (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * B_ACC_STRIDE + 96u;
      const int ub = k % U_STAGES;
      mbar_wait(b5b_full(acc), acc_phase);
      tc_fence_after();
      uint32_t rr[4][16];
#pragma unroll
      for (int g = 0; g < 4; ++g) tc_ld16(t_row + 16u * g, rr[g]);
      tc_wait_ld();
      mbar_wait(u_empty(ub), ((uint32_t)(k / U_STAGES) & 1u) ^ 1u);     // the b5c MMAs that read this U tile have completed
      const uint32_t sw = (uint32_t)((i >> 1) & 3);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        // k-block kb of U: row i is 64 bytes; logical 16-byte chunk c sits at c ^ ((i >> 1) & 3)  (SWIZZLE_64B)
        const uint32_t ubh = u_blk(ub, 0, kb) + (uint32_t)i * 64u;
        const uint32_t ubl = u_blk(ub, PARTS - 1, kb) + (uint32_t)i * 64u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[4], pl[4];
          const int e0 = kb * 32 + c * 8;                      // first b5b output channel of this 16-byte chunk
          const float bv[8] = {p.biasv[144 + e0], p.biasv[144 + e0 + 1], p.biasv[144 + e0 + 2], p.biasv[144 + e0 + 3],
                               p.biasv[144 + e0 + 4], p.biasv[144 + e0 + 5], p.biasv[144 + e0 + 6], p.biasv[144 + e0 + 7]};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = e0 + 2 * k;
            float a = relu_keep_nan(__uint_as_float(rr[e >> 4][e & 15]) + bv[2 * k]);
            float b = relu_keep_nan(__uint_as_float(rr[(e + 1) >> 4][(e + 1) & 15]) + bv[2 * k + 1]);
            if (zero) { a = 0.f; b = 0.f; }
            pk[k] = pack16<F16>(a, b);
            if constexpr (SPLIT) pl[k] = pack_lo(a, b, pk[k]);
          }
          sts128(ubh + (((uint32_t)c ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          if constexpr (SPLIT) sts128(ubl + (((uint32_t)c ^ sw) << 4), make_uint4(pl[0], pl[1], pl[2], pl[3]));
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(u_full(ub));
      if (++acc == B_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ---------------- epilogue warps: b3, b4, b5 = relu(stem + b5c) -> Y[:, 96:240] ----------------
    // accumulator column t (b3 [0,48) | b4 [48,96) | b5c [96,144)) is output channel 96 + t and bias index t.
    // Five 32-column groups per row tile: t0 = 0, 32, 64, 96 and 112 (the last one overlaps the fourth by 16 columns so
    // that every group is a full [32 rows x 32 columns] box).  A warp converts its 32 rows of a group, writes them as a
    // SWIZZLE_64B tile into one of its two staging buffers and issues one TMA store: whole 64-byte row segments, no
    // shared-memory read-back by the threads.  Rows >= M inside the tensor are written too (never consumed).
    // Split mode: the hi tile goes through staging buffer 0, the lo tile through buffer 1.
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const uint32_t stage = base + BSmem::stage + (uint32_t)(warp - 6) * (uint32_t)BSmem::STG * EPI_STAGE_BYTES;
    uint32_t stage_sel = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int i = q * 32 + lane;
      const int row = tile * 128 + i;
      const int wrow0 = tile * 128 + q * 32;
      const bool halo = (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * B_ACC_STRIDE;
      // the residual (stem plane) is an input: fetch it before waiting for the accumulators.
      // warp set 1 owns group 3 (stem channels 0..31), warp set 0 owns group 4 (stem channels 16..47).
      uint4 res[4], resl[4];
      {
        const uint4* tp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.PS) + (size_t)(row < p.M ? row : 0) * 48);
        const int o = half == 1 ? 0 : 2;
        res[0] = __ldg(tp + o); res[1] = __ldg(tp + o + 1); res[2] = __ldg(tp + o + 2); res[3] = __ldg(tp + o + 3);
        if constexpr (SPLIT) {
          const uint4* tl = tp + p.lo_ps / 8;
          resl[0] = __ldg(tl + o); resl[1] = __ldg(tl + o + 1); resl[2] = __ldg(tl + o + 2); resl[3] = __ldg(tl + o + 3);
        }
      }
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      for (int g = half; g < 5; g += 2) {
        const int t0 = g < 4 ? g * 32 : 112;
        uint32_t rr[2][16];
        __syncwarp();
        tc_ld16(t_row + (uint32_t)t0, rr[0]);
        tc_ld16(t_row + (uint32_t)(t0 + 16), rr[1]);
        tc_wait_ld();
        uint32_t pk[16], pl[16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int t = t0 + 16 * h2;
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(rr[h2][k]) + p.biasv[t + k];
          if (g >= 3) {                                   // + stem
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint4 rv = res[2 * h2 + hh];
              const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = unpack16<F16>(w[k]);
                v[8 * hh + 2 * k] += f.x;
                v[8 * hh + 2 * k + 1] += f.y;
              }
              if constexpr (SPLIT) {
                const uint4 lv = resl[2 * h2 + hh];
                const uint32_t wl[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = unpack16<F16>(wl[k]);
                  v[8 * hh + 2 * k] += f.x;
                  v[8 * hh + 2 * k + 1] += f.y;
                }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = halo ? 0.f : relu_keep_nan(v[k]);
#pragma unroll
          for (int k = 0; k < 8; ++k) pk[8 * h2 + k] = pack16<F16>(v[2 * k], v[2 * k + 1]);
          if constexpr (SPLIT) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pl[8 * h2 + k] = pack_lo(v[2 * k], v[2 * k + 1], pk[8 * h2 + k]);
          }
        }
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);
        if constexpr (SPLIT && BSmem::STG == 1) {
          // one staging buffer: the hi tile and the lo tile go through it one after the other
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            const uint32_t (&pp)[16] = part == 0 ? pk : pl;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int pc = 0; pc < 4; ++pc)
              sts128(stage + (uint32_t)lane * 64u + (((uint32_t)pc ^ sw) << 4), make_uint4(pp[4 * pc], pp[4 * pc + 1], pp[4 * pc + 2], pp[4 * pc + 3]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(p.ablate & 2)) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(part == 0 ? &maps.ys : &maps.ysl)), "r"(stage), "r"(96 + t0), "r"(wrow0) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        } else if constexpr (SPLIT) {
          // both staging buffers were the sources of the previous group's two stores
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            sts128(stage + (uint32_t)lane * 64u + (((uint32_t)pc ^ sw) << 4), make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]));
            sts128(stage + EPI_STAGE_BYTES + (uint32_t)lane * 64u + (((uint32_t)pc ^ sw) << 4), make_uint4(pl[4 * pc], pl[4 * pc + 1], pl[4 * pc + 2], pl[4 * pc + 3]));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.ablate & 2)) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&maps.ys)), "r"(stage), "r"(96 + t0), "r"(wrow0) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&maps.ysl)), "r"(stage + EPI_STAGE_BYTES), "r"(96 + t0), "r"(wrow0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          // the buffer about to be rewritten was the source of the store issued two groups ago
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          const uint32_t stg = stage + stage_sel * EPI_STAGE_BYTES;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
            sts128(stg + (uint32_t)lane * 64u + (((uint32_t)pc ^ sw) << 4), make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.ablate & 2)) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&maps.ys)), "r"(stg), "r"(96 + t0), "r"(wrow0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          stage_sel ^= 1u;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      if (++acc == B_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory stays valid until every store has read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <class K> void set_smem_once(K kern, bool* flags) {
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  if (!flags[dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    flags[dev & 63] = true;
  }
}

}  // namespace

void launch_incept_a(const InceptALaunch& L, int M, int sm_count, int ablate, cudaStream_t s) {
  if (M <= 0) return;
  const int num_tiles = (M + INCEPT_A_TILE_ROWS - 1) / INCEPT_A_TILE_ROWS;
  const int grid = num_tiles < sm_count ? num_tiles : sm_count;
  InceptAParams p = L.p;
  p.M = M;
  p.ablate = ablate;
  const size_t smem = ASmem::total + 1024;
  static bool cfg[3][64] = {{false}};
  if (L.split) {
    if (!L.fp16) fail(DSB_ERR_INVALID, "the split-operand mode needs fp16 operands");
    set_smem_once(incept_a_kernel<true, true>, cfg[2]);
    launch_pdl(incept_a_kernel<true, true>, grid, A_THREADS, smem, s, L.maps, p, num_tiles);
  } else if (L.fp16) {
    set_smem_once(incept_a_kernel<true, false>, cfg[1]);
    launch_pdl(incept_a_kernel<true, false>, grid, A_THREADS, smem, s, L.maps, p, num_tiles);
  } else {
    set_smem_once(incept_a_kernel<false, false>, cfg[0]);
    launch_pdl(incept_a_kernel<false, false>, grid, A_THREADS, smem, s, L.maps, p, num_tiles);
  }
  DSB_CUDA(cudaGetLastError());
}

void launch_incept_b(const InceptBLaunch& L, int M, int sm_count, int ablate, cudaStream_t s) {
  if (M <= 0) return;
  const int num_tiles = (M + 127) / 128;
  const int grid = num_tiles < sm_count ? num_tiles : sm_count;
  InceptBParams p = L.p;
  p.M = M;
  p.ablate = ablate;
  static bool cfg[3][64] = {{false}};
  if (L.split) {
    if (!L.fp16) fail(DSB_ERR_INVALID, "the split-operand mode needs fp16 operands");
    set_smem_once(incept_b_kernel<true, true>, cfg[2]);
    launch_pdl(incept_b_kernel<true, true>, grid, B_THREADS, BSmemT<true>::total + 1024, s, L.maps, p, num_tiles);
  } else if (L.fp16) {
    set_smem_once(incept_b_kernel<true, false>, cfg[1]);
    launch_pdl(incept_b_kernel<true, false>, grid, B_THREADS, BSmemT<false>::total + 1024, s, L.maps, p, num_tiles);
  } else {
    set_smem_once(incept_b_kernel<false, false>, cfg[0]);
    launch_pdl(incept_b_kernel<false, false>, grid, B_THREADS, BSmemT<false>::total + 1024, s, L.maps, p, num_tiles);
  }
  DSB_CUDA(cudaGetLastError());
}

}  // namespace dsb
