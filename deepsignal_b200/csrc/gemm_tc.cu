// gemm_tc.cu — bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA) with the conv / LSTM epilogues fused.  sm_100a only; this is the DSB_PREC_BF16 path.
//
// One persistent CTA per SM, 320 threads:
//   warp 0      TMA producer  (one elected lane): A tile [128 x BK] + B tile [BN x BK] per stage
//   warp 1      TMEM allocator + MMA issuer (one elected lane): BK/16 tcgen05.mma per stage,
//               tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..9  epilogue (two warps per TMEM lane quadrant): tcgen05.ld 32 lanes x 16 columns, bias / residual / ReLU / halo-row zeroing
//               (conv) or the LSTM cell update, bf16 stores
// Accumulators are double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  A-operand "segments" give k-tap convs (row-shifted TMA boxes; out-of-range rows are
// zero-filled by TMA, sample edges read the zero halo rows) and [x_t | h_{t-1}] for the LSTM.
#include <algorithm>
#include <cstdlib>

#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace dsb {
namespace {

using namespace tc;

constexpr int TC_THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps
constexpr int BM_TC = 128;

template <int BK, int EPI, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const GemmTCParams p, const int m_tiles, const uint32_t tmem_cols, const int ablate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t a_bytes = BM_TC * BK * 2, b_bytes = (uint32_t)BN * BK * 2;
  const bool resident = p.resident != 0;
  int total_kb = 0;
  for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;
  const uint32_t stage_bytes = resident ? a_bytes : a_bytes + b_bytes;
  const uint32_t bres_base = smem_base;                                   // resident B: total_kb blocks of b_bytes
  const uint32_t ring_base = smem_base + (resident ? (uint32_t)p.wkb_total * b_bytes : 0u);
  const int stages = p.stages;
  const uint32_t bar_base = ring_base + stages * stage_bytes;   // 8-byte barriers
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t bres_bar = bar_base + 8u * (2 * stages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 6);
  float* sbias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));   // [N] floats
  const uint32_t stage_base = (tmem_slot + 16u + (uint32_t)p.epi.N * 4u + 15u) & ~15u;           // 8 x 2 KB epilogue staging
  const int n_tiles_n = p.epi.N / BN;
  // tile walk: streaming mode interleaves (m, n) tiles over all CTAs; resident mode pins a CTA to one n tile
  const int my_n = resident ? (int)(blockIdx.x % n_tiles_n) : 0;
  const int t_begin = resident ? (int)(blockIdx.x / n_tiles_n) : (int)blockIdx.x;
  const int t_step = resident ? (int)(gridDim.x / n_tiles_n) : (int)gridDim.x;
  const int num_tiles = resident ? m_tiles : m_tiles * n_tiles_n;
  auto tile_m = [&](int t) { return resident ? t : t / n_tiles_n; };
  auto tile_n = [&](int t) { return resident ? my_n : t % n_tiles_n; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.epi.N; i += TC_THREADS) sbias[i] = p.epi.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait_prior_grid();        // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (resident && t_begin < num_tiles) {
        mbar_expect_tx(bres_bar, (uint32_t)p.wkb_total * b_bytes);
        for (int kb = 0; kb < p.wkb_total; ++kb) tma_load_2d(bres_base + kb * b_bytes, &maps.b, bres_bar, kb * BK, my_n * BN);
      }
      for (int tile = t_begin; tile < num_tiles; tile += t_step) {
        const int row0 = tile_m(tile) * BM_TC, n0 = tile_n(tile) * BN;
        for (int s = 0; s < p.nseg; ++s) {
          const SegTC sg = p.seg[s];
          for (int j = 0; j < sg.nkb; ++j) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), stage_bytes);
            const uint32_t a_dst = ring_base + stage * stage_bytes;
            tma_load_2d(a_dst, &maps.a[sg.map], full_bar(stage), sg.col0 + j * BK, row0 + sg.drow);
            if (!resident) tma_load_2d(a_dst + a_bytes, &maps.b, full_bar(stage), (sg.wkb + j) * BK, n0);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A/B format at bits 7-9 / 10-12, K-major A/B,
      // N>>3 at bits 17-22, M>>4 at bits 24-28.
      const uint32_t fmt = F16 ? 0u : 1u;   // F16F32Format: 0 = F16, 1 = BF16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM_TC >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      if (resident && t_begin < num_tiles) mbar_wait(bres_bar, 0);
      for (int tile = t_begin; tile < num_tiles; tile += t_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int wkb0 = p.seg[s].wkb, nkb = p.seg[s].nkb;
          for (int j = 0; j < nkb; ++j, ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = ring_base + stage * stage_bytes;
            const uint64_t adesc = make_kmajor_desc<BK>(a_addr);
            const uint64_t bdesc = make_kmajor_desc<BK>(resident ? bres_base + (uint32_t)(wkb0 + j) * b_bytes : a_addr + a_bytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
        tc_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ================= epilogue warps (8: two per TMEM lane quadrant, interleaved column groups) =================
    const int q = warp & 3;                   // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;         // which of the two interleaved column-group sets
    const EpiParams& e = p.epi;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = t_begin; tile < num_tiles; tile += t_step) {
      const int row = tile_m(tile) * BM_TC + q * 32 + lane;
      const int n0 = tile_n(tile) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool row_ok = row < e.M && !(ablate & 2);
      const uint32_t stage = stage_base + (uint32_t)(warp - 2) * EPI_STAGE_BYTES;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const int wrow0 = tile_m(tile) * BM_TC + q * 32;
      if constexpr (EPI == EPI_CONV) {
        const bool halo = e.Lp > 0 && (row % e.Lp) >= e.L;
        // staging pays off for narrow row pitches (fused kernels' planes); for the 256/512-byte pitches of conv2/conv3 the
        // direct 2 x 16-byte stores per lane measured faster, so the generic conv epilogue keeps them.
        const bool staged = false;
        // 32-column groups; group g belongs to the warp set (g & 1) == half
        for (int c = half * 32; c < BN; c += 64) {
          const bool two = c + 32 <= BN;       // BN is a multiple of 16: the last group may hold 16 columns
          uint32_t r[2][16];
          __syncwarp();
          tc_ld16(t_row + (uint32_t)c, r[0]);
          if (two) tc_ld16(t_row + (uint32_t)(c + 16), r[1]);
          tc_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (h2 == 1 && !two) break;
            const int col = n0 + c + 16 * h2;
            float v[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bb = *reinterpret_cast<const float4*>(sbias + col + 4 * i);
              v[4 * i + 0] = __uint_as_float(r[h2][4 * i + 0]) + bb.x;
              v[4 * i + 1] = __uint_as_float(r[h2][4 * i + 1]) + bb.y;
              v[4 * i + 2] = __uint_as_float(r[h2][4 * i + 2]) + bb.z;
              v[4 * i + 3] = __uint_as_float(r[h2][4 * i + 3]) + bb.w;
            }
            const bool first = col < e.split;
            if (e.res != nullptr && first && row_ok) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(e.res) + (size_t)row * e.res_ld + e.res_c0 + col);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint4 rv = __ldg(rp + h);
                const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack16<F16>(w[i]);
                  v[8 * h + 2 * i] += f.x;
                  v[8 * h + 2 * i + 1] += f.y;
                }
                if (e.res_lo_off != 0) {                   // split-operand mode: the residual is hi + lo
                  const uint4 lv = __ldg(rp + h + e.res_lo_off / 8);
                  const uint32_t wl[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack16<F16>(wl[i]);
                    v[8 * h + 2 * i] += f.x;
                    v[8 * h + 2 * i + 1] += f.y;
                  }
                }
              }
            }
            if ((e.relu_mask >> (col >> 4)) & 1u) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = relu_keep_nan(v[i]);
            }
            if (halo) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[8 * h2 + i] = pack16<F16>(v[2 * i], v[2 * i + 1]);
            if (!staged && row_ok) {
              uint16_t* dst = first ? reinterpret_cast<uint16_t*>(e.d0) + (size_t)row * e.ld0 + e.c0 + col
                                    : reinterpret_cast<uint16_t*>(e.d1) + (size_t)row * e.ld1 + e.c1 + col - e.split;
              uint4* dp = reinterpret_cast<uint4*>(dst);
              dp[0] = make_uint4(pk[8 * h2 + 0], pk[8 * h2 + 1], pk[8 * h2 + 2], pk[8 * h2 + 3]);
              dp[1] = make_uint4(pk[8 * h2 + 4], pk[8 * h2 + 5], pk[8 * h2 + 6], pk[8 * h2 + 7]);
              const long long lo_off = first ? e.lo_off0 : e.lo_off1;
              if (lo_off != 0) {                           // lo = fp16(v - hi): the part of v the 16-bit value lost
                uint32_t pl[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float2 hf = unpack16<F16>(pk[8 * h2 + i]);
                  pl[i] = pack16<F16>(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
                }
                uint4* lp = reinterpret_cast<uint4*>(dst + lo_off);
                lp[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                lp[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
              }
            }
          }
          if (staged) {
            const int col = n0 + c;
            const bool first = col < e.split;
            const size_t pitch = (size_t)(first ? e.ld0 : e.ld1) * 2;
            uint8_t* dst = reinterpret_cast<uint8_t*>(first ? e.d0 : e.d1) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch +
                           (size_t)(first ? e.c0 + col : e.c1 + col - e.split) * 2;
            if (two) staged_store<4>(stage, pk, lane, ok_mask, dst, pitch);
            else staged_store<2>(stage, pk, lane, ok_mask, dst, pitch);
          }
        }
      } else {
        // LSTM: columns 4*u+g.  64 columns (16 hidden units) per iteration, iterations interleaved over the two warp sets.
        // The cell state is private to this kernel, so it is kept in a thread-interleaved layout: block
        // (m tile, 64-column iteration, lane quadrant) holds 4 float4 per lane, lane-minor -> every load/store
        // instruction of a warp touches 512 contiguous bytes.
        for (int c = half * 64; c < BN; c += 128) {
          uint32_t r[4][16];
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) tc_ld16(t_row + (uint32_t)(c + 16 * i), r[i]);
          const int col = n0 + c;
          const int u0 = col >> 2;
          float4* cblk = reinterpret_cast<float4*>(e.c) + ((((size_t)tile_m(tile) * (GATES / 64) + (size_t)(col >> 6)) * 4 + q) * 4) * 32 + lane;
          float4 cprev[4];
          if (!(ablate & 2)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cprev[i] = cblk[i * 32];
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) cprev[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
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
            if (!(ablate & 2)) cblk[i * 32] = make_float4(cn[0], cn[1], cn[2], cn[3]);
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
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || sym == nullptr) fail(DSB_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

template <int BK, int EPI, bool F16>
void launch_one(const TcLaunch& L, int m_tiles, int grid, uint32_t tmem_cols, int ablate, cudaStream_t s) {
  auto kern = gemm_tc_kernel<BK, EPI, F16>;
  static bool configured[64] = {false};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[dev & 63] = true;
  }
  launch_pdl(kern, grid, TC_THREADS, L.smem, s, L.maps, L.p, m_tiles, tmem_cols, ablate);
}

}  // namespace

void encode_map_16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                   int box_rows, int box_cols, bool fp16) {
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (box_cols * 2 == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (sw == CU_TENSOR_MAP_SWIZZLE_NONE) fail(DSB_ERR_INVALID, "tensor map box must span 64 or 128 bytes");
  if (((uintptr_t)base & 15) || ((ld * 2) & 15)) fail(DSB_ERR_INVALID, "tensor map base/stride must be 16-byte aligned");
  CUresult r = get_encode_fn()(out, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(DSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld box=%dx%d)",
                              (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
}

void finalize_tc_launch(TcLaunch& L) {
  if (L.p.BN % 16 != 0 || L.p.BN < 16 || L.p.BN > 256 || L.p.epi.N % L.p.BN != 0)
    fail(DSB_ERR_INVALID, "bad GEMM tile: N=%d BN=%d", L.p.epi.N, L.p.BN);
  int total_kb = 0;                                   // weight blocks (segments may share them)
  for (int s = 0; s < L.p.nseg; ++s) total_kb = std::max(total_kb, L.p.seg[s].wkb + L.p.seg[s].nkb);
  L.p.wkb_total = total_kb;
  const size_t a_bytes = (size_t)BM_TC * L.BK * 2;
  const size_t budget = 196 * 1024;
  // conv weights are small: keep the whole [BN x K] operand in shared memory when it leaves room for >= 4 A stages
  L.p.resident = 0;
  if (L.epi_kind == EPI_CONV) {
    int bn = L.p.BN;
    while (true) {
      const size_t bres = (size_t)total_kb * bn * L.BK * 2;
      if (bres + 4 * a_bytes <= budget && (148 % (L.p.epi.N / bn)) == 0) { L.p.BN = bn; L.p.resident = 1; break; }
      if (bn % 32 != 0 || L.p.epi.N % (bn / 2) != 0) break;
      bn /= 2;
    }
  }
  const size_t b_bytes = (size_t)L.p.BN * L.BK * 2;
  const size_t bres = L.p.resident ? (size_t)total_kb * b_bytes : 0;
  const size_t stage_bytes = L.p.resident ? a_bytes : a_bytes + b_bytes;
  int stages = (int)((budget - bres) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  L.p.stages = stages;
  size_t smem = 1024 + bres + (size_t)stages * stage_bytes + 8 * (2 * stages + 6) + 16 + (size_t)L.p.epi.N * 4 + 32 + 8 * 2048;
  // 2*BN TMEM columns per CTA: keep co-resident CTAs within the 512-column budget
  const int cols = 2 * L.p.BN;
  if (cols > 256 && smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM
  else if (cols > 128 && smem < 80 * 1024) smem = 80 * 1024;   // at most two
  L.smem = smem;
}

void launch_gemm_tc(const TcLaunch& L, int M, int sm_count, int ablate, cudaStream_t s) {
  const int m_tiles = (M + BM_TC - 1) / BM_TC;
  const int ntn = L.p.epi.N / L.p.BN;
  const int num_tiles = m_tiles * ntn;
  int grid = num_tiles < sm_count ? num_tiles : sm_count;
  if (L.p.resident) grid = (grid / ntn) * ntn;
  if (grid <= 0) grid = L.p.resident ? ntn : 1;
  if (m_tiles <= 0) return;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * L.p.BN)) cols <<= 1;
  TcLaunch Lr = L;
  Lr.p.epi.M = M;
  const int key = (L.BK == 64 ? 0 : 4) | (L.epi_kind == EPI_CONV ? 0 : 2) | (L.fp16 ? 1 : 0);
  switch (key) {
    case 0: launch_one<64, EPI_CONV, false>(Lr, m_tiles, grid, cols, ablate, s); break;
    case 1: launch_one<64, EPI_CONV, true>(Lr, m_tiles, grid, cols, ablate, s); break;
    case 2: launch_one<64, EPI_LSTM, false>(Lr, m_tiles, grid, cols, ablate, s); break;
    case 3: launch_one<64, EPI_LSTM, true>(Lr, m_tiles, grid, cols, ablate, s); break;
    case 4: launch_one<32, EPI_CONV, false>(Lr, m_tiles, grid, cols, ablate, s); break;
    case 5: launch_one<32, EPI_CONV, true>(Lr, m_tiles, grid, cols, ablate, s); break;
    default: fail(DSB_ERR_INVALID, "LSTM epilogue needs BK=64");
  }
}

}  // namespace dsb
