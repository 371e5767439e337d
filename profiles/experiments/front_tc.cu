// front_tc.cu — fused CNN front end: conv1 (1x7 s2, fp32 CUDA cores) + maxpool (1x3 s2) -> conv2 (1x1) -> conv3 (1x3),
// see front_tc.cuh.  Persistent, warp-specialised, 704 threads:
//   warp 0        TMA: the CTA's weights (conv2 whole, its 128-channel half of conv3), once
//   warp 1        MMA issuer: conv2 [128x64]·[64x128] from the P1 tile, conv3 as three row-shifted taps on the staged C2 tile
//   warps 2..9    stem: 128 rows x 64 channels of relu(maxpool(relu(conv1))) per tile -> 16-bit P1 tile (SWIZZLE_128B)
//   warps 10..13  C2 staging: TMEM -> +bias, ReLU, halo rows = 0 -> 16-bit C2 tile (four [136 x 32] SWIZZLE_64B k-blocks,
//                 stored one row down so that tap t of output row i is tile row i + t)
//   warps 14..21  epilogue: TMEM -> +bias, ReLU, halo rows = 0 -> X0[:, nh*128 .. nh*128+128)
// A CTA owns one half of conv3's 256 output channels (96 KB of weights; the whole matrix does not fit beside the tiles),
// so each 128-row tile is visited by two CTAs; the C2 tile and both accumulators are double buffered across tiles.
// Arithmetic and rounding points are those of stem_kernel + two gemm_tc launches: results are bit-identical.
#include <cstdlib>

#include "front_tc.cuh"
#include "tc_common.cuh"

namespace dsb {
namespace {

using namespace tc;

constexpr int F_THREADS = 704;
constexpr uint32_t F_W3_BLK = 128 * 32 * 2;            // one [128 x 32] SWIZZLE_64B weight block (8 KB); 12 = 3 taps x 4 k-blocks
constexpr uint32_t F_W2_BYTES = 128 * 64 * 2;          // conv2 weights [128 x 64] SWIZZLE_128B
constexpr uint32_t F_P1_BYTES = 128 * 64 * 2;          // P1 tile [128 x 64] SWIZZLE_128B
constexpr int F_C2_ROWS = 136;
constexpr uint32_t F_C2_KB = F_C2_ROWS * 32 * 2;       // one [136 x 32] SWIZZLE_64B k-block of the C2 tile (8704 B)

struct FSmem {
  static constexpr uint32_t w3 = 0;
  static constexpr uint32_t w2 = w3 + 12 * F_W3_BLK;                 // 98304
  static constexpr uint32_t p1 = w2 + F_W2_BYTES;                    // one buffer: conv2 frees it a few hundred cycles after it is full
  static constexpr uint32_t c2 = p1 + F_P1_BYTES;                    // 2 buffers x 4 k-blocks
  static constexpr uint32_t stage = c2 + 8 * F_C2_KB;                // 8 epilogue warps x 2 KB
  static constexpr uint32_t bars = stage + 8 * EPI_STAGE_BYTES;
  // barriers: wbar, then for buf in {0,1}: p1_full p1_empty acc2_full acc2_empty c2_full c2_empty acc3_full acc3_empty
  static constexpr uint32_t nbars = 1 + 16;
  static constexpr uint32_t tmem_slot = (bars + 8 * nbars + 15u) & ~15u;
  static constexpr uint32_t b2 = tmem_slot + 16;
  static constexpr uint32_t b3 = b2 + 128 * 4;
  static constexpr uint32_t win = b3 + 128 * 4;                      // [128][12] fp32 signal windows of the tile
  static constexpr uint32_t meta = win + 128 * 12 * 4;               // [128] int
  static constexpr uint32_t total = meta + 128 * 4;
};

__device__ __forceinline__ uint32_t idesc16(bool f16, int n) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <bool F16>
__global__ void __launch_bounds__(F_THREADS, 1)
front_kernel(const __grid_constant__ FrontMaps maps, const FrontParams p, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbar = base + FSmem::bars;
  auto bar = [&](int kind, int buf) { return base + FSmem::bars + 8u * (1 + buf * 8 + kind); };
  enum { P1_FULL = 0, P1_EMPTY, ACC2_FULL, ACC2_EMPTY, C2_FULL, C2_EMPTY, ACC3_FULL, ACC3_EMPTY };
  const uint32_t tmem_slot = base + FSmem::tmem_slot;
  float* sb2 = reinterpret_cast<float*>(gbase + FSmem::b2);
  float* sb3 = reinterpret_cast<float*>(gbase + FSmem::b3);

  const int nh = (int)(blockIdx.x & 1);                  // which 128-channel half of conv3 this CTA owns
  const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);
  const int my_tiles = t_first < num_tiles ? (num_tiles - t_first + t_step - 1) / t_step : 0;

  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(P1_FULL, b), 8);   mbar_init(bar(P1_EMPTY, b), 1);
      mbar_init(bar(ACC2_FULL, b), 1); mbar_init(bar(ACC2_EMPTY, b), 4);
      mbar_init(bar(C2_FULL, b), 4);   mbar_init(bar(C2_EMPTY, b), 1);
      mbar_init(bar(ACC3_FULL, b), 1); mbar_init(bar(ACC3_EMPTY, b), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128; i += F_THREADS) sb2[i] = p.b2[i];
  for (int i = threadIdx.x; i < 128; i += F_THREADS) sb3[i] = p.b3[nh * 128 + i];
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
    // ---------------- weights, once ----------------
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(wbar, F_W2_BYTES + 12 * F_W3_BLK);
      tma_load_2d(base + FSmem::w2, &maps.w2, wbar, 0, 0);
      for (int j = 0; j < 12; ++j) tma_load_2d(base + FSmem::w3 + j * F_W3_BLK, &maps.w3, wbar, j * 32, nh * 128);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && my_tiles > 0) {
      const uint32_t id128 = idesc16(F16, 128);
      mbar_wait(wbar, 0);
      auto conv2 = [&](int kk) {
        const int b = kk & 1;
        const uint32_t ph = (uint32_t)(kk >> 1) & 1u;
        mbar_wait(bar(ACC2_EMPTY, b), ph ^ 1u);
        mbar_wait(bar(P1_FULL, 0), (uint32_t)kk & 1u);
        tc_fence_after();
        const uint64_t ad = make_kmajor_desc<64>(base + FSmem::p1);
        const uint64_t bd = make_kmajor_desc<64>(base + FSmem::w2);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base + b * 128, ad + 2 * k, bd + 2 * k, id128, k != 0 ? 1u : 0u);
        tc_commit(bar(P1_EMPTY, 0));
        tc_commit(bar(ACC2_FULL, b));
      };
      auto conv3 = [&](int kk) {
        const int b = kk & 1;
        const uint32_t ph = (uint32_t)(kk >> 1) & 1u;
        mbar_wait(bar(ACC3_EMPTY, b), ph ^ 1u);
        mbar_wait(bar(C2_FULL, b), ph);
        tc_fence_after();
        const uint32_t d = tmem_base + 256 + b * 128;
        for (int tap = 0; tap < 3; ++tap)
          for (int kb = 0; kb < 4; ++kb) {
            // output row i, tap t reads C2 row i + t - 1 = tile row i + t (the staging warps store one row down)
            const uint64_t ad = make_kmajor_desc<32>(base + FSmem::c2 + (b * 4 + kb) * F_C2_KB + (uint32_t)tap * 64u);
            const uint64_t bd = make_kmajor_desc<32>(base + FSmem::w3 + (tap * 4 + kb) * F_W3_BLK);
#pragma unroll
            for (int k = 0; k < 2; ++k) tc_mma_bf16(d, ad + 2 * k, bd + 2 * k, id128, (tap | kb | k) != 0 ? 1u : 0u);
          }
        tc_commit(bar(C2_EMPTY, b));
        tc_commit(bar(ACC3_FULL, b));
      };
      // conv2 of tile k+1 is issued before conv3 of tile k: the TMEM -> smem staging of C2 is off the critical path
      conv2(0);
      for (int kk = 0; kk < my_tiles; ++kk) {
        if (kk + 1 < my_tiles) conv2(kk + 1);
        conv3(kk);
      }
    }
  } else if (warp < 10) {
    // ---------------- stem warps: conv1 + ReLU + maxpool for 128 rows x 64 channels ----------------
    // A warp owns a strip of 16 tile rows, a lane owns two channels with their 2 x 7 weights in registers.  The 11-sample
    // signal window of a row is staged in shared memory once per tile and read as broadcast float4s; consecutive rows of a
    // sample slide the window by four samples and re-use the previous row's last conv value, as stem_kernel does.
    const int st = threadIdx.x - 64;             // 0..255
    const int ws = warp - 2;                     // strip: tile rows ws*16 .. ws*16+15
    const int c0 = 2 * lane;                     // channels c0, c0+1
    float wa[7], wb[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) { wa[t] = p.w1[c0 * 7 + t]; wb[t] = p.w1[(c0 + 1) * 7 + t]; }
    const float ba = p.b1[c0], bb = p.b1[c0 + 1];
    float* win = reinterpret_cast<float*>(gbase + FSmem::win);       // [128][12]: x[k] at [row][k+1]
    int* meta = reinterpret_cast<int*>(gbase + FSmem::meta);         // [128]: bit 4 row valid, bit 3 first row of a sample, bits 0-2 conv cell valid
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      named_bar_sync(3, 256);                    // every stem thread is done with the previous tile's windows
      {
        const int i = st >> 1, hf = st & 1;
        const int r = tile * FRONT_TILE_ROWS - 1 + i;
        bool valid = r >= 0 && r < p.M;
        const int s = valid ? r / p.Lp : 0;
        const int pos = r - s * p.Lp;
        valid = valid && pos < p.L;
        const int q0 = 2 * pos - p.pool_pl;      // first of the three conv positions under this pooled row
        const float* sig = p.signals + (size_t)s * p.S;
        const int j0 = 2 * q0 - p.conv_pl;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int kx = hf * 6 + k;             // hf 0: x[0..5], hf 1: x[6..10] (+ one pad slot)
          const int j = j0 + kx;
          if (kx < 11) win[i * 12 + kx + 1] = (valid && j >= 0 && j < p.S) ? __ldg(sig + j) : 0.f;
        }
        if (hf == 0) {
          int m = 0;
          if (valid) {
            m = 16 | (pos == 0 ? 8 : 0);
#pragma unroll
            for (int a = 0; a < 3; ++a) m |= (q0 + a >= 0 && q0 + a < p.L1) ? (1 << a) : 0;
          }
          meta[i] = m;
        }
      }
      named_bar_sync(3, 256);
      mbar_wait(bar(P1_EMPTY, 0), ((uint32_t)kk & 1u) ^ 1u);
      float x[11];
      float pa = 0.f, pb = 0.f;                  // conv value of the previous row's last cell (= this row's first)
      bool have_prev = false;
#pragma unroll 1
      for (int rr = 0; rr < 16; ++rr) {
        const int i = ws * 16 + rr;
        const int m = meta[i];
        float oa = 0.f, ob = 0.f;
        if (m & 16) {
          const float4* wr = reinterpret_cast<const float4*>(win + i * 12);
          const bool restart = !have_prev || (m & 8);
          if (restart) {
            const float4 v0 = wr[0], v1 = wr[1];
            x[0] = v0.y; x[1] = v0.z; x[2] = v0.w; x[3] = v1.x; x[4] = v1.y; x[5] = v1.z; x[6] = v1.w;
          } else {
#pragma unroll
            for (int k = 0; k < 7; ++k) x[k] = x[k + 4];
          }
          const float4 v2 = wr[2];
          x[7] = v2.x; x[8] = v2.y; x[9] = v2.z; x[10] = v2.w;
          auto conv = [&](int a, float& ra, float& rb) {
            float fa = ba, fb = bb;
#pragma unroll
            for (int t = 0; t < 7; ++t) { fa = fmaf(wa[t], x[2 * a + t], fa); fb = fmaf(wb[t], x[2 * a + t], fb); }
            const bool ok = (m >> a) & 1;
            ra = ok ? fmaxf(fa, 0.f) : 0.f;      // ReLU; invalid conv cells contribute 0 (<= every valid value)
            rb = ok ? fmaxf(fb, 0.f) : 0.f;
          };
          float a0 = pa, b0 = pb, a1, b1, a2, b2;
          if (restart) conv(0, a0, b0);
          conv(1, a1, b1);
          conv(2, a2, b2);
          oa = fmaxf(a0, fmaxf(a1, a2));
          ob = fmaxf(b0, fmaxf(b1, b2));
          pa = a2; pb = b2;
          have_prev = true;
        } else {
          have_prev = false;
        }
        // channels c0, c0+1 = 4 bytes at byte 4*lane of the 128-byte row; SWIZZLE_128B: 16-byte piece j sits at j ^ (row & 7)
        const uint32_t addr = base + FSmem::p1 + (uint32_t)i * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)i & 7u)) << 4) + (((uint32_t)lane & 3u) << 2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack16<F16>(oa, ob)) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P1_FULL, 0));
    }
  } else if (warp < 14) {
    // ---------------- C2 staging warps: relu(conv2 + bias) -> 16-bit A operand of conv3 ----------------
    const int q = warp & 3;
    const int i = q * 32 + lane;
    const uint32_t rr = (uint32_t)(i + 1);                   // tile row (one row down)
    const uint32_t sw = (rr >> 1) & 3u;                      // SWIZZLE_64B: piece c sits at c ^ ((row >> 1) & 3)
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      const int b = kk & 1;
      const uint32_t ph = (uint32_t)(kk >> 1) & 1u;
      const int row = tile * FRONT_TILE_ROWS - 1 + i;
      const bool zero = row < 0 || row >= p.M || (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128);
      mbar_wait(bar(ACC2_FULL, b), ph);
      mbar_wait(bar(C2_EMPTY, b), ph ^ 1u);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        uint32_t r2[2][16];
        tc_ld16(t_row + (uint32_t)(kb * 32), r2[0]);
        tc_ld16(t_row + (uint32_t)(kb * 32 + 16), r2[1]);
        tc_wait_ld();
        const uint32_t ub = base + FSmem::c2 + (b * 4 + kb) * F_C2_KB + rr * 64u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = *reinterpret_cast<const float4*>(sb2 + kb * 32 + c * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sb2 + kb * 32 + c * 8 + 4);
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t pk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = c * 8 + 2 * k;
            float a = fmaxf(__uint_as_float(r2[e >> 4][e & 15]) + bv[2 * k], 0.f);
            float d = fmaxf(__uint_as_float(r2[(e + 1) >> 4][(e + 1) & 15]) + bv[2 * k + 1], 0.f);
            if (zero) { a = 0.f; d = 0.f; }
            pk[k] = pack16<F16>(a, d);
          }
          sts128(ub + (((uint32_t)c ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(C2_FULL, b)); mbar_arrive(bar(ACC2_EMPTY, b)); }
    }
  } else {
    // ---------------- epilogue warps: relu(conv3 + bias) -> X0 ----------------
    const int q = warp & 3;
    const int half = (warp - 14) >> 2;
    const uint32_t stage = base + FSmem::stage + (uint32_t)(warp - 14) * EPI_STAGE_BYTES;
    for (int kk = 0; kk < my_tiles; ++kk) {
      const int tile = t_first + kk * t_step;
      const int b = kk & 1;
      const uint32_t ph = (uint32_t)(kk >> 1) & 1u;
      const int i = q * 32 + lane;
      const int row = tile * FRONT_TILE_ROWS - 1 + i;
      const int wrow0 = tile * FRONT_TILE_ROWS - 1 + q * 32;
      const bool row_ok = i >= 1 && i <= FRONT_TILE_ROWS && row < p.M;
      const uint32_t ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const bool halo = row < 0 || (row % p.Lp) >= p.L;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(b * 128);
      mbar_wait(bar(ACC3_FULL, b), ph);
      tc_fence_after();
      for (int g = half; g < 4; g += 2) {
        uint32_t r2[2][16];
        __syncwarp();
        tc_ld16(t_row + (uint32_t)(g * 32), r2[0]);
        tc_ld16(t_row + (uint32_t)(g * 32 + 16), r2[1]);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int col = g * 32 + 16 * h2;
          float v[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 bb = *reinterpret_cast<const float4*>(sb3 + col + 4 * k);
            v[4 * k + 0] = __uint_as_float(r2[h2][4 * k + 0]) + bb.x;
            v[4 * k + 1] = __uint_as_float(r2[h2][4 * k + 1]) + bb.y;
            v[4 * k + 2] = __uint_as_float(r2[h2][4 * k + 2]) + bb.z;
            v[4 * k + 3] = __uint_as_float(r2[h2][4 * k + 3]) + bb.w;
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = halo ? 0.f : fmaxf(v[k], 0.f);
#pragma unroll
          for (int k = 0; k < 8; ++k) pk[8 * h2 + k] = pack16<F16>(v[2 * k], v[2 * k + 1]);
        }
        const size_t pitch = (size_t)p.ldX * 2;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.X0) + (ptrdiff_t)wrow0 * (ptrdiff_t)pitch + (size_t)(nh * 128 + g * 32) * 2;
        staged_store<4>(stage, pk, lane, ok_mask, dst, pitch);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(ACC3_EMPTY, b));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

void launch_front(const FrontLaunch& L, const float* signals, int n, int sm_count, cudaStream_t s) {
  if (n <= 0) return;
  FrontParams p = L.p;
  p.signals = signals;
  p.M = n * p.Lp;
  const int num_tiles = (p.M + FRONT_TILE_ROWS - 1) / FRONT_TILE_ROWS;
  int grid = sm_count & ~1;
  if (grid > 2 * num_tiles) grid = 2 * num_tiles;
  if (grid < 2) grid = 2;
  const size_t smem = FSmem::total + 1024;
  static bool cfg[2][64] = {{false}};
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  auto kern = L.fp16 ? front_kernel<true> : front_kernel<false>;
  if (!cfg[L.fp16 ? 1 : 0][dev & 63]) {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg[L.fp16 ? 1 : 0][dev & 63] = true;
  }
  launch_pdl(kern, grid, F_THREADS, smem, s, L.maps, p, num_tiles);
  DSB_CUDA(cudaGetLastError());
}

}  // namespace dsb
