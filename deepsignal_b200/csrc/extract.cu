// extract.cu — per-read feature math of the reference's _extract_features
// (deepsignal/extract_features.py:143-190,211-286; motif scan utils/process_utils.py:95-110) as
// HBM-bound CUDA kernels over packed read tables.
//
//  1. read_norm_kernel    per read: int16 histogram in shared memory -> exact median and MAD as weighted
//                         order statistics (np.median / statsmodels robust.mad semantics), or mean/std
//                         for zscore.  Normalisation is elementwise in the raw value, so downstream
//                         kernels recompute  around((scaling*(v+offset) - shift)/scale, 6)  from the int16
//                         sample in float64 with IEEE-rounded primitives (bit-identical to numpy).
//  2. lut_kernel          per read: float64 table of the normalised value of every distinct int16 sample value.
//  3. site_count / scan / site_emit   motif scan + ordered compaction of the targeted sites.
//  4. site_feature_kernel per site (one warp): lanes 0..k-1 compute the clipped slice length, mean and population std
//                         of their base in numpy's pairwise summation order (float64), then the warp writes the k-mer
//                         codes, genome position and the 360-sample window (extract_features.py:154-190).  (Round 1
//                         computed the base statistics in a separate per-base kernel over flagged bases: sparse warps,
//                         0.49 ms of the 1.17 ms chain; per site they are dense and need no flag / per-base arrays.)
#include <cmath>

#include "common.cuh"

namespace dsb {

struct ExtractWork {
  DevBuf counts, offsets, shift, scale, scratch, total, tmp_read, tmp_loc;
  DevBuf vmin, bins, lut_off, lut;
};
void destroy_extract(ExtractWork* p) { delete p; }

namespace {

constexpr double MAD_C = 0.6744897501960817;   // scipy.stats.norm.ppf(0.75), statsmodels robust.mad
constexpr int NORM_THREADS = 256;
constexpr int SMEM_BINS = 11264;               // 44 KB of uint32 counters; wider reads use a global scratch

__device__ __forceinline__ double rescale(double scaling, double offset, int v) {
  return __dmul_rn(scaling, __dadd_rn((double)v, offset));          // extract_features.py:211-212
}
__device__ __forceinline__ double normalise(double x, double shift, double scale) {
  const double q = __ddiv_rn(__dsub_rn(x, shift), scale);           // extract_features.py:150
  return __ddiv_rn(rint(__dmul_rn(q, 1e6)), 1e6);                   // np.around(., 6)
}

template <class T> __device__ __forceinline__ T block_reduce(T v, T* sh, bool is_max, bool is_min) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (other > v ? other : v) : is_min ? (other < v ? other : v) : v + other;
  }
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  T r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? (sh[i] > r ? sh[i] : r) : is_min ? (sh[i] < r ? sh[i] : r) : r + sh[i];
  return r;
}

struct NormParams {
  const int16_t* raw; const int64_t* raw_off; const double* scaling; const double* offset;
  int64_t R; int method; double* shift; double* scale; uint32_t* gscratch;
  int32_t* vmin; int64_t* bins;      // per-read raw minimum and number of distinct-value bins (LUT size)
};

__global__ void __launch_bounds__(NORM_THREADS) read_norm_kernel(const NormParams p) {
  __shared__ uint32_t hist_s[SMEM_BINS];
  __shared__ int red_i[NORM_THREADS / 32];
  __shared__ double red_d[NORM_THREADS / 32];
  __shared__ uint32_t part[NORM_THREADS];
  __shared__ int found[4];
  __shared__ double dres[2];
  const int tid = threadIdx.x;
  for (int64_t r = blockIdx.x; r < p.R; r += gridDim.x) {
    const int16_t* x = p.raw + p.raw_off[r];
    const int64_t N = p.raw_off[r + 1] - p.raw_off[r];
    const double sc = p.scaling[r], of = p.offset[r];
    if (N <= 0 || !(sc > 0.0)) {     // empty read / non-physical scaling: numpy would give nan
      if (tid == 0) { p.shift[r] = nan(""); p.scale[r] = nan(""); p.vmin[r] = 0; p.bins[r] = 0; }
      continue;
    }
    // ---- min / max (vectorised 16-byte loads over the aligned body) ----
    int lo = 32767, hi = -32768;
    const int64_t head = min(N, (int64_t)(((16 - ((uintptr_t)x & 15)) & 15) >> 1));
    const int64_t nvec = (N - head) / 8;
    const int64_t tail0 = head + nvec * 8;
    for (int64_t i = tid; i < head; i += NORM_THREADS) { const int v = x[i]; lo = min(lo, v); hi = max(hi, v); }
    const uint4* xv = reinterpret_cast<const uint4*>(x + head);
    for (int64_t i = tid; i < nvec; i += NORM_THREADS) {
      const uint4 q = __ldg(xv + i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = (int16_t)(w[j] & 0xffffu), b = (int16_t)(w[j] >> 16);
        lo = min(lo, min(a, b)); hi = max(hi, max(a, b));
      }
    }
    for (int64_t i = tail0 + tid; i < N; i += NORM_THREADS) { const int v = x[i]; lo = min(lo, v); hi = max(hi, v); }
    lo = block_reduce<int>(lo, red_i, false, true);
    hi = block_reduce<int>(hi, red_i, true, false);
    const int bins = hi - lo + 1;
    if (tid == 0) { p.vmin[r] = lo; p.bins[r] = bins; }
    uint32_t* hist = (bins <= SMEM_BINS) ? hist_s : p.gscratch + (size_t)blockIdx.x * 65536;
    for (int i = tid; i < bins; i += NORM_THREADS) hist[i] = 0;
    __syncthreads();
    // ---- histogram ----
    for (int64_t i = tid; i < head; i += NORM_THREADS) atomicAdd(&hist[x[i] - lo], 1u);
    for (int64_t i = tid; i < nvec; i += NORM_THREADS) {
      const uint4 q = __ldg(xv + i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&hist[(int)(int16_t)(w[j] & 0xffffu) - lo], 1u);
        atomicAdd(&hist[(int)(int16_t)(w[j] >> 16) - lo], 1u);
      }
    }
    for (int64_t i = tail0 + tid; i < N; i += NORM_THREADS) atomicAdd(&hist[x[i] - lo], 1u);
    __syncthreads();

    if (p.method == 1) {
      // ---- zscore: mean / population std from the histogram (float64) ----
      double s = 0.0;
      for (int i = tid; i < bins; i += NORM_THREADS) s += (double)hist[i] * rescale(sc, of, lo + i);
      const double mean = block_reduce<double>(s, red_d, false, false) / (double)N;
      double q = 0.0;
      for (int i = tid; i < bins; i += NORM_THREADS) { const double d = rescale(sc, of, lo + i) - mean; q += (double)hist[i] * d * d; }
      const double var = block_reduce<double>(q, red_d, false, false) / (double)N;
      if (tid == 0) { p.shift[r] = mean; p.scale[r] = sqrt(var); }
      __syncthreads();
      continue;
    }

    // ---- inclusive prefix sum of the histogram, in place ----
    const int per = (bins + NORM_THREADS - 1) / NORM_THREADS;
    const int b0 = min(bins, tid * per), b1 = min(bins, b0 + per);
    uint32_t local = 0;
    for (int i = b0; i < b1; ++i) local += hist[i];
    part[tid] = local;
    __syncthreads();
    if (tid < 32) {      // 256 partials: 8 per lane
      uint32_t v[8], run = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[j] = part[tid * 8 + j]; run += v[j]; }
      uint32_t incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
      uint32_t excl = incl - run;
#pragma unroll
      for (int j = 0; j < 8; ++j) { part[tid * 8 + j] = excl; excl += v[j]; }
    }
    __syncthreads();
    {
      uint32_t run = part[tid];
      for (int i = b0; i < b1; ++i) { run += hist[i]; hist[i] = run; }
    }
    __syncthreads();
    // P(i) = hist[i] (inclusive), P(-1) = 0
    auto P = [&](int i) -> uint32_t { return i < 0 ? 0u : hist[i]; };
    // ---- median: order statistics k1=(N-1)/2, k2=N/2 (np.median averages the two middle values) ----
    const uint32_t k1 = (uint32_t)((N - 1) / 2), k2 = (uint32_t)(N / 2);
    for (int i = tid; i < bins; i += NORM_THREADS) {
      const uint32_t pi = hist[i], pm = P(i - 1);
      if (pm <= k1 && k1 < pi) found[0] = i;
      if (pm <= k2 && k2 < pi) found[1] = i;
    }
    __syncthreads();
    const double med = __dmul_rn(__dadd_rn(rescale(sc, of, lo + found[0]), rescale(sc, of, lo + found[1])), 0.5);
    // ---- MAD: median of |x - med| / c, a V-shaped function of the bin index ----
    auto dev = [&](int i) -> double { return __ddiv_rn(fabs(__dsub_rn(rescale(sc, of, lo + i), med)), MAD_C); };
    for (int i = tid; i < bins; i += NORM_THREADS) {     // m0 = first bin with x >= med
      const bool c = rescale(sc, of, lo + i) >= med;
      const bool cp = (i > 0) && (rescale(sc, of, lo + i - 1) >= med);
      if (c && !cp) found[2] = i;
    }
    __syncthreads();
    const int m0 = found[2];
    const uint32_t PL = P(m0 - 1);                       // samples left of the median
    for (int i = tid; i < bins; i += NORM_THREADS) {
      const uint32_t cnt = hist[i] - P(i - 1);
      if (cnt == 0) continue;
      const double di = dev(i);
      uint32_t less, leq;
      if (i >= m0) {
        // left side: dev decreasing in j. first j in [0,m0) with dev(j) < di / <= di
        int a = 0, b = m0;
        while (a < b) { const int m = (a + b) >> 1; if (dev(m) < di) b = m; else a = m + 1; }
        const int jl = a;
        a = 0; b = m0;
        while (a < b) { const int m = (a + b) >> 1; if (dev(m) <= di) b = m; else a = m + 1; }
        const int jle = a;
        less = (P(i - 1) - PL) + (PL - P(jl - 1));
        leq = (hist[i] - PL) + (PL - P(jle - 1));
      } else {
        // right side: dev increasing in j. first j in [m0,bins) with dev(j) >= di / > di
        int a = m0, b = bins;
        while (a < b) { const int m = (a + b) >> 1; if (dev(m) >= di) b = m; else a = m + 1; }
        const int jg = a;
        a = m0; b = bins;
        while (a < b) { const int m = (a + b) >> 1; if (dev(m) > di) b = m; else a = m + 1; }
        const int jgt = a;
        less = (PL - hist[i]) + (P(jg - 1) - PL);
        leq = (PL - P(i - 1)) + (P(jgt - 1) - PL);
      }
      if (less <= k1 && k1 < leq) dres[0] = di;
      if (less <= k2 && k2 < leq) dres[1] = di;
    }
    __syncthreads();
    if (tid == 0) {
      p.shift[r] = med;
      p.scale[r] = __dmul_rn(__dadd_rn(dres[0], dres[1]), 0.5);
    }
    __syncthreads();
  }
}

// Normalisation is elementwise in the raw int16 value, so each read gets a float64 table
//   lut[v - vmin] = around((scaling*(v+offset) - shift)/scale, 6)
// (a few hundred to a few thousand entries) and every later kernel is a gather from it.
__global__ void lut_kernel(const int64_t* __restrict__ lut_off, const int32_t* __restrict__ vmin, const double* __restrict__ scaling,
                           const double* __restrict__ offset, const double* __restrict__ shift, const double* __restrict__ scale,
                           int64_t R, int64_t total_bins, double* __restrict__ lut) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_bins) return;
  int64_t a = 0, b = R;        // largest r with lut_off[r] <= i
  while (b - a > 1) { const int64_t m = (a + b) >> 1; if (lut_off[m] <= i) a = m; else b = m; }
  const int r = (int)a;
  const int v = vmin[r] + (int)(i - lut_off[r]);
  lut[i] = normalise(rescale(scaling[r], offset[r], v), shift[r], scale[r]);
}

// numpy's pairwise summation (DOUBLE_pairwise_sum): < 8 sequential, <= 128 eight-way unrolled, else halves.
template <class F> __device__ double pairwise_block(const F& f, int a0, int n) {
  // The additions happen in numpy's order; the (dependent raw -> LUT) loads of the sequential parts are issued together
  // first, so that their latencies overlap instead of adding up element by element.
  double res = 0.0;
  int i = 0;
  if (n >= 8) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(a0 + j);
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(a0 + i + j));
    }
    res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                    __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  }
  const int nt = n - i;                                  // fewer than 8 left (the whole slice when n < 8): sequential adds
  double t[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) t[j] = j < nt ? f(a0 + i + j) : 0.0;
#pragma unroll
  for (int j = 0; j < 7; ++j)
    if (j < nt) res = __dadd_rn(res, t[j]);
  return res;
}
// The n > 128 recursion of numpy (split at n/2 rounded down to a multiple of 8) with an explicit stack; `leaf(a0, n)`
// evaluates one block of at most 128 elements.
template <class Leaf> __device__ double pairwise_tree(int a0, int n, Leaf&& leaf) {
  if (n <= 128) return leaf(a0, n);
  struct Frame { int a0, n, state; double left; };
  Frame st[26];
  int sp = 0;
  st[0] = Frame{a0, n, 0, 0.0};
  double ret = 0.0;
  while (sp >= 0) {
    Frame& fr = st[sp];
    if (fr.n <= 128) { ret = leaf(fr.a0, fr.n); --sp; continue; }
    int n2 = fr.n / 2;
    n2 -= n2 % 8;
    if (fr.state == 0) { fr.state = 1; st[sp + 1] = Frame{fr.a0, n2, 0, 0.0}; ++sp; continue; }
    if (fr.state == 1) { fr.left = ret; fr.state = 2; st[sp + 1] = Frame{fr.a0 + n2, fr.n - n2, 0, 0.0}; ++sp; continue; }
    ret = __dadd_rn(fr.left, ret);
    --sp;
  }
  return ret;
}
// A slice of more than 128 samples, summed by a whole warp (all 32 lanes call it with the same arguments): lane 0 lists
// the leaf blocks of the split tree, the lanes sum one leaf each with the eight-accumulator scheme, lane 0 adds the leaf
// sums in the tree's order.  Every addition is the one the serial recursion performs, so the result is bit-identical -- but
// a 3,000-sample stall costs ~100 dependent additions per lane instead of 3,000 on one lane while 31 wait (and a 100,000-
// sample one no longer holds a warp for milliseconds).
constexpr int LEAF_CAP = 128;
struct LeafScratch { int a0[LEAF_CAP]; short len[LEAF_CAP]; double sum[LEAF_CAP]; };
template <class F> __device__ double pairwise_sum_warp(const F& f, int n, int lane, LeafScratch& sc) {
  int L = 0;
  if (lane == 0)
    pairwise_tree(0, n, [&](int a, int m) { if (L < LEAF_CAP) { sc.a0[L] = a; sc.len[L] = (short)m; } ++L; return 0.0; });
  L = __shfl_sync(0xffffffffu, L, 0);
  double r = 0.0;
  if (L > LEAF_CAP) {                                   // more than ~12 k samples in one base: serial on lane 0
    if (lane == 0) r = pairwise_tree(0, n, [&](int a, int m) { return pairwise_block(f, a, m); });
    return __shfl_sync(0xffffffffu, r, 0);
  }
  __syncwarp();
  for (int k = lane; k < L; k += 32) sc.sum[k] = pairwise_block(f, sc.a0[k], sc.len[k]);
  __syncwarp();
  if (lane == 0) { int k = 0; r = pairwise_tree(0, n, [&](int, int) { return sc.sum[k++]; }); }
  __syncwarp();
  return __shfl_sync(0xffffffffu, r, 0);
}

struct NormAt {
  const int16_t* x; const double* lut;       // lut already offset by -vmin
  __device__ double operator()(int i) const { return __ldg(lut + x[i]); }
};
struct SqDevAt {
  NormAt f; double mean;
  __device__ double operator()(int i) const { const double d = __dsub_rn(f(i), mean); return __dmul_rn(d, d); }
};

struct MotifSet { int n, len, mod_loc, nbh; char seq[512]; };

__device__ __forceinline__ bool motif_hit(const uint8_t* b, const MotifSet& m) {
  for (int k = 0; k < m.n; ++k) {
    bool ok = true;
    for (int j = 0; j < m.len; ++j) ok = ok && (b[j] == (uint8_t)m.seq[k * m.len + j]);
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ bool site_at(const uint8_t* bases, int nb, int i, const MotifSet& m) {
  if (i + m.len > nb) return false;
  const int loc = i + m.mod_loc;
  if (loc < m.nbh || loc >= nb - m.nbh) return false;               // extract_features.py:254
  return motif_hit(bases + i, m);
}

// A read whose normalisation is undefined (no raw samples, non-positive scaling: read_norm_kernel leaves it without a
// LUT) contributes no sites, like a read the reference fails on and drops (extract_features.py:281-283).
__device__ __forceinline__ bool read_usable(const int64_t* raw_off, const double* scaling, int64_t r) {
  return raw_off[r + 1] > raw_off[r] && scaling[r] > 0.0;
}

__global__ void __launch_bounds__(128) site_count_kernel(const uint8_t* ev_base, const int64_t* ev_off, int64_t R,
                                                         const int64_t* raw_off, const double* scaling,
                                                         const MotifSet m, int64_t* counts) {
  __shared__ int sh[4];
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    const uint8_t* b = ev_base + ev_off[r];
    const int nb = read_usable(raw_off, scaling, r) ? (int)(ev_off[r + 1] - ev_off[r]) : 0;
    int c = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) c += site_at(b, nb, i, m) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) counts[r] = (int64_t)sh[0] + sh[1] + sh[2] + sh[3];
    __syncthreads();
  }
}

// exclusive scan of counts[R] -> offsets[R+1]; single CTA, 1024 threads
__global__ void __launch_bounds__(1024) scan_kernel(const int64_t* counts, int64_t* offsets, int64_t R, int64_t* total) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t base = 0; base < R; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < R ? counts[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int64_t s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
      wsum[lane] = s;
    }
    __syncthreads();
    const int64_t before = carry + (w > 0 ? wsum[w - 1] : 0) + incl - v;
    if (i < R) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) { offsets[R] = carry; *total = carry; }
}

__global__ void __launch_bounds__(128) site_emit_kernel(const uint8_t* ev_base, const int64_t* ev_off, int64_t R,
                                                        const int64_t* raw_off, const double* scaling, const MotifSet m,
                                                        const int64_t* offsets, int64_t capacity, int32_t* site_read, int32_t* site_loc) {
  __shared__ int wcnt[4];
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    const uint8_t* b = ev_base + ev_off[r];
    const int nb = read_usable(raw_off, scaling, r) ? (int)(ev_off[r + 1] - ev_off[r]) : 0;
    int64_t run = offsets[r];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += blockDim.x) {
      const int i = base + threadIdx.x;
      const bool hit = i < nb && site_at(b, nb, i, m);
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) wcnt[w] = __popc(bal);
      __syncthreads();
      int before = __popc(bal & ((1u << lane) - 1));
      for (int k = 0; k < w; ++k) before += wcnt[k];
      const int tot = wcnt[0] + wcnt[1] + wcnt[2] + wcnt[3];
      if (hit) {
        const int64_t o = run + before;
        if (o < capacity) {
          if (site_read) site_read[o] = (int32_t)r;
          site_loc[o] = i + m.mod_loc;
        }
      }
      run += tot;
      __syncthreads();
    }
  }
}

struct SiteParams {
  dsb_reads rd;
  const double* lut; const int64_t* lut_off; const int32_t* vmin;
  const int32_t* site_read_in; const int32_t* site_loc_in;
  dsb_sites out;
  int64_t n_sites;
  int klen, slen;
};

__device__ __forceinline__ int base_code(uint8_t c) {
  switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': case 'U': return 3; case 'N': return 4; default: return 255; }
}

constexpr int SITE_WARPS = 8;
constexpr int MAX_KLEN = 33;
constexpr int LONG_BASE = 128;       // slices longer than this are summed by the whole warp (pairwise_sum_warp)

__global__ void __launch_bounds__(SITE_WARPS * 32) site_feature_kernel(const SiteParams p) {
  __shared__ int cum_s[SITE_WARPS][MAX_KLEN + 1];
  __shared__ int st_s[SITE_WARPS][MAX_KLEN];
  __shared__ LeafScratch leaf_s[SITE_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t site = (int64_t)blockIdx.x * SITE_WARPS + w;
  if (site >= p.n_sites) return;
  const int K = p.klen, S = p.slen, nbh = (K - 1) / 2;
  const int r = p.site_read_in[site], loc = p.site_loc_in[site];
  const int64_t e0 = p.rd.ev_off[r];
  const int nb = (int)(p.rd.ev_off[r + 1] - e0);
  const int64_t N = p.rd.raw_off[r + 1] - p.rd.raw_off[r];
  const int16_t* x = p.rd.raw + p.rd.raw_off[r];
  int* cum = cum_s[w];
  int* st = st_s[w];
  // per-base gathers; K <= 33 -> up to two bases per lane
  const double* lut = p.lut + p.lut_off[r] - p.vmin[r];
  auto put_stats = [&](int j, double mu, double sd) {
    if (p.out.means) p.out.means[site * K + j] = (float)mu;
    if (p.out.stds) p.out.stds[site * K + j] = (float)sd;
    if (p.out.means64) p.out.means64[site * K + j] = mu;
    if (p.out.stds64) p.out.stds64[site * K + j] = sd;
  };
  for (int j = lane; j < K; j += 32) {
    const int64_t e = e0 + loc - nbh + j;
    // python slice norm[s : s+l] for s >= 0 (negative starts are rejected on the host side), then np.mean / np.std of it
    // in numpy's pairwise summation order (the sums are bit-identical to the reference's)
    const int64_t s_raw = p.rd.ev_start[e], l_raw = p.rd.ev_len[e];
    const int64_t s = s_raw < 0 ? 0 : (s_raw > N ? N : s_raw);
    int64_t en = s_raw + l_raw;
    en = en < s ? s : (en > N ? N : en);
    const int len = (int)(en - s);
    st[j] = (int)s;
    cum[j + 1] = len;
    const int code = base_code(p.rd.ev_base[e]);
    if (p.out.kmer) p.out.kmer[site * K + j] = code;
    if (p.out.lens) p.out.lens[site * K + j] = (float)len;
    if (p.out.lens_i32) p.out.lens_i32[site * K + j] = len;
    if (len > LONG_BASE) continue;                          // stalls: summed by the whole warp below
    double mu = nan(""), sd = nan("");                      // np.mean([]) -> nan
    if (len > 0) {
      NormAt f{x + s, lut};
      mu = __ddiv_rn(pairwise_block(f, 0, len), (double)len);
      SqDevAt g{f, mu};
      sd = sqrt(__ddiv_rn(pairwise_block(g, 0, len), (double)len));
    }
    put_stats(j, mu, sd);
  }
  __syncwarp();
  for (int j = 0; j < K; ++j) {                             // warp-uniform: st[] / cum[] are shared
    const int len = cum[j + 1];                             // (still the slice length: the prefix sum follows)
    if (len <= LONG_BASE) continue;
    NormAt f{x + st[j], lut};
    const double mu = __ddiv_rn(pairwise_sum_warp(f, len, lane, leaf_s[w]), (double)len);
    SqDevAt g{f, mu};
    const double sd = sqrt(__ddiv_rn(pairwise_sum_warp(g, len, lane, leaf_s[w]), (double)len));
    if (lane == 0) put_stats(j, mu, sd);
  }
  __syncwarp();
  if (lane == 0) {
    cum[0] = 0;
    for (int j = 0; j < K; ++j) cum[j + 1] += cum[j];
    // genome coordinates (extract_features.py:255-260)
    const int64_t cs = p.rd.chrom_start[r], cl = p.rd.chrom_len ? p.rd.chrom_len[r] : -1;
    int64_t pos, pis;
    if (p.rd.strand_minus[r]) { pos = cs + nb - 1 - loc; pis = cl >= 0 ? cl - 1 - pos : -1; }
    else { pos = cs + loc; pis = cl >= 0 ? pos : -1; }
    if (p.out.pos) p.out.pos[site] = pos;
    if (p.out.pos_in_strand) p.out.pos_in_strand[site] = pis;
    if (p.out.site_read) p.out.site_read[site] = r;
    if (p.out.site_loc) p.out.site_loc[site] = loc;
  }
  __syncwarp();
  // ---- window rule (extract_features.py:154-190) ----
  const int T = cum[K];
  const int mid = (K - 1) / 2;
  const int midlen = cum[mid + 1] - cum[mid];
  int mode, start, nreal;
  if (T < S) { mode = 0; start = 0; nreal = T; }
  else if (midlen >= S) { mode = 1; start = cum[mid]; nreal = S; }
  else {
    int left = (S - midlen) / 2, right = S - left;
    const int Ls = cum[mid], Rs = T - cum[mid];
    if (left > Ls) { right += left - Ls; left = Ls; }
    else if (right > Rs) { left += right - Rs; right = Rs; }
    mode = 2; start = Ls - left; nreal = S;
  }
  if (lane == 0 && p.out.window) {
    p.out.window[site * 3] = mode; p.out.window[site * 3 + 1] = start; p.out.window[site * 3 + 2] = nreal;
  }
  float* sig32 = p.out.signals ? p.out.signals + site * S : nullptr;
  double* sig64 = p.out.signals64 ? p.out.signals64 + site * S : nullptr;
  int b = mode == 1 ? mid : 0;                              // q is non-decreasing in j: the base pointer only moves forward
  for (int j = lane; j < S; j += 32) {
    double v = 0.0;
    if (j < nreal) {
      // concatenated index q -> (base b, offset); mode 1 resamples the middle base (64-bit product)
      int q = start + j;
      if (mode == 1) q = start + (int)(((int64_t)j * midlen) / S);
      else while (b + 1 < K && cum[b + 1] <= q) ++b;
      v = __ldg(lut + x[st[b] + (q - cum[b])]);
    }
    if (sig32) sig32[j] = (float)v;
    if (sig64) sig64[j] = v;
  }
}

}  // namespace

void extract_dev(dsb_ctx* ctx, const dsb_reads* rd, const dsb_extract_params* prm, const dsb_sites* sites,
                 int64_t* n_sites, double* read_shift, double* read_scale, cudaStream_t s) {
  if (!rd || !prm || !sites || !n_sites) fail(DSB_ERR_INVALID, "null argument");
  if (prm->kmer_len < 1 || prm->kmer_len > MAX_KLEN || prm->kmer_len % 2 == 0) fail(DSB_ERR_INVALID, "kmer_len must be odd");
  if (prm->signal_len < 1) fail(DSB_ERR_INVALID, "signal_len must be positive");
  if (prm->normalize_method != 0 && prm->normalize_method != 1) fail(DSB_ERR_INVALID, "normalize_method must be 0 (mad) or 1 (zscore)");
  if (prm->n_motifs < 1 || prm->motif_len < 1 || prm->n_motifs * prm->motif_len > 512 || !prm->motifs) fail(DSB_ERR_INVALID, "bad motif set");
  if (prm->mod_loc < 0 || prm->mod_loc >= prm->motif_len) fail(DSB_ERR_INVALID, "mod_loc outside the motif");
  *n_sites = 0;
  const int64_t R = rd->n_reads;
  if (R <= 0) return;
  DSB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->xw) ctx->xw.reset(new ExtractWork());
  ExtractWork& w = *ctx->xw;
  // (raw_off[0] == ev_off[0] == 0 is part of the contract; dsb_extract_host checks it on its host arrays)
  const int grid_norm = (int)std::min<int64_t>(R, (int64_t)ctx->sm_count * 4);
  w.shift.ensure(R * sizeof(double));
  w.scale.ensure(R * sizeof(double));
  w.scratch.ensure((size_t)grid_norm * 65536 * sizeof(uint32_t));
  w.counts.ensure(R * sizeof(int64_t));
  w.offsets.ensure((R + 1) * sizeof(int64_t));

  w.vmin.ensure(R * sizeof(int32_t));
  w.bins.ensure(R * sizeof(int64_t));
  w.lut_off.ensure((R + 1) * sizeof(int64_t));
  NormParams np{rd->raw, rd->raw_off, rd->scaling, rd->offset, R, prm->normalize_method,
                w.shift.as<double>(), w.scale.as<double>(), w.scratch.as<uint32_t>(), w.vmin.as<int32_t>(), w.bins.as<int64_t>()};
  {
    Timer tm(ctx, s, KC_X_NORM, 0);
    read_norm_kernel<<<grid_norm, NORM_THREADS, 0, s>>>(np);
  }
  DSB_CUDA(cudaGetLastError());
  MotifSet ms;
  memset(&ms, 0, sizeof ms);
  ms.n = prm->n_motifs; ms.len = prm->motif_len; ms.mod_loc = prm->mod_loc; ms.nbh = (prm->kmer_len - 1) / 2;
  memcpy(ms.seq, prm->motifs, (size_t)ms.n * ms.len);
  const int grid_r = (int)std::min<int64_t>(R, (int64_t)ctx->sm_count * 16);
  w.total.ensure(2 * sizeof(int64_t));
  {
    Timer tm(ctx, s, KC_X_SITES, 0);
    site_count_kernel<<<grid_r, 128, 0, s>>>(rd->ev_base, rd->ev_off, R, rd->raw_off, rd->scaling, ms, w.counts.as<int64_t>());
    DSB_CUDA(cudaGetLastError());
    scan_kernel<<<1, 1024, 0, s>>>(w.counts.as<int64_t>(), w.offsets.as<int64_t>(), R, w.total.as<int64_t>());
    scan_kernel<<<1, 1024, 0, s>>>(w.bins.as<int64_t>(), w.lut_off.as<int64_t>(), R, w.total.as<int64_t>() + 1);
  }
  DSB_CUDA(cudaGetLastError());
  ctx->launches += 4;
  int64_t totals[2] = {0, 0};
  DSB_CUDA(cudaMemcpyAsync(totals, w.total.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  DSB_CUDA(cudaStreamSynchronize(s));
  const int64_t total = totals[0], total_bins = totals[1];
  *n_sites = total;
  w.lut.ensure(std::max<int64_t>(total_bins, 1) * sizeof(double));
  if (total_bins > 0) {
    {
      Timer tm(ctx, s, KC_X_LUT, 0);
      lut_kernel<<<(unsigned)((total_bins + 255) / 256), 256, 0, s>>>(w.lut_off.as<int64_t>(), w.vmin.as<int32_t>(), rd->scaling, rd->offset,
                                                                      w.shift.as<double>(), w.scale.as<double>(), R, total_bins, w.lut.as<double>());
    }
    DSB_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  const int64_t emit = std::min<int64_t>(total, sites->capacity);
  if (read_shift) DSB_CUDA(cudaMemcpyAsync(read_shift, w.shift.p, R * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (read_scale) DSB_CUDA(cudaMemcpyAsync(read_scale, w.scale.p, R * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (emit > 0) {
    // site_read / site_loc are needed by the feature kernel even if the caller does not want them
    int32_t* sr = sites->site_read;
    int32_t* sl = sites->site_loc;
    if (!sr || !sl) {
      w.tmp_read.ensure(emit * sizeof(int32_t));
      w.tmp_loc.ensure(emit * sizeof(int32_t));
      if (!sr) sr = w.tmp_read.as<int32_t>();
      if (!sl) sl = w.tmp_loc.as<int32_t>();
    }
    {
      Timer tm(ctx, s, KC_X_SITES, 0);
      site_emit_kernel<<<grid_r, 128, 0, s>>>(rd->ev_base, rd->ev_off, R, rd->raw_off, rd->scaling, ms, w.offsets.as<int64_t>(), emit, sr, sl);
    }
    DSB_CUDA(cudaGetLastError());
    SiteParams sp;
    sp.rd = *rd;
    sp.lut = w.lut.as<double>(); sp.lut_off = w.lut_off.as<int64_t>(); sp.vmin = w.vmin.as<int32_t>();
    sp.site_read_in = sr; sp.site_loc_in = sl;
    sp.out = *sites;
    sp.n_sites = emit;
    sp.klen = prm->kmer_len; sp.slen = prm->signal_len;
    {
      Timer tm(ctx, s, KC_X_FEAT, 0);
      site_feature_kernel<<<(unsigned)((emit + SITE_WARPS - 1) / SITE_WARPS), SITE_WARPS * 32, 0, s>>>(sp);
    }
    DSB_CUDA(cudaGetLastError());
    ctx->launches += 2;      // site_emit + site_feature
  }
  if (total > sites->capacity) fail(DSB_ERR_CAPACITY, "%lld sites found but the output capacity is %lld", (long long)total, (long long)sites->capacity);
}

}  // namespace dsb
