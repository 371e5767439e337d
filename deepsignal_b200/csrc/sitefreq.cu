// sitefreq.cu — per-site aggregation of call_mods results on the GPU: the reduction the reference does with
// a Python dict update per line (scripts/call_modification_frequency.py:17-49, scripts/txt_formater.py ModRecord /
// SiteStats): for every (chromosome, position) key the sums of prob_0 / prob_1, the counts of modified / all calls and
// the index of the first record.
//
// A segmented reduce: stable sort of (key, line index) — keys sort by (chromosome id, position), ties keep input order —
// then one thread per segment adds its records IN LINE ORDER in float64, which is exactly the order of the reference's
// `+=`, so the sums are bit-identical to the reference's.  The sort is CUB's radix sort (CUDA toolkit library code; this
// is a post-processing step, not the hot path); the segment detection / reduction kernels are ours.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace dsb {
namespace {

__global__ void sf_keys_kernel(const int32_t* __restrict__ chrom, const int64_t* __restrict__ pos, const uint8_t* __restrict__ use,
                               int64_t n, uint64_t* __restrict__ keys, int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // unused records (|p0 - p1| < prob_cf) sort behind everything else
  keys[i] = use[i] ? (((uint64_t)(uint32_t)chrom[i] << 40) | ((uint64_t)pos[i] & 0xffffffffffull)) : ~0ull;
  idx[i] = (int32_t)i;
}

// head[i] = 1 if sorted record i starts a segment of used records
__global__ void sf_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ seg_start, int32_t* __restrict__ n_seg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  if (k == ~0ull) return;
  if (i == 0 || keys[i - 1] != k) seg_start[atomicAdd(n_seg, 1)] = (int32_t)i;      // order restored by the host (sort by first index)
}

__global__ void sf_reduce_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx, const int32_t* __restrict__ seg_start,
                                 int32_t n_seg, int64_t n, const double* __restrict__ p0, const double* __restrict__ p1,
                                 const uint8_t* __restrict__ label, int64_t* __restrict__ first, double* __restrict__ s0,
                                 double* __restrict__ s1, int64_t* __restrict__ met, int64_t* __restrict__ cov) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  int64_t i = seg_start[s];
  const uint64_t k = keys[i];
  double a0 = 0.0, a1 = 0.0;
  int64_t m = 0, c = 0;
  first[s] = idx[i];                                        // stable sort: the first record of the segment is the earliest line
  for (; i < n && keys[i] == k; ++i) {
    const int32_t j = idx[i];
    a0 = __dadd_rn(a0, p0[j]);
    a1 = __dadd_rn(a1, p1[j]);
    m += label[j] ? 1 : 0;
    ++c;
  }
  s0[s] = a0; s1[s] = a1; met[s] = m; cov[s] = c;
}

}  // namespace

void site_frequency_host(dsb_ctx* ctx, int64_t n, const int32_t* chrom, const int64_t* pos, const double* p0, const double* p1,
                         const uint8_t* label, const uint8_t* use, int64_t capacity, int64_t* n_sites, int64_t* first,
                         double* s0, double* s1, int64_t* met, int64_t* cov) {
  if (n < 0 || !n_sites) fail(DSB_ERR_INVALID, "bad arguments");
  *n_sites = 0;
  if (n == 0) return;
  if (n >= (1ll << 31)) fail(DSB_ERR_INVALID, "more than 2^31 records in one call");
  if (!chrom || !pos || !p0 || !p1 || !label || !use) fail(DSB_ERR_INVALID, "null input pointer");
  DSB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = nullptr;
  DevBuf d_chrom, d_pos, d_p0, d_p1, d_lab, d_use, d_keys[2], d_idx[2], d_seg, d_cnt, d_tmp, d_first, d_s0, d_s1, d_met, d_cov;
  auto up = [&](DevBuf& b, const void* src, size_t bytes) { b.ensure(bytes); DSB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s)); };
  up(d_chrom, chrom, n * 4); up(d_pos, pos, n * 8); up(d_p0, p0, n * 8); up(d_p1, p1, n * 8); up(d_lab, label, n); up(d_use, use, n);
  for (int k = 0; k < 2; ++k) { d_keys[k].ensure(n * 8); d_idx[k].ensure(n * 4); }
  d_seg.ensure(n * 4);
  d_cnt.ensure(4);
  DSB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4, s));
  const unsigned grid = (unsigned)((n + 255) / 256);
  sf_keys_kernel<<<grid, 256, 0, s>>>(d_chrom.as<int32_t>(), d_pos.as<int64_t>(), d_use.as<uint8_t>(), n, d_keys[0].as<uint64_t>(), d_idx[0].as<int32_t>());
  DSB_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys[0].as<uint64_t>(), d_keys[1].as<uint64_t>(), d_idx[0].as<int32_t>(),
                                  d_idx[1].as<int32_t>(), (int)n, 0, 64, s);
  d_tmp.ensure(tmp_bytes + 16);
  DSB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys[0].as<uint64_t>(), d_keys[1].as<uint64_t>(), d_idx[0].as<int32_t>(),
                                           d_idx[1].as<int32_t>(), (int)n, 0, 64, s));          // LSD radix sort: stable
  sf_heads_kernel<<<grid, 256, 0, s>>>(d_keys[1].as<uint64_t>(), n, d_seg.as<int32_t>(), d_cnt.as<int32_t>());
  DSB_CUDA(cudaGetLastError());
  int32_t n_seg = 0;
  DSB_CUDA(cudaMemcpyAsync(&n_seg, d_cnt.p, 4, cudaMemcpyDeviceToHost, s));
  DSB_CUDA(cudaStreamSynchronize(s));
  ctx->launches += 3;
  *n_sites = n_seg;
  if (n_seg == 0) return;
  if (n_seg > capacity) fail(DSB_ERR_CAPACITY, "%d sites found but the output capacity is %lld", n_seg, (long long)capacity);
  if (!first || !s0 || !s1 || !met || !cov) fail(DSB_ERR_INVALID, "null output pointer");
  d_first.ensure((size_t)n_seg * 8); d_s0.ensure((size_t)n_seg * 8); d_s1.ensure((size_t)n_seg * 8); d_met.ensure((size_t)n_seg * 8); d_cov.ensure((size_t)n_seg * 8);
  sf_reduce_kernel<<<(unsigned)((n_seg + 127) / 128), 128, 0, s>>>(d_keys[1].as<uint64_t>(), d_idx[1].as<int32_t>(), d_seg.as<int32_t>(), n_seg, n,
                                                                d_p0.as<double>(), d_p1.as<double>(), d_lab.as<uint8_t>(), d_first.as<int64_t>(),
                                                                d_s0.as<double>(), d_s1.as<double>(), d_met.as<int64_t>(), d_cov.as<int64_t>());
  DSB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  auto down = [&](void* dst, DevBuf& b) { DSB_CUDA(cudaMemcpyAsync(dst, b.p, (size_t)n_seg * 8, cudaMemcpyDeviceToHost, s)); };
  down(first, d_first); down(s0, d_s0); down(s1, d_s1); down(met, d_met); down(cov, d_cov);
  DSB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace dsb
