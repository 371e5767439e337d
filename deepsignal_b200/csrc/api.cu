// api.cu — the extern "C" surface of libdsb200.so (include/dsb200.h).  No exception crosses the ABI.
#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace {
std::string g_create_error;

// read-and-reset of the head kernel's non-finite-logit counter (the caller has synchronised the producing stream)
int64_t take_nonfinite(dsb_ctx* ctx) {
  if (!ctx->nonfinite_dev) return 0;
  int h = 0;
  DSB_CUDA(cudaMemcpy(&h, ctx->nonfinite_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (h != 0) DSB_CUDA(cudaMemset(ctx->nonfinite_dev, 0, sizeof(int)));
  return h;
}

template <class F> int guarded(dsb_ctx* ctx, F&& f) {
  try {
    f();
    return DSB_OK;
  } catch (const dsb::Error& e) {
    if (ctx) ctx->err = e.msg; else g_create_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    if (ctx) ctx->err = "out of host memory"; else g_create_error = "out of host memory";
    return DSB_ERR_NOMEM;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what(); else g_create_error = e.what();
    return DSB_ERR_INVALID;
  } catch (...) {
    if (ctx) ctx->err = "unknown error"; else g_create_error = "unknown error";
    return DSB_ERR_INVALID;
  }
}

void ensure_streams(dsb_ctx* ctx) {
  if (ctx->stream_compute) return;
  DSB_CUDA(cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking));
  DSB_CUDA(cudaStreamCreateWithFlags(&ctx->stream_compute, cudaStreamNonBlocking));
  DSB_CUDA(cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    DSB_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
    DSB_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
    DSB_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
  }
}
}  // namespace

extern "C" {

int dsb_abi_version(void) { return DSB_ABI_VERSION; }

int dsb_create(int device, dsb_ctx** out) {
  if (!out) return DSB_ERR_INVALID;
  *out = nullptr;
  return guarded(nullptr, [&] {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      dsb::fail(DSB_ERR_CUDA, "no CUDA device available (%s); libdsb200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) dsb::fail(DSB_ERR_INVALID, "device %d out of range (found %d)", device, count);
    DSB_CUDA(cudaSetDevice(device));
    std::unique_ptr<dsb_ctx> ctx(new dsb_ctx());
    ctx->device = device;
    cudaDeviceProp prop;
    DSB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    *out = ctx.release();
  });
}

int dsb_destroy(dsb_ctx* ctx) {
  if (!ctx) return DSB_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& p : ctx->plan) p.reset();
  ctx->xw.reset();
  if (ctx->nonfinite_dev) cudaFree(ctx->nonfinite_dev);
  for (cudaStream_t s : {ctx->stream_h2d, ctx->stream_compute, ctx->stream_d2h}) if (s) cudaStreamDestroy(s);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  delete ctx;
  return DSB_OK;
}

const char* dsb_last_error(const dsb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dsb_load_weights(dsb_ctx* ctx, const dsb_tensor* tensors, int32_t n_tensors, uint32_t flags, int32_t kmer_len, int32_t signal_len) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    if (!tensors || n_tensors <= 0) dsb::fail(DSB_ERR_INVALID, "no tensors given");
    ctx->have_model = false;
    for (auto& p : ctx->plan) p.reset();
    dsb::fold_model(tensors, n_tensors, flags, kmer_len, signal_len, ctx->model);
    ctx->have_model = true;
  });
}

int dsb_set_max_chunk(dsb_ctx* ctx, int64_t max_chunk) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    if (max_chunk < 128 || max_chunk > (1 << 20)) dsb::fail(DSB_ERR_INVALID, "max_chunk must be in [128, 1048576]");
    ctx->max_chunk = (max_chunk + 127) / 128 * 128;
  });
}

int dsb_get_plan_options(const dsb_ctx* ctx, dsb_plan_options* out) {
  if (!ctx || !out) return DSB_ERR_INVALID;
  *out = ctx->opts;
  return DSB_OK;
}

int dsb_set_plan_options(dsb_ctx* ctx, const dsb_plan_options* opts) {
  if (!ctx || !opts) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    DSB_CUDA(cudaSetDevice(ctx->device));
    DSB_CUDA(cudaDeviceSynchronize());          // no launch of an old plan may still be in flight
    for (auto& p : ctx->plan) p.reset();
    ctx->opts = *opts;
  });
}

int dsb_nonfinite_count(dsb_ctx* ctx, int64_t* count) {
  if (!ctx || !count) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    DSB_CUDA(cudaSetDevice(ctx->device));
    *count = take_nonfinite(ctx);
  });
}

size_t dsb_workspace_bytes(const dsb_ctx* ctx, int64_t n, int32_t precision) {
  if (!ctx || !ctx->have_model) return 0;
  try {
    const int64_t chunk = std::min<int64_t>(ctx->max_chunk, (std::max<int64_t>(n, 1) + 127) / 128 * 128);
    return dsb::forward_workspace_bytes(ctx, chunk, precision);
  } catch (...) {
    return 0;
  }
}

int dsb_forward_dev(dsb_ctx* ctx, int32_t precision, int64_t n, const int32_t* kmer, const float* means, const float* stds,
                    const float* lens, const float* signals, float* probs, int32_t* pred, void* cuda_stream) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    dsb::forward_dev(ctx, precision, n, kmer, means, stds, lens, signals, probs, pred, s);
    if (ctx->prof.enabled) {
      DSB_CUDA(cudaStreamSynchronize(s));
      ctx->prof.collect();
    }
  });
}

int dsb_forward_host(dsb_ctx* ctx, int32_t precision, int64_t n, const int32_t* kmer, const float* means, const float* stds,
                     const float* lens, const float* signals, float* probs, int32_t* pred) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    if (!ctx->have_model) dsb::fail(DSB_ERR_STATE, "dsb_forward called before dsb_load_weights");
    if (n < 0) dsb::fail(DSB_ERR_INVALID, "negative sample count");
    if (n == 0) return;
    if (!signals || !probs || !pred) dsb::fail(DSB_ERR_INVALID, "null input/output pointer");
    DSB_CUDA(cudaSetDevice(ctx->device));
    ensure_streams(ctx);
    const int KL = ctx->model.kmer_len, S = ctx->model.signal_len;
    const int64_t hc = std::min<int64_t>(n, std::max<int64_t>(ctx->max_chunk, 65536));   // samples per pipeline slot
    const size_t in_bytes = (size_t)hc * (KL * 16 + S * 4);
    const size_t out_bytes = (size_t)hc * 12;
    for (int k = 0; k < 2; ++k) { ctx->stage_in[k].ensure(in_bytes + 4096); ctx->stage_out[k].ensure(out_bytes + 1024); }
    int slot = 0;
    int64_t used[2] = {0, 0};
    for (int64_t off = 0; off < n; off += hc, slot ^= 1) {
      const int64_t m = std::min<int64_t>(hc, n - off);
      uint8_t* base = ctx->stage_in[slot].as<uint8_t>();
      int32_t* d_kmer = reinterpret_cast<int32_t*>(base);
      float* d_means = reinterpret_cast<float*>(base + (size_t)hc * KL * 4);
      float* d_stds = reinterpret_cast<float*>(base + (size_t)hc * KL * 8);
      float* d_lens = reinterpret_cast<float*>(base + (size_t)hc * KL * 12);
      float* d_sig = reinterpret_cast<float*>(base + (size_t)hc * KL * 16);
      float* d_probs = ctx->stage_out[slot].as<float>();
      int32_t* d_pred = reinterpret_cast<int32_t*>(ctx->stage_out[slot].as<uint8_t>() + (size_t)hc * 8);
      // inputs of this slot may be overwritten once the previous forward that read them is done
      if (used[slot]) DSB_CUDA(cudaStreamWaitEvent(ctx->stream_h2d, ctx->ev_done[slot], 0));
      if (kmer) DSB_CUDA(cudaMemcpyAsync(d_kmer, kmer + off * KL, (size_t)m * KL * 4, cudaMemcpyHostToDevice, ctx->stream_h2d));
      if (means) DSB_CUDA(cudaMemcpyAsync(d_means, means + off * KL, (size_t)m * KL * 4, cudaMemcpyHostToDevice, ctx->stream_h2d));
      if (stds) DSB_CUDA(cudaMemcpyAsync(d_stds, stds + off * KL, (size_t)m * KL * 4, cudaMemcpyHostToDevice, ctx->stream_h2d));
      if (lens) DSB_CUDA(cudaMemcpyAsync(d_lens, lens + off * KL, (size_t)m * KL * 4, cudaMemcpyHostToDevice, ctx->stream_h2d));
      DSB_CUDA(cudaMemcpyAsync(d_sig, signals + off * S, (size_t)m * S * 4, cudaMemcpyHostToDevice, ctx->stream_h2d));
      DSB_CUDA(cudaEventRecord(ctx->ev_in[slot], ctx->stream_h2d));
      DSB_CUDA(cudaStreamWaitEvent(ctx->stream_compute, ctx->ev_in[slot], 0));
      if (used[slot]) DSB_CUDA(cudaStreamWaitEvent(ctx->stream_compute, ctx->ev_out[slot], 0));   // outputs drained
      dsb::forward_dev(ctx, precision, m, kmer ? d_kmer : nullptr, means ? d_means : nullptr, stds ? d_stds : nullptr,
                       lens ? d_lens : nullptr, d_sig, d_probs, d_pred, ctx->stream_compute);
      DSB_CUDA(cudaEventRecord(ctx->ev_done[slot], ctx->stream_compute));
      DSB_CUDA(cudaStreamWaitEvent(ctx->stream_d2h, ctx->ev_done[slot], 0));
      DSB_CUDA(cudaMemcpyAsync(probs + off * 2, d_probs, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream_d2h));
      DSB_CUDA(cudaMemcpyAsync(pred + off, d_pred, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream_d2h));
      DSB_CUDA(cudaEventRecord(ctx->ev_out[slot], ctx->stream_d2h));
      used[slot] = 1;
    }
    DSB_CUDA(cudaStreamSynchronize(ctx->stream_d2h));
    DSB_CUDA(cudaStreamSynchronize(ctx->stream_compute));
    if (ctx->prof.enabled) ctx->prof.collect();
    const int64_t bad = take_nonfinite(ctx);
    if (bad > 0)
      dsb::fail(DSB_ERR_NONFINITE, "%lld of %lld samples have non-finite logits (16-bit activation overflow; outputs were written, "
                "rerun with DSB_PREC_FP32)", (long long)bad, (long long)n);
  });
}

int dsb_set_profiling(dsb_ctx* ctx, int32_t enabled) {
  if (!ctx) return DSB_ERR_INVALID;
  ctx->prof.enabled = enabled != 0;
  ctx->prof.reset();
  return DSB_OK;
}

int dsb_get_profile(dsb_ctx* ctx, int32_t cls, double* ms_total, int64_t* launches, double* flops) {
  if (!ctx || cls < 0 || cls >= dsb::KC_COUNT) return DSB_ERR_INVALID;
  if (ms_total) *ms_total = ctx->prof.ms[cls];
  if (launches) *launches = ctx->prof.launches[cls];
  if (flops) *flops = ctx->prof.flops[cls];
  return DSB_OK;
}

int64_t dsb_launch_count(const dsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dsb_extract_dev(dsb_ctx* ctx, const dsb_reads* reads, const dsb_extract_params* params, const dsb_sites* sites,
                    int64_t* n_sites, double* read_shift, double* read_scale, void* cuda_stream) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    struct Collect {        // also when extract_dev leaves with DSB_ERR_CAPACITY
      dsb_ctx* c; cudaStream_t s;
      ~Collect() { if (c->prof.enabled) { cudaStreamSynchronize(s); c->prof.collect(); } }
    } collect{ctx, s};
    dsb::extract_dev(ctx, reads, params, sites, n_sites, read_shift, read_scale, s);
  });
}

int64_t dsb_extract_max_sites(const dsb_reads* r) {
  if (!r || !r->ev_off || r->n_reads <= 0) return 0;
  return r->ev_off[r->n_reads] - r->ev_off[0];
}

int dsb_extract_host(dsb_ctx* ctx, const dsb_reads* rd, const dsb_extract_params* params, const dsb_sites* sites,
                     int64_t* n_sites, double* read_shift, double* read_scale) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    if (!rd || !params || !sites || !n_sites) dsb::fail(DSB_ERR_INVALID, "null argument");
    *n_sites = 0;
    const int64_t R = rd->n_reads;
    if (R <= 0) return;
    DSB_CUDA(cudaSetDevice(ctx->device));
    ensure_streams(ctx);
    cudaStream_t s = ctx->stream_compute;
    if (rd->raw_off[0] != 0 || rd->ev_off[0] != 0) dsb::fail(DSB_ERR_INVALID, "raw_off[0] and ev_off[0] must be 0");
    const int64_t total_raw = rd->raw_off[R], total_ev = rd->ev_off[R];
    for (int64_t i = 0; i < total_ev; ++i)
      if (rd->ev_start[i] < 0) dsb::fail(DSB_ERR_INVALID, "event %lld has a negative start (read_start_rel_to_raw + start < 0)", (long long)i);
    dsb::DevBuf* in = ctx->x_in;
    auto up = [&](int k, const void* src, size_t bytes) -> void* {
      in[k].ensure(std::max<size_t>(bytes, 16));
      if (bytes) DSB_CUDA(cudaMemcpyAsync(in[k].p, src, bytes, cudaMemcpyHostToDevice, s));
      return in[k].p;
    };
    dsb_reads d = *rd;
    d.raw = (const int16_t*)up(0, rd->raw, total_raw * 2);
    d.raw_off = (const int64_t*)up(1, rd->raw_off, (R + 1) * 8);
    d.scaling = (const double*)up(2, rd->scaling, R * 8);
    d.offset = (const double*)up(3, rd->offset, R * 8);
    d.ev_start = (const int32_t*)up(4, rd->ev_start, total_ev * 4);
    d.ev_len = (const int32_t*)up(5, rd->ev_len, total_ev * 4);
    d.ev_base = (const uint8_t*)up(6, rd->ev_base, total_ev);
    d.ev_off = (const int64_t*)up(7, rd->ev_off, (R + 1) * 8);
    d.chrom_start = (const int64_t*)up(8, rd->chrom_start, R * 8);
    d.strand_minus = (const uint8_t*)up(9, rd->strand_minus, R);
    d.chrom_len = rd->chrom_len ? (const int64_t*)up(10, rd->chrom_len, R * 8) : nullptr;
    const int64_t cap = sites->capacity;
    const int K = params->kmer_len, S = params->signal_len;
    dsb::DevBuf* out = ctx->x_out;
    dsb_sites ds = *sites;
    struct Item { void** dev; const void* host; size_t bytes; int idx; };
    std::vector<Item> items;
    auto mk = [&](int idx, auto*& devp, const void* hostp, size_t bytes) {
      if (!hostp) { devp = nullptr; return; }
      out[idx].ensure(std::max<size_t>(bytes, 16));
      devp = reinterpret_cast<std::remove_reference_t<decltype(devp)>>(out[idx].p);
      items.push_back(Item{nullptr, hostp, bytes, idx});
    };
    mk(0, ds.site_read, sites->site_read, cap * 4);
    mk(1, ds.site_loc, sites->site_loc, cap * 4);
    mk(2, ds.pos, sites->pos, cap * 8);
    mk(3, ds.pos_in_strand, sites->pos_in_strand, cap * 8);
    mk(4, ds.kmer, sites->kmer, cap * K * 4);
    mk(5, ds.means, sites->means, cap * K * 4);
    mk(6, ds.stds, sites->stds, cap * K * 4);
    mk(7, ds.lens, sites->lens, cap * K * 4);
    mk(8, ds.signals, sites->signals, (size_t)cap * S * 4);
    mk(9, ds.means64, sites->means64, cap * K * 8);
    mk(10, ds.stds64, sites->stds64, cap * K * 8);
    mk(11, ds.lens_i32, sites->lens_i32, cap * K * 4);
    mk(12, ds.signals64, sites->signals64, (size_t)cap * S * 8);
    mk(13, ds.window, sites->window, cap * 3 * 4);
    dsb::DevBuf dshift, dscale;
    if (read_shift) dshift.ensure(R * 8);
    if (read_scale) dscale.ensure(R * 8);
    int code = DSB_OK;
    std::string msg;
    try {
      dsb::extract_dev(ctx, &d, params, &ds, n_sites, read_shift ? dshift.as<double>() : nullptr,
                       read_scale ? dscale.as<double>() : nullptr, s);
    } catch (const dsb::Error& e) {
      if (e.code != DSB_ERR_CAPACITY) throw;
      code = e.code;
      msg = e.msg;
    }
    const int64_t emit = std::min<int64_t>(*n_sites, cap);
    for (auto& it : items) {
      const size_t per = cap ? it.bytes / (size_t)cap : 0;
      if (emit) DSB_CUDA(cudaMemcpyAsync(const_cast<void*>(it.host), out[it.idx].p, per * (size_t)emit, cudaMemcpyDeviceToHost, s));
    }
    if (read_shift) DSB_CUDA(cudaMemcpyAsync(read_shift, dshift.p, R * 8, cudaMemcpyDeviceToHost, s));
    if (read_scale) DSB_CUDA(cudaMemcpyAsync(read_scale, dscale.p, R * 8, cudaMemcpyDeviceToHost, s));
    DSB_CUDA(cudaStreamSynchronize(s));
    if (ctx->prof.enabled) ctx->prof.collect();
    if (code != DSB_OK) throw dsb::Error{code, msg};
  });
}

int dsb_site_frequency_host(dsb_ctx* ctx, int64_t n, const int32_t* chrom_id, const int64_t* pos, const double* p0, const double* p1,
                            const uint8_t* label, const uint8_t* use, int64_t capacity, int64_t* n_sites, int64_t* first,
                            double* sum_p0, double* sum_p1, int64_t* met, int64_t* coverage) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    dsb::site_frequency_host(ctx, n, chrom_id, pos, p0, p1, label, use, capacity, n_sites, first, sum_p0, sum_p1, met, coverage);
  });
}

int dsb_extract_call_dev(dsb_ctx* ctx, int32_t precision, const dsb_reads* reads, const dsb_extract_params* params,
                         const dsb_sites* sites, int64_t* n_sites, float* probs, int32_t* pred, void* cuda_stream) {
  if (!ctx) return DSB_ERR_INVALID;
  return guarded(ctx, [&] {
    if (!sites || !sites->kmer || !sites->means || !sites->stds || !sites->lens || !sites->signals)
      dsb::fail(DSB_ERR_INVALID, "dsb_extract_call_dev needs the kmer/means/stds/lens/signals site buffers");
    if (!ctx->have_model) dsb::fail(DSB_ERR_STATE, "dsb_extract_call_dev called before dsb_load_weights");
    if (params && (params->kmer_len != ctx->model.kmer_len || params->signal_len != ctx->model.signal_len))
      dsb::fail(DSB_ERR_INVALID, "extraction k-mer/signal length differs from the loaded model");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    dsb::extract_dev(ctx, reads, params, sites, n_sites, nullptr, nullptr, s);
    dsb::forward_dev(ctx, precision, *n_sites, sites->kmer, sites->means, sites->stds, sites->lens, sites->signals, probs, pred, s);
    if (ctx->prof.enabled) {
      DSB_CUDA(cudaStreamSynchronize(s));
      ctx->prof.collect();
    }
  });
}

}  // extern "C"
