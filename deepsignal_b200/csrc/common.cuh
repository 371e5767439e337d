// common.cuh — context, error plumbing and shared device helpers of libdsb200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dsb200.h"

namespace dsb {

constexpr int HID = 256;          // LSTM hidden size (reference model.py:51)
constexpr int GATES = 4 * HID;    // 1024
constexpr int NLAYER = 3;         // reference model.py:50
constexpr int CH = 240;           // inception output channels (5 x 48, layers.py:139)
constexpr int NBLOCK = 11;
constexpr int HALO = 2;           // zero rows after every sample in the padded row layout

struct Error {
  int code;
  std::string msg;
};

// Throwing inside the library is fine; api.cu catches everything at the ABI.
[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error{code, buf};
}

#define DSB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      ::dsb::fail(DSB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                  __FILE__, __LINE__);                                                       \
  } while (0)

// RAII device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t n) {
    if (n <= bytes) return;
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) fail(DSB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
    bytes = n;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMallocHost(&p, n);
    if (e != cudaSuccess) fail(DSB_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", n, cudaGetErrorString(e));
    bytes = n;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- folded model (host, float64 -> float32) ------------------------------------------------
struct FoldedGemm {      // B operand of one GEMM: out[n] = sum_k A[k] * w[n*K + k] + bias[n]
  int N = 0, K = 0;
  std::vector<float> w, bias;
};

struct FoldedBlock {     // one inception block (reference layers.py:87-139)
  int cin = 0;
  FoldedGemm b1;         // [48][cin]      on maxpool3(x)
  FoldedGemm p1;         // [192][cin]     b2 | b3a | b4a | b5a | stem (stem has no ReLU)
  FoldedGemm b3;         // [48][3*32]
  FoldedGemm b4;         // [48][5*32]
  FoldedGemm b5b;        // [64][3*32]
  FoldedGemm b5c;        // [48][64]       + stem residual, then ReLU
};

struct FoldedLstm {      // one direction, one layer; columns permuted to 4*u+g (g = i,j,f,o)
  FoldedGemm g;          // layer0: K = 256 (h) + 8 feature rows ; else K = 512 (x | h)
};

struct FoldedModel {
  uint32_t flags = 0;
  int kmer_len = 17, signal_len = 360;
  int L1 = 180, L2 = 90, L3 = 45, L4 = 23;
  FoldedGemm conv1;      // [64][7]
  FoldedGemm conv2;      // [128][64]
  FoldedGemm conv3;      // [256][3*128]
  FoldedBlock blocks[NBLOCK];
  FoldedLstm lstm[2][NLAYER];
  std::vector<float> head_rnn;   // [512][2]
  std::vector<float> head_cnn;   // [L4*240][2], avgpool folded in
  bool has_cnn() const { return flags & DSB_FLAG_CNN; }
  bool has_rnn() const { return flags & DSB_FLAG_RNN; }
  bool has_base() const { return flags & DSB_FLAG_BASE; }
};

void fold_model(const dsb_tensor* tensors, int n, uint32_t flags, int kmer_len, int signal_len,
                FoldedModel& out);

// ---- profiling ------------------------------------------------------------------------------
// forward: KC_INCEPT_A is the dominant kernel, timed separately; KC_INCEPT_B the second inception kernel; KC_FRONT conv1..conv3.
// extraction (extract.cu): one class per kernel of the chain.
enum KClass { KC_GEMM = 0, KC_LSTM = 1, KC_ELEM = 2, KC_INCEPT_A = 3, KC_INCEPT_B = 4, KC_FRONT = 5,
              KC_X_NORM = 6, KC_X_SITES = 7, KC_X_LUT = 8, KC_X_BASE = 9, KC_X_FEAT = 10, KC_COUNT = 11 };

struct Profiler {
  bool enabled = false;
  struct Rec { cudaEvent_t a, b; int cls; double flops; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[KC_COUNT] = {};
  double flops[KC_COUNT] = {};
  int64_t launches[KC_COUNT] = {};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    DSB_CUDA(cudaEventCreate(&e));
    return e;
  }
  void collect() {   // call after the stream is synchronised
    for (auto& r : recs) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
        ms[r.cls] += t; flops[r.cls] += r.flops; launches[r.cls] += 1;
      }
      pool.push_back(r.a); pool.push_back(r.b);
    }
    recs.clear();
  }
  void reset() { for (int i = 0; i < KC_COUNT; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; } }
  ~Profiler() { for (auto e : pool) cudaEventDestroy(e); for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

struct ForwardPlan;   // forward.cu
struct ExtractWork;   // extract.cu
void destroy_plan(ForwardPlan* p);
void destroy_extract(ExtractWork* p);
struct PlanDeleter { void operator()(ForwardPlan* p) const { destroy_plan(p); } };
struct ExtractDeleter { void operator()(ExtractWork* p) const { destroy_extract(p); } };

}  // namespace dsb

struct dsb_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  std::string err;
  int64_t max_chunk = 37888;   // 148 SMs x 256 samples: the persistent LSTM kernel gives every SM exactly two item pairs
  int64_t launches = 0;
  bool have_model = false;
  dsb::FoldedModel model;
  dsb::Profiler prof;
  std::unique_ptr<dsb::ForwardPlan, dsb::PlanDeleter> plan[5];   // per precision
  dsb_plan_options opts{1, 1, 1, 0, 0, 0, 0, 0};
  int* nonfinite_dev = nullptr;   // device counter written by the head kernel
  std::unique_ptr<dsb::ExtractWork, dsb::ExtractDeleter> xw;
  // staging for the _host entry points
  dsb::DevBuf stage_in[2], stage_out[2];
  dsb::DevBuf x_in[12], x_out[14];   // dsb_extract_host staging
  dsb::PinnedBuf pin_in[2], pin_out[2];
  cudaStream_t stream_h2d = nullptr, stream_compute = nullptr, stream_d2h = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

namespace dsb {
struct Timer {       // optional CUDA-event bracket around one or more launches of a class (ctx->prof.enabled)
  dsb_ctx* ctx;
  cudaStream_t s;
  bool on;
  Profiler::Rec rec;
  Timer(dsb_ctx* c, cudaStream_t st, int cls, double flops) : ctx(c), s(st), on(c->prof.enabled) {
    if (!on) return;
    rec.a = c->prof.get();
    rec.b = c->prof.get();
    rec.cls = cls;
    rec.flops = flops;
    cudaEventRecord(rec.a, s);
  }
  ~Timer() {
    if (!on) return;
    cudaEventRecord(rec.b, s);
    ctx->prof.recs.push_back(rec);
  }
};

// forward.cu
void forward_dev(dsb_ctx* ctx, int precision, int64_t n, const int32_t* kmer, const float* means,
                 const float* stds, const float* lens, const float* signals, float* probs, int32_t* pred,
                 cudaStream_t stream);
size_t forward_workspace_bytes(const dsb_ctx* ctx, int64_t chunk, int precision);
// sitefreq.cu
void site_frequency_host(dsb_ctx* ctx, int64_t n, const int32_t* chrom, const int64_t* pos, const double* p0, const double* p1,
                         const uint8_t* label, const uint8_t* use, int64_t capacity, int64_t* n_sites, int64_t* first,
                         double* s0, double* s1, int64_t* met, int64_t* cov);
// extract.cu
void extract_dev(dsb_ctx* ctx, const dsb_reads* reads, const dsb_extract_params* params, const dsb_sites* sites,
                 int64_t* n_sites, double* read_shift, double* read_scale, cudaStream_t stream);
}  // namespace dsb
