// forward.cu — the call_mods forward (reference model.py:57-108 + layers.py) as a static plan of
// GEMM-with-epilogue ops and a few HBM-bound helpers, executed chunk by chunk on device-resident
// inputs.  Two backends share the plan: fp32 CUDA-core GEMMs (DSB_PREC_FP32) and 16-bit tcgen05 GEMMs
// (DSB_PREC_BF16 / DSB_PREC_FP16, and DSB_PREC_FP16X2 = the same kernels on split hi + lo fp16 activations).
// Which kernels realise the plan (fused or one launch per conv) comes from ctx->opts (dsb_plan_options).
//
// Activation layout ("padded rows"): a stage with L positions stores sample s, position p at row
// s*(L+HALO)+p of a row-major [rows][C] matrix; the HALO rows after each sample are zero, so a k-tap
// conv is a sum of row-shifted GEMMs and a sample never sees its neighbour.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <type_traits>

#include "elementwise.cuh"
#include "gemm_tc.cuh"
#include "incept_tc.cuh"
#include "lstm_seq.cuh"
#include "front_tc.cuh"

namespace dsb {

enum StepKind { ST_GEMM, ST_STEM, ST_POOL1, ST_POOL2, ST_FEAT, ST_HEAD, ST_ZERO, ST_INCEPT_A, ST_INCEPT_B, ST_LSTM_SEQ, ST_FRONT };

struct SegHost { const void* a; int ld; int cols; int col0; int drow; int K; int map_rows_per_sample; long long lo = 0; };   // lo: element offset of the lo twin (split mode)

struct Step {
  int kind = ST_GEMM;
  int stream = 0;             // 0 cnn, 1 fw lstm, 2 bw lstm
  int rows_per_sample = 1;    // M = n * rows_per_sample
  // GEMM
  GemmF32Params f32;
  TcLaunch tc;
  InceptALaunch ia;
  InceptBLaunch ib;
  LstmSeqLaunch seq;
  FrontLaunch front;
  int epi_kind = EPI_CONV;
  int kclass = KC_GEMM;
  double flops_per_row = 0;
  // helpers
  StemParams stem;
  const void* src = nullptr; void* dst = nullptr;
  int C = 0, ld = 0, Lin = 0, Lpin = 0, Lout = 0, Lpout = 0, pl = 0;
  long long lo_in = 0, lo_out = 0;      // split mode: element offsets of the lo twins of src / dst
  size_t zero_bytes_per_sample = 0;
  HeadParams head;
};

struct ForwardPlan {
  int precision = 0;
  int64_t cap = 0;            // samples per chunk the buffers are sized for
  DevBuf weights, work;
  std::vector<Step> steps;
  void* F = nullptr; int FW = 0;
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  bool multi_stream = true;
  size_t work_bytes = 0;
  ~ForwardPlan() {
    for (auto s : side) if (s) cudaStreamDestroy(s);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (auto e : ev_join) if (e) cudaEventDestroy(e);
  }
};

void destroy_plan(ForwardPlan* p) { delete p; }

namespace {

struct Arena {          // bump allocator over one DevBuf (two passes: size, then place)
  uint8_t* base = nullptr;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

// Ktot: columns of the device matrix; Klo: column where the lo parts of the weights start (0 = single-rounded weights)
template <class T> struct DevGemm { const T* w; const float* bias; int N; int Ktot; int Klo = 0; };

// Upload a folded GEMM: K is re-laid out as the concatenation of `parts` (src_k0, len, padded_len).
template <class T>
DevGemm<T> upload_gemm(Arena& wa, const FoldedGemm& g, const std::vector<std::array<int, 3>>& parts, cudaStream_t s,
                       bool with_lo = false) {
  int K1 = 0;
  for (auto& p : parts) K1 += p[2];
  // with_lo (DSB_PREC_FP16X3): [N][K1 | K1] = fp16(w) followed by fp16(w - fp16(w)), the third term of the split product
  const int Ktot = with_lo ? 2 * K1 : K1;
  DevGemm<T> d{nullptr, nullptr, g.N, Ktot, with_lo ? K1 : 0};
  T* w = reinterpret_cast<T*>(wa.take((size_t)g.N * Ktot * sizeof(T)));
  float* b = reinterpret_cast<float*>(wa.take((size_t)g.N * sizeof(float)));
  d.w = w;
  d.bias = b;
  if (wa.base) {
    std::vector<T> h((size_t)g.N * Ktot);
    for (int n = 0; n < g.N; ++n) {
      int ko = 0;
      for (auto& p : parts) {
        for (int k = 0; k < p[2]; ++k) {
          const float v = (k < p[1]) ? g.w[(size_t)n * g.K + p[0] + k] : 0.f;
          if constexpr (std::is_same<T, float>::value) h[(size_t)n * Ktot + ko + k] = v;
          else if constexpr (std::is_same<T, __half>::value) {
            const __half hi = __float2half_rn(v);
            h[(size_t)n * Ktot + ko + k] = hi;
            if (with_lo) h[(size_t)n * Ktot + K1 + ko + k] = __float2half_rn(v - __half2float(hi));
          } else h[(size_t)n * Ktot + ko + k] = __float2bfloat16_rn(v);
        }
        ko += p[2];
      }
    }
    DSB_CUDA(cudaMemcpyAsync(w, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    DSB_CUDA(cudaMemcpyAsync(b, g.bias.data(), g.N * sizeof(float), cudaMemcpyHostToDevice, s));
    DSB_CUDA(cudaStreamSynchronize(s));   // h goes out of scope
  }
  return d;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <class T, bool SPLIT = false>
struct Builder {
  static constexpr bool TC = !std::is_same<T, float>::value;
  static_assert(!SPLIT || std::is_same<T, __half>::value, "the split-operand mode carries fp16 hi + lo parts");
  dsb_ctx* ctx;
  ForwardPlan& plan;
  const FoldedModel& fm;
  int64_t cap;
  Arena wa, ba;       // weights, work buffers
  bool place;         // second pass: pointers are real
  cudaStream_t up_stream;
  bool w3 = false;    // DSB_PREC_FP16X3: weights as hi + lo too (three product terms), one launch per conv

  T* buf(size_t elems) { return reinterpret_cast<T*>(ba.take(elems * sizeof(T))); }
  // CNN activation matrix; in split mode its lo twin follows it (`elems` elements further on)
  T* abuf(size_t elems) { return reinterpret_cast<T*>(ba.take(elems * sizeof(T) * (SPLIT ? 2 : 1))); }
  static long long lo_of(size_t elems) { return SPLIT ? (long long)elems : 0; }

  // one GEMM step -------------------------------------------------------------------------------
  // alg_rows: rows per sample that carry real outputs (L for convs, 1 for the LSTM); kalg: algorithmic K
  void gemm(int stream, int rows_per_sample, int64_t rows_cap, const std::vector<SegHost>& segs,
            const DevGemm<T>& W, int BK, EpiParams epi, int epi_kind, int kclass, int alg_rows, int kalg) {
    Step st;
    st.kind = ST_GEMM;
    st.stream = stream;
    st.rows_per_sample = rows_per_sample;
    st.epi_kind = epi_kind;
    st.kclass = kclass;
    epi.N = W.N;
    epi.bias = W.bias;
    st.flops_per_row = 2.0 * (double)kalg * W.N * (double)alg_rows / (double)rows_per_sample;
    if (!place) { plan.steps.push_back(st); return; }
    if constexpr (!TC) {
      GemmF32Params& g = st.f32;
      g.nseg = (int)segs.size();
      int kt = 0;
      for (int i = 0; i < g.nseg; ++i) {
        g.seg[i] = SegF32{reinterpret_cast<const float*>(segs[i].a), segs[i].ld, segs[i].col0, segs[i].drow, round_up(segs[i].K, 16)};
        kt += g.seg[i].K;
      }
      if (kt != W.Ktot) fail(DSB_ERR_STATE, "internal: fp32 GEMM K mismatch %d vs %d", kt, W.Ktot);
      g.B = reinterpret_cast<const float*>(W.w);
      g.Ktot = W.Ktot;
      g.epi = epi;
    } else {
      TcLaunch& L = st.tc;
      L.BK = BK;
      L.epi_kind = epi_kind;
      L.fp16 = std::is_same<T, __half>::value;
      const void* map_base[2] = {nullptr, nullptr};
      int nmaps = 0, kt = 0, ns = 0;
      auto map_of = [&](const void* base, const SegHost& sg) {
        for (int j = 0; j < nmaps; ++j) if (map_base[j] == base) return j;
        if (nmaps == 2) fail(DSB_ERR_STATE, "internal: more than two A matrices in one GEMM");
        map_base[nmaps] = base;
        encode_map_16(&L.maps.a[nmaps], base, rows_cap, sg.cols, sg.ld, 128, BK, std::is_same<T, __half>::value);
        return nmaps++;
      };
      for (size_t i = 0; i < segs.size(); ++i) {
        const int nkb = round_up(segs[i].K, BK) / BK;
        const int wkb = kt / BK;
        if (ns + (segs[i].lo ? 2 : 1) + (W.Klo ? 1 : 0) > MAX_SEG) fail(DSB_ERR_STATE, "internal: too many GEMM segments");
        L.p.seg[ns++] = SegTC{map_of(segs[i].a, segs[i]), segs[i].col0, segs[i].drow, nkb, wkb};
        if (segs[i].lo)      // the lo part multiplies the same weight blocks
          L.p.seg[ns++] = SegTC{map_of(reinterpret_cast<const T*>(segs[i].a) + segs[i].lo, segs[i]), segs[i].col0, segs[i].drow, nkb, wkb};
        if (W.Klo)           // third term: hi activations x lo weights (xl . wl ~ 2^-22 is dropped)
          L.p.seg[ns++] = SegTC{map_of(segs[i].a, segs[i]), segs[i].col0, segs[i].drow, nkb, wkb + W.Klo / BK};
        kt += nkb * BK;
      }
      L.p.nseg = ns;
      if (nmaps == 1) L.maps.a[1] = L.maps.a[0];
      if (kt != (W.Klo ? W.Klo : W.Ktot)) fail(DSB_ERR_STATE, "internal: 16-bit GEMM K mismatch %d vs %d", kt, W.Klo ? W.Klo : W.Ktot);
      L.p.BN = W.N <= 256 ? W.N : 256;
      L.p.epi = epi;
      L.flops_per_row = st.flops_per_row;
      finalize_tc_launch(L);      // may shrink BN so that the weights stay resident in shared memory
      encode_map_16(&L.maps.b, W.w, W.N, W.Ktot, W.Ktot, L.p.BN, BK, std::is_same<T, __half>::value);
    }
    plan.steps.push_back(st);
  }

  std::vector<std::array<int, 3>> kparts(std::initializer_list<std::pair<int, int>> src, int BK) {
    std::vector<std::array<int, 3>> v;
    const int pad = TC ? BK : 16;
    for (auto& p : src) v.push_back({p.first, p.second, round_up(p.second, pad)});
    return v;
  }
  std::vector<std::array<int, 3>> tap_parts(int taps, int cin, int BK) {
    std::vector<std::array<int, 3>> v;
    const int pad = TC ? BK : 16;
    for (int t = 0; t < taps; ++t) v.push_back({t * cin, cin, round_up(cin, pad)});
    return v;
  }

  void build() {
    const dsb_plan_options& opt = ctx->opts;
    const int S = fm.signal_len, KL = fm.kmer_len;
    const int L2 = fm.L2, L3 = fm.L3, L4 = fm.L4;
    const int Lp2 = L2 + HALO, Lp3 = L3 + HALO, Lp4 = L4 + HALO;
    (void)Lp3; (void)Lp4;
    const int64_t R2 = cap * Lp2;
    void* final_y = nullptr;
    int final_Lp = Lp4;
    long long final_lo = 0;

    if (fm.has_cnn()) {
      // ---- device weights ----
      float* w1 = reinterpret_cast<float*>(wa.take(64 * 7 * sizeof(float)));
      float* b1 = reinterpret_cast<float*>(wa.take(64 * sizeof(float)));
      if (place) {
        DSB_CUDA(cudaMemcpyAsync(w1, fm.conv1.w.data(), 64 * 7 * sizeof(float), cudaMemcpyHostToDevice, up_stream));
        DSB_CUDA(cudaMemcpyAsync(b1, fm.conv1.bias.data(), 64 * sizeof(float), cudaMemcpyHostToDevice, up_stream));
      }
      DevGemm<T> Wc2 = upload_gemm<T>(wa, fm.conv2, kparts({{0, 64}}, 64), up_stream, w3);
      DevGemm<T> Wc3 = upload_gemm<T>(wa, fm.conv3, tap_parts(3, 128, 64), up_stream, w3);
      // ---- buffers (split mode: every matrix is followed by its lo twin) ----
      constexpr int YLD = 256;            // 240 channels padded to a 512-byte row pitch: every TMA row chunk is one aligned 128-byte line
      const size_t eP1 = (size_t)R2 * 64, eC2 = (size_t)R2 * 128, eX = (size_t)R2 * 256, eY = (size_t)R2 * YLD;
      const size_t e32 = (size_t)R2 * 32, e48 = (size_t)R2 * 48, eT = (size_t)R2 * 144, eU = (size_t)R2 * 64;
      T* P1 = abuf(eP1);
      T* C2 = abuf(eC2);
      T* X0 = abuf(eX);
      T* XP = abuf(eX);
      T* YA = abuf(eY);
      T* YB = abuf(eY);
      T* TP3 = abuf(e32);                 // fused path: the block's 1x1 outputs as separate planes (contiguous TMA boxes)
      T* TP4 = abuf(e32);
      T* TP5 = abuf(e32);
      T* TPS = abuf(e48);
      T* Tb = abuf(eT);
      T* Ub = abuf(eU);
      const long long loP1 = lo_of(eP1), loC2 = lo_of(eC2), loX = lo_of(eX), loY = lo_of(eY), lo32 = lo_of(e32), lo48 = lo_of(e48),
                      loT = lo_of(eT), loU = lo_of(eU);

      auto conv_epi = [&](void* d0, int ld0, int c0, uint32_t relu, int Lp, int L, long long lo0) {
        EpiParams e;
        e.d0 = d0; e.ld0 = ld0; e.c0 = c0; e.split = 1 << 30; e.relu_mask = relu; e.Lp = Lp; e.L = L; e.lo_off0 = lo0;
        return e;
      };
      const int conv_total = std::max((fm.L1 - 1) * 2 + 7 - S, 0);
      const int pool_total = std::max((L2 - 1) * 2 + 3 - fm.L1, 0);
      // un-fused reference (fused_front = 0): stem on the CUDA cores in fp32 followed by conv2 / conv3 as generic GEMM launches
      const bool front = TC && opt.fused_front != 0 && !w3;
      T* wB1 = reinterpret_cast<T*>(wa.take((size_t)192 * 64 * sizeof(T)));
      if (front) {
        if constexpr (TC) {
          // conv1 + pool + conv2 + conv3 in one kernel (front2_tc.cu): P1 and C2 stay in shared memory
          Step st;
          st.kind = ST_FRONT;
          st.stream = 0;
          st.rows_per_sample = Lp2;
          st.kclass = KC_FRONT;
          st.flops_per_row = 2.0 * (64.0 * 128.0 + 384.0 * 256.0) * (double)L2 / (double)Lp2;
          if (place) {
            const bool f16 = std::is_same<T, __half>::value;
            FrontLaunch& Q = st.front;
            Q.fp16 = f16;
            Q.split = SPLIT;
            Q.p.lo_x0 = loX;
            encode_map_16(&Q.maps.w2, Wc2.w, 128, Wc2.Ktot, Wc2.Ktot, 128, 64, f16);
            if (Wc3.N != 256 || Wc2.N != 128) fail(DSB_ERR_STATE, "internal: front kernel expects conv2 N=128, conv3 N=256");
            encode_map_16(&Q.maps.w3, Wc3.w, 256, Wc3.Ktot, Wc3.Ktot, 128, 32, f16);
            Q.p.S = S; Q.p.w1 = w1; Q.p.b1 = b1; Q.p.b2 = Wc2.bias; Q.p.b3 = Wc3.bias;
            Q.p.L1 = fm.L1; Q.p.L = L2; Q.p.Lp = Lp2; Q.p.conv_pl = conv_total / 2; Q.p.pool_pl = pool_total / 2;
            Q.p.X0 = X0; Q.p.ldX = 256;
            // conv1 as a 16-bit (hi | hi | lo) x (hi | lo | hi) MMA -- weights of channel c shifted to pool tap a in row a*64+c
            std::vector<__half> hB1((size_t)192 * 64, __float2half_rn(0.f));     // fp16 parts in both precision modes
            for (int a = 0; a < 3; ++a)
              for (int c = 0; c < 64; ++c)
                for (int t = 0; t < 7; ++t) {
                  const float wf = fm.conv1.w[(size_t)c * 7 + t];
                  const __half wh = __float2half_rn(wf);
                  const __half wl = __float2half_rn(wf - __half2float(wh));
                  const size_t row = (size_t)(a * 64 + c) * 64;
                  const int k = 2 * a + t;
                  hB1[row + k] = wh; hB1[row + 11 + k] = wh; hB1[row + 22 + k] = wl;
                  hB1[row + 33 + a] = __float2half_rn(1.f);      // picks up the kernel's out-of-range penalty of pool tap a
                }
            DSB_CUDA(cudaMemcpyAsync(wB1, hB1.data(), hB1.size() * sizeof(__half), cudaMemcpyHostToDevice, up_stream));
            DSB_CUDA(cudaStreamSynchronize(up_stream));
            encode_map_16(&Q.maps.w1, wB1, 192, 64, 64, 192, 64, true);
            for (int k = 0; k < 64; ++k) Q.p.b1v[k] = fm.conv1.bias[k];
            for (int k = 0; k < 128; ++k) Q.p.b2v[k] = fm.conv2.bias[k];
            for (int k = 0; k < 256; ++k) Q.p.b3v[k] = fm.conv3.bias[k];
          }
          plan.steps.push_back(st);
        }
      } else {
        {
          Step st;
          st.kind = ST_STEM;
          st.stream = 0;
          st.stem = StemParams{nullptr, S, w1, b1, fm.L1, L2, Lp2, conv_total / 2, pool_total / 2, P1, 0, loP1};
          plan.steps.push_back(st);
        }
        gemm(0, Lp2, R2, {SegHost{P1, 64, 64, 0, 0, 64, Lp2, loP1}}, Wc2, 64, conv_epi(C2, 128, 0, 0xffffffffu, Lp2, L2, loC2), EPI_CONV, KC_GEMM, L2, 64);
        gemm(0, Lp2, R2, {SegHost{C2, 128, 128, 0, -1, 128, Lp2, loC2}, SegHost{C2, 128, 128, 0, 0, 128, Lp2, loC2}, SegHost{C2, 128, 128, 0, 1, 128, Lp2, loC2}},
             Wc3, 64, conv_epi(X0, 256, 0, 0xffffffffu, Lp2, L2, loX), EPI_CONV, KC_GEMM, L2, 384);
      }

      const T* X = X0;
      int xld = 256;
      long long xlo = loX;
      int L = L2, Lp = Lp2;
      T* Y = YA;
      for (int i = 0; i < NBLOCK; ++i) {
        const FoldedBlock& fb = fm.blocks[i];
        const int cin = fb.cin;
        const int64_t rows_cap = cap * Lp;
        DevGemm<T> Wb1 = upload_gemm<T>(wa, fb.b1, kparts({{0, cin}}, 64), up_stream, w3);
        DevGemm<T> Wp1 = upload_gemm<T>(wa, fb.p1, kparts({{0, cin}}, 64), up_stream, w3);
        DevGemm<T> Wb3 = upload_gemm<T>(wa, fb.b3, tap_parts(3, 32, 32), up_stream, w3);
        DevGemm<T> Wb4 = upload_gemm<T>(wa, fb.b4, tap_parts(5, 32, 32), up_stream, w3);
        DevGemm<T> Wb5b = upload_gemm<T>(wa, fb.b5b, tap_parts(3, 32, 32), up_stream, w3);
        DevGemm<T> Wb5c = upload_gemm<T>(wa, fb.b5c, kparts({{0, 64}}, 64), up_stream, w3);
        const bool fused = TC && opt.fused_blocks != 0 && !w3;
        if (fused) {
          if constexpr (TC) {
            const bool f16 = std::is_same<T, __half>::value;
            // concatenated biases in accumulator-column order
            float* biasA = reinterpret_cast<float*>(wa.take(240 * sizeof(float)));
            float* biasB = reinterpret_cast<float*>(wa.take(208 * sizeof(float)));
            Step sa, sb;
            sa.kind = ST_INCEPT_A; sa.stream = 0; sa.rows_per_sample = Lp; sa.kclass = KC_INCEPT_A;
            sb.kind = ST_INCEPT_B; sb.stream = 0; sb.rows_per_sample = Lp; sb.kclass = KC_INCEPT_B;
            sa.flops_per_row = 2.0 * (double)cin * 240.0 * (double)L / (double)Lp;
            sb.flops_per_row = 2.0 * (96.0 * 48 + 160.0 * 48 + 96.0 * 64 + 64.0 * 48) * (double)L / (double)Lp;
            if (place) {
              std::vector<float> ha(240), hb(208);
              for (int k = 0; k < 48; ++k) ha[k] = fb.b1.bias[k];          // accumulator order: b1 | p1
              for (int k = 0; k < 192; ++k) ha[48 + k] = fb.p1.bias[k];
              for (int k = 0; k < 48; ++k) { hb[k] = fb.b3.bias[k]; hb[48 + k] = fb.b4.bias[k]; hb[96 + k] = fb.b5c.bias[k]; }
              for (int k = 0; k < 64; ++k) hb[144 + k] = fb.b5b.bias[k];     // b3 | b4 | b5c | b5b
              DSB_CUDA(cudaMemcpyAsync(biasA, ha.data(), 240 * sizeof(float), cudaMemcpyHostToDevice, up_stream));
              DSB_CUDA(cudaMemcpyAsync(biasB, hb.data(), 208 * sizeof(float), cudaMemcpyHostToDevice, up_stream));
              DSB_CUDA(cudaStreamSynchronize(up_stream));
              InceptALaunch& A = sa.ia;
              A.fp16 = f16;
              A.split = SPLIT;
              encode_map_16(&A.maps.x, X, rows_cap, cin, xld, 128, 64, f16);
              encode_map_16(&A.maps.wp1, Wp1.w, 192, Wp1.Ktot, Wp1.Ktot, 192, 64, f16);
              encode_map_16(&A.maps.wb1, Wb1.w, 48, Wb1.Ktot, Wb1.Ktot, 48, 64, f16);
              encode_map_16(&A.maps.ys, Y, rows_cap, CH, YLD, 32, 32, f16);                // epilogue TMA stores
              encode_map_16(&A.maps.p3s, TP3, rows_cap, 32, 32, 32, 32, f16);
              encode_map_16(&A.maps.p4s, TP4, rows_cap, 32, 32, 32, 32, f16);
              encode_map_16(&A.maps.p5s, TP5, rows_cap, 32, 32, 32, 32, f16);
              encode_map_16(&A.maps.pss, TPS, rows_cap, 48, 48, 32, 32, f16);
              if (SPLIT) {
                encode_map_16(&A.maps.xl, X + xlo, rows_cap, cin, xld, 128, 64, f16);
                encode_map_16(&A.maps.ysl, Y + loY, rows_cap, CH, YLD, 32, 32, f16);
                encode_map_16(&A.maps.p3sl, TP3 + lo32, rows_cap, 32, 32, 32, 32, f16);
                encode_map_16(&A.maps.p4sl, TP4 + lo32, rows_cap, 32, 32, 32, 32, f16);
                encode_map_16(&A.maps.p5sl, TP5 + lo32, rows_cap, 32, 32, 32, 32, f16);
                encode_map_16(&A.maps.pssl, TPS + lo48, rows_cap, 48, 48, 32, 32, f16);
                A.p.lo_y = loY; A.p.lo_p3 = A.p.lo_p4 = A.p.lo_p5 = lo32; A.p.lo_ps = lo48;
              }
              for (int k = 0; k < 240; ++k) A.p.biasv[k] = ha[k];
              A.p.Lp = Lp; A.p.L = L; A.p.Y = Y; A.p.ldY = YLD; A.p.P3 = TP3; A.p.P4 = TP4; A.p.P5 = TP5; A.p.PS = TPS; A.p.bias = biasA;
              InceptBLaunch& B = sb.ib;
              B.fp16 = f16;
              B.split = SPLIT;
              encode_map_16(&B.maps.t3, TP3, rows_cap, 32, 32, 136, 32, f16);      // 136-row boxes: taps are row-shifted
              encode_map_16(&B.maps.t4, TP4, rows_cap, 32, 32, 136, 32, f16);      // descriptors on one staged tile
              encode_map_16(&B.maps.t5, TP5, rows_cap, 32, 32, 136, 32, f16);
              encode_map_16(&B.maps.w3, Wb3.w, 48, Wb3.Ktot, Wb3.Ktot, 48, 32, f16);
              encode_map_16(&B.maps.w4, Wb4.w, 48, Wb4.Ktot, Wb4.Ktot, 48, 32, f16);
              encode_map_16(&B.maps.w5b, Wb5b.w, 64, Wb5b.Ktot, Wb5b.Ktot, 64, 32, f16);
              encode_map_16(&B.maps.w5c, Wb5c.w, 48, Wb5c.Ktot, Wb5c.Ktot, 48, 32, f16);
              encode_map_16(&B.maps.ys, Y, rows_cap, CH, YLD, 32, 32, f16);               // epilogue TMA stores
              if (SPLIT) {
                encode_map_16(&B.maps.t3l, TP3 + lo32, rows_cap, 32, 32, 136, 32, f16);
                encode_map_16(&B.maps.t4l, TP4 + lo32, rows_cap, 32, 32, 136, 32, f16);
                encode_map_16(&B.maps.t5l, TP5 + lo32, rows_cap, 32, 32, 136, 32, f16);
                encode_map_16(&B.maps.ysl, Y + loY, rows_cap, CH, YLD, 32, 32, f16);
                B.p.lo_ps = lo48;
              }
              for (int k = 0; k < 208; ++k) B.p.biasv[k] = hb[k];
              B.p.Lp = Lp; B.p.L = L; B.p.PS = TPS; B.p.Y = Y; B.p.ldY = YLD; B.p.bias = biasB;
            }
            plan.steps.push_back(sa);
            plan.steps.push_back(sb);
          }
        } else {
          {   // branch-1 max pool (1x3, stride 1)
            Step st;
            st.kind = ST_POOL1;
            st.stream = 0;
            st.rows_per_sample = Lp;
            st.src = X; st.dst = XP; st.C = cin; st.ld = xld; st.lo_in = xlo; st.lo_out = loX;
            st.Lin = L; st.Lpin = Lp; st.Lout = L; st.Lpout = Lp; st.pl = 1;
            plan.steps.push_back(st);
          }
          gemm(0, Lp, rows_cap, {SegHost{XP, xld, cin, 0, 0, cin, Lp, loX}}, Wb1, 64, conv_epi(Y, YLD, 0, 0xffffffffu, Lp, L, loY), EPI_CONV, KC_GEMM, L, cin);
          {   // fused 1x1s: b2 -> Y[:,48:96] ; b3a|b4a|b5a|stem -> T[:,0:144] (stem without ReLU)
            EpiParams e = conv_epi(Y, YLD, 48, 0x1ffu /* cols 0..143 */, Lp, L, loY);
            e.split = 48;
            e.d1 = Tb; e.ld1 = 144; e.c1 = 0; e.lo_off1 = loT;
            gemm(0, Lp, rows_cap, {SegHost{X, xld, cin, 0, 0, cin, Lp, xlo}}, Wp1, 64, e, EPI_CONV, KC_GEMM, L, cin);
          }
          gemm(0, Lp, rows_cap, {SegHost{Tb, 144, 144, 0, -1, 32, Lp, loT}, SegHost{Tb, 144, 144, 0, 0, 32, Lp, loT}, SegHost{Tb, 144, 144, 0, 1, 32, Lp, loT}},
               Wb3, 32, conv_epi(Y, YLD, 96, 0xffffffffu, Lp, L, loY), EPI_CONV, KC_GEMM, L, 96);
          gemm(0, Lp, rows_cap, {SegHost{Tb, 144, 144, 32, -2, 32, Lp, loT}, SegHost{Tb, 144, 144, 32, -1, 32, Lp, loT}, SegHost{Tb, 144, 144, 32, 0, 32, Lp, loT},
                                 SegHost{Tb, 144, 144, 32, 1, 32, Lp, loT}, SegHost{Tb, 144, 144, 32, 2, 32, Lp, loT}},
               Wb4, 32, conv_epi(Y, YLD, 144, 0xffffffffu, Lp, L, loY), EPI_CONV, KC_GEMM, L, 160);
          gemm(0, Lp, rows_cap, {SegHost{Tb, 144, 144, 64, -1, 32, Lp, loT}, SegHost{Tb, 144, 144, 64, 0, 32, Lp, loT}, SegHost{Tb, 144, 144, 64, 1, 32, Lp, loT}},
               Wb5b, 32, conv_epi(Ub, 64, 0, 0xffffffffu, Lp, L, loU), EPI_CONV, KC_GEMM, L, 96);
          {
            EpiParams e = conv_epi(Y, YLD, 192, 0xffffffffu, Lp, L, loY);
            e.res = Tb; e.res_ld = 144; e.res_c0 = 96; e.res_lo_off = loT;
            gemm(0, Lp, rows_cap, {SegHost{Ub, 64, 64, 0, 0, 64, Lp, loU}}, Wb5c, 64, e, EPI_CONV, KC_GEMM, L, 64);
          }
        }
        X = Y;
        xld = YLD;
        xlo = loY;
        T* other = (Y == YA) ? YB : YA;
        if (i == 2 || i == 7) {   // max pool 1x3 stride 2 between stages
          const int Lo = (i == 2) ? L3 : L4, Lpo = Lo + HALO;
          const int total = std::max((Lo - 1) * 2 + 3 - L, 0);
          Step st;
          st.kind = ST_POOL2;
          st.stream = 0;
          st.src = Y; st.dst = other; st.C = CH; st.ld = YLD; st.Lin = L; st.Lpin = Lp; st.Lout = Lo; st.Lpout = Lpo; st.pl = total / 2;
          st.lo_in = loY; st.lo_out = loY;
          plan.steps.push_back(st);
          X = other;
          other = Y;
          L = Lo; Lp = Lpo;
        }
        Y = other;
      }
      final_y = const_cast<T*>(X);
      final_Lp = Lp;
      final_lo = xlo;
    }

    void* hfw = nullptr;
    void* hbw = nullptr;
    if (fm.has_rnn()) {
      const int FW = TC ? 64 : 16;
      plan.FW = FW;
      T* F = buf((size_t)KL * cap * FW);
      plan.F = F;
      {
        Step st;
        st.kind = ST_FEAT;
        st.stream = 0;    // runs before the fork
        plan.steps.insert(plan.steps.begin(), st);
      }
      T* H[2][NLAYER];
      float* Cst[2][NLAYER];
      DevGemm<T> Wl[2][NLAYER];
      for (int d = 0; d < 2; ++d)
        for (int l = 0; l < NLAYER; ++l) {
          H[d][l] = buf((size_t)(KL + 1) * cap * HID);
          Cst[d][l] = reinterpret_cast<float*>(ba.take((size_t)cap * HID * sizeof(float)));
          Wl[d][l] = (l == 0)
              ? upload_gemm<T>(wa, fm.lstm[d][l].g, std::vector<std::array<int, 3>>{{0, HID, HID}, {HID, 8, FW}}, up_stream)
              : upload_gemm<T>(wa, fm.lstm[d][l].g, std::vector<std::array<int, 3>>{{0, HID, HID}, {HID, HID, HID}}, up_stream);
        }
      const bool seq = TC && opt.lstm_seq != 0 && NLAYER == LSTM_SEQ_LAYERS;
      if (seq) {
        if constexpr (TC) {
          // one persistent kernel for all directions, layers and time steps of the chunk (lstm_seq2.cu)
          Step st;
          st.kind = ST_LSTM_SEQ;
          st.stream = 1;
          st.rows_per_sample = 1;
          st.kclass = KC_LSTM;
          st.flops_per_row = 2.0 * 2.0 * (double)KL * (double)GATES * ((double)HID + 2.0 * (2.0 * HID));
          if (place) {
            const bool f16 = std::is_same<T, __half>::value;
            LstmSeqLaunch& Q = st.seq;
            Q.fp16 = f16;
            encode_map_16(&Q.maps.f, F, (int64_t)KL * cap, FW, FW, 128, 64, f16);
            for (int d = 0; d < 2; ++d) {
              Q.p.c[d] = Cst[d][0];
              for (int l = 0; l < NLAYER; ++l) {
                encode_map_16(&Q.maps.h[d][l], H[d][l], (int64_t)(KL + 1) * cap, HID, HID, 128, 64, f16);
                encode_map_16(&Q.maps.w2[d][l], Wl[d][l].w, GATES, Wl[d][l].Ktot, Wl[d][l].Ktot, 128, 64, f16);
                Q.p.bias[d][l] = Wl[d][l].bias;
                Q.p.H[d][l] = H[d][l];
              }
            }
            Q.p.cap = cap;
            Q.p.KL = KL;
          }
          plan.steps.push_back(st);
        }
      } else {
        for (int d = 0; d < 2; ++d) {
          for (int l = 0; l < NLAYER; ++l) {
            Step z;
            z.kind = ST_ZERO;
            z.stream = 1 + d;
            z.dst = H[d][l]; z.zero_bytes_per_sample = HID * sizeof(T);      // slot 0 = h_init
            plan.steps.push_back(z);
            Step zc = z;
            zc.dst = Cst[d][l]; zc.zero_bytes_per_sample = HID * sizeof(float);
            plan.steps.push_back(zc);
          }
          for (int l = 0; l < NLAYER; ++l) {
            for (int t = 0; t < KL; ++t) {
              const int pos = (d == 0) ? t : KL - 1 - t;          // bw consumes the k-mer reversed
              T* hprev = H[d][l] + (size_t)t * cap * HID;
              T* hnext = H[d][l] + (size_t)(t + 1) * cap * HID;
              EpiParams e;
              e.c = Cst[d][l];
              e.h_out = hnext;
              std::vector<SegHost> segs;
              if (l == 0) {
                segs.push_back(SegHost{hprev, HID, HID, 0, 0, HID, 1});
                segs.push_back(SegHost{F + (size_t)pos * cap * FW, FW, FW, 0, 0, FW, 1});
              } else {
                segs.push_back(SegHost{H[d][l - 1] + (size_t)(t + 1) * cap * HID, HID, HID, 0, 0, HID, 1});
                segs.push_back(SegHost{hprev, HID, HID, 0, 0, HID, 1});
              }
              gemm(1 + d, 1, cap, segs, Wl[d][l], 64, e, EPI_LSTM, KC_LSTM, 1, l == 0 ? HID : 2 * HID);
            }
          }
        }
      }
      hfw = H[0][NLAYER - 1] + (size_t)KL * cap * HID;
      hbw = H[1][NLAYER - 1] + (size_t)KL * cap * HID;
    }

    // head
    {
      float* Wc = reinterpret_cast<float*>(wa.take(std::max<size_t>(fm.head_cnn.size(), 4) * sizeof(float)));
      float* Wr = reinterpret_cast<float*>(wa.take(std::max<size_t>(fm.head_rnn.size(), 4) * sizeof(float)));
      if (place) {
        if (!fm.head_cnn.empty()) DSB_CUDA(cudaMemcpyAsync(Wc, fm.head_cnn.data(), fm.head_cnn.size() * sizeof(float), cudaMemcpyHostToDevice, up_stream));
        if (!fm.head_rnn.empty()) DSB_CUDA(cudaMemcpyAsync(Wr, fm.head_rnn.data(), fm.head_rnn.size() * sizeof(float), cudaMemcpyHostToDevice, up_stream));
        DSB_CUDA(cudaStreamSynchronize(up_stream));
      }
      Step st;
      st.kind = ST_HEAD;
      st.stream = 0;
      st.head = HeadParams{final_y, final_Lp, fm.L4, 256, hfw, hbw, Wc, Wr, nullptr, nullptr, 0, final_lo, ctx->nonfinite_dev};
      plan.steps.push_back(st);
    }
  }
};

template <class T, bool SPLIT>
void build_plan(dsb_ctx* ctx, ForwardPlan& plan, int64_t cap, cudaStream_t s, bool w3 = false) {
  // pass 1: sizes
  {
    ForwardPlan scratch;
    Builder<T, SPLIT> b{ctx, scratch, ctx->model, cap, Arena{}, Arena{}, false, s};
    b.w3 = w3;
    b.build();
    plan.weights.ensure(b.wa.off + 1024);
    plan.work.ensure(b.ba.off + 1024);
    plan.work_bytes = b.ba.off + 1024;
  }
  DSB_CUDA(cudaMemsetAsync(plan.work.p, 0, plan.work.bytes, s));
  plan.steps.clear();
  Arena wa, ba;
  wa.base = reinterpret_cast<uint8_t*>(plan.weights.p);
  ba.base = reinterpret_cast<uint8_t*>(plan.work.p);
  Builder<T, SPLIT> b{ctx, plan, ctx->model, cap, wa, ba, true, s};
  b.w3 = w3;
  b.build();
  DSB_CUDA(cudaStreamSynchronize(s));
  plan.cap = cap;
}

template <class T, bool SPLIT>
void run_chunk(dsb_ctx* ctx, ForwardPlan& plan, int64_t n, const int32_t* kmer, const float* means, const float* stds,
               const float* lens, const float* signals, float* probs, int32_t* pred, cudaStream_t main_stream) {
  const FoldedModel& fm = ctx->model;
  constexpr bool TC16 = !std::is_same<T, float>::value;
  const bool multi = plan.multi_stream && !ctx->prof.enabled && fm.has_rnn() && fm.has_cnn();
  cudaStream_t streams[3] = {main_stream, multi ? plan.side[0] : main_stream, multi ? plan.side[1] : main_stream};
  // every persistent kernel sizes its grid to the whole GPU; the LSTM stream (L2 / tensor bound) and the CNN stream
  // (HBM bound) time-slice the SMs (partitioning them measured slower: profiles/r1_ablation.md)
  const dsb_plan_options& opt = ctx->opts;
  int sms[3] = {ctx->sm_count, ctx->sm_count, ctx->sm_count};
  if (multi && opt.lstm_sms > 0 && opt.lstm_sms < ctx->sm_count) {       // SM budget per stream (dsb_plan_options.lstm_sms)
    sms[1] = sms[2] = opt.lstm_sms & ~1;
    sms[0] = ctx->sm_count - sms[1];
  }
  bool forked = false;
  for (const Step& st : plan.steps) {
    cudaStream_t s = streams[st.stream];
    if (multi && !forked && st.kind != ST_FEAT) {
      // everything after the feature rows may run on the three streams concurrently
      DSB_CUDA(cudaEventRecord(plan.ev_fork, main_stream));
      DSB_CUDA(cudaStreamWaitEvent(plan.side[0], plan.ev_fork, 0));
      DSB_CUDA(cudaStreamWaitEvent(plan.side[1], plan.ev_fork, 0));
      forked = true;
    }
    const int64_t M = n * st.rows_per_sample;
    switch (st.kind) {
      case ST_GEMM: {
        Timer tm(ctx, s, st.kclass, st.flops_per_row * (double)M);
        if constexpr (std::is_same<T, float>::value) {
          GemmF32Params g = st.f32;
          g.epi.M = (int)M;
          g.rows_valid = (int)M;
          launch_gemm_f32(g, st.epi_kind, s);
        } else {
          launch_gemm_tc(st.tc, (int)M, sms[st.stream], opt.ablate_gemm, s);
        }
        ctx->launches++;
        break;
      }
      case ST_INCEPT_A: {
        if constexpr (!std::is_same<T, float>::value) {
          Timer tm(ctx, s, st.kclass, st.flops_per_row * (double)M);
          launch_incept_a(st.ia, (int)M, sms[st.stream], opt.ablate_a, s);
          ctx->launches++;
        }
        break;
      }
      case ST_INCEPT_B: {
        if constexpr (!std::is_same<T, float>::value) {
          Timer tm(ctx, s, st.kclass, st.flops_per_row * (double)M);
          launch_incept_b(st.ib, (int)M, sms[st.stream], opt.ablate_b, s);
          ctx->launches++;
        }
        break;
      }
      case ST_LSTM_SEQ: {
        if constexpr (!std::is_same<T, float>::value) {
          Timer tm(ctx, s, st.kclass, st.flops_per_row * (double)M);
          launch_lstm_seq2(st.seq, (int)n, sms[st.stream], opt.ablate_lstm, s);
          ctx->launches++;
        }
        break;
      }
      case ST_FRONT: {
        if constexpr (!std::is_same<T, float>::value) {
          Timer tm(ctx, s, st.kclass, st.flops_per_row * (double)M);
          launch_front2(st.front, signals, (int)n, sms[st.stream], s);
          ctx->launches++;
        }
        break;
      }
      case ST_STEM: {
        Timer tm(ctx, s, KC_ELEM, 0);
        StemParams p = st.stem;
        p.signals = signals;
        p.n = (int)n;
        const size_t smem = (size_t)(p.S + 16 + 64 * 7 + 64) * sizeof(float);
        stem_kernel<T><<<(unsigned)n, 256, smem, s>>>(p);
        DSB_CUDA(cudaGetLastError());
        ctx->launches++;
        break;
      }
      case ST_POOL1: {
        Timer tm(ctx, s, KC_ELEM, 0);
        const int64_t total = M * (st.C / 8);
        if constexpr (SPLIT) {
          // the warp-per-row pool kernel of the stage boundaries with stride 1 (6 TB/s against 3.6 for the per-vector kernel)
          const int rows_out = (int)M;
          const unsigned grid = (unsigned)std::min<int64_t>((rows_out + 7) / 8, (int64_t)ctx->sm_count * 16);
          pool3s2_16_kernel<true, true><<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(st.src), reinterpret_cast<uint16_t*>(st.dst), rows_out,
                                                             st.C / 8, st.ld, st.Lin, st.Lpin, st.Lout, st.Lpout, st.pl, st.lo_in, st.lo_out, 1);
        } else
          pool3s1_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const T*>(st.src), reinterpret_cast<T*>(st.dst), M, st.C, st.ld);
        DSB_CUDA(cudaGetLastError());
        ctx->launches++;
        break;
      }
      case ST_POOL2: {
        Timer tm(ctx, s, KC_ELEM, 0);
        const int64_t total = n * st.Lpout * (st.C / 8);
        if constexpr (TC16) {
          const int rows_out = (int)(n * st.Lpout);
          const unsigned grid = (unsigned)std::min<int64_t>((rows_out + 7) / 8, (int64_t)ctx->sm_count * 16);
          pool3s2_16_kernel<SPLIT, std::is_same<T, __half>::value><<<grid, 256, 0, s>>>(
              reinterpret_cast<const uint16_t*>(st.src), reinterpret_cast<uint16_t*>(st.dst), rows_out, st.C / 8, st.ld,
              st.Lin, st.Lpin, st.Lout, st.Lpout, st.pl, st.lo_in, st.lo_out);
        } else {
          pool3s2_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const T*>(st.src), reinterpret_cast<T*>(st.dst), n, st.C, st.ld,
                                                                            st.Lin, st.Lpin, st.Lout, st.Lpout, st.pl);
        }
        DSB_CUDA(cudaGetLastError());
        ctx->launches++;
        break;
      }
      case ST_FEAT: {
        Timer tm(ctx, s, KC_ELEM, 0);
        const int64_t total = n * fm.kmer_len;
        lstm_feat_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(kmer, means, stds, lens, reinterpret_cast<T*>(plan.F), n, plan.cap,
                                                                            fm.kmer_len, plan.FW, fm.has_base() ? 1 : 0);
        DSB_CUDA(cudaGetLastError());
        ctx->launches++;
        break;
      }
      case ST_ZERO:
        // whole 128-row tiles: the tensor-core path keeps the cell state in a tile-interleaved layout
        DSB_CUDA(cudaMemsetAsync(st.dst, 0, st.zero_bytes_per_sample * (size_t)((n + 127) / 128 * 128), s));
        break;
      case ST_HEAD: {
        if (multi) {
          DSB_CUDA(cudaEventRecord(plan.ev_join[0], plan.side[0]));
          DSB_CUDA(cudaEventRecord(plan.ev_join[1], plan.side[1]));
          DSB_CUDA(cudaStreamWaitEvent(main_stream, plan.ev_join[0], 0));
          DSB_CUDA(cudaStreamWaitEvent(main_stream, plan.ev_join[1], 0));
        }
        Timer tm(ctx, s, KC_ELEM, 0);
        HeadParams p = st.head;
        p.probs = probs;
        p.pred = pred;
        p.n = n;
        const size_t hsm = ((p.y ? (size_t)p.L * CH * 2 : 0) + 4 * HID) * sizeof(float);
        static bool head_cfg[64] = {false};
        if (!head_cfg[ctx->device & 63]) {
          DSB_CUDA(cudaFuncSetAttribute(head_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
          head_cfg[ctx->device & 63] = true;
        }
        const unsigned hgrid = (unsigned)std::min<int64_t>((n + 31) / 32, (int64_t)ctx->sm_count * 3);      // 8 warps x 4 samples per block iteration
        head_kernel<T><<<hgrid, 256, hsm, s>>>(p);
        DSB_CUDA(cudaGetLastError());
        ctx->launches++;
        break;
      }
    }
  }
}

}  // namespace

template <class T, bool SPLIT>
size_t plan_bytes(const dsb_ctx* ctx, int64_t chunk, bool w3 = false) {
  ForwardPlan scratch;
  Builder<T, SPLIT> b{const_cast<dsb_ctx*>(ctx), scratch, ctx->model, chunk, Arena{}, Arena{}, false, nullptr};
  b.w3 = w3;
  b.build();
  return b.ba.off + b.wa.off;
}

size_t forward_workspace_bytes(const dsb_ctx* ctx, int64_t chunk, int precision) {
  if (!ctx->have_model) return 0;
  size_t bytes = 0;
  switch (precision) {
    case DSB_PREC_FP32: bytes = plan_bytes<float, false>(ctx, chunk); break;
    case DSB_PREC_FP16: bytes = plan_bytes<__half, false>(ctx, chunk); break;
    case DSB_PREC_FP16X2: bytes = plan_bytes<__half, true>(ctx, chunk); break;
    case DSB_PREC_FP16X3: bytes = plan_bytes<__half, true>(ctx, chunk, true); break;
    case DSB_PREC_BF16: bytes = plan_bytes<__nv_bfloat16, false>(ctx, chunk); break;
    default: return 0;
  }
  return bytes + 2048;
}

void forward_dev(dsb_ctx* ctx, int precision, int64_t n, const int32_t* kmer, const float* means, const float* stds,
                 const float* lens, const float* signals, float* probs, int32_t* pred, cudaStream_t stream) {
  if (!ctx->have_model) fail(DSB_ERR_STATE, "dsb_forward called before dsb_load_weights");
  if (precision < DSB_PREC_FP32 || precision > DSB_PREC_FP16X3) fail(DSB_ERR_INVALID, "unknown precision %d", precision);
  if (n < 0) fail(DSB_ERR_INVALID, "negative sample count");
  if (n == 0) return;
  if (!signals || !probs || !pred || (ctx->model.has_rnn() && (!means || !stds || !lens || (ctx->model.has_base() && !kmer))))
    fail(DSB_ERR_INVALID, "null input/output pointer");
  if (precision != DSB_PREC_FP32 && ctx->cc_major != 10)
    fail(DSB_ERR_ARCH, "the 16-bit tensor-core path needs an sm_100 device (found sm_%d%d); there is no fallback", ctx->cc_major, ctx->cc_minor);
  DSB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->nonfinite_dev) {
    DSB_CUDA(cudaMalloc(&ctx->nonfinite_dev, sizeof(int)));
    DSB_CUDA(cudaMemset(ctx->nonfinite_dev, 0, sizeof(int)));
  }
  const int64_t want = std::min<int64_t>(ctx->max_chunk, (n + 127) / 128 * 128);
  auto& slot = ctx->plan[precision];
  if (!slot || slot->cap < want || slot->cap > ctx->max_chunk) {
    slot.reset();
    std::unique_ptr<ForwardPlan, PlanDeleter> np(new ForwardPlan());
    np->precision = precision;
    DSB_CUDA(cudaStreamCreateWithFlags(&np->side[0], cudaStreamNonBlocking));
    DSB_CUDA(cudaStreamCreateWithFlags(&np->side[1], cudaStreamNonBlocking));
    DSB_CUDA(cudaEventCreateWithFlags(&np->ev_fork, cudaEventDisableTiming));
    DSB_CUDA(cudaEventCreateWithFlags(&np->ev_join[0], cudaEventDisableTiming));
    DSB_CUDA(cudaEventCreateWithFlags(&np->ev_join[1], cudaEventDisableTiming));
    switch (precision) {
      case DSB_PREC_FP32: build_plan<float, false>(ctx, *np, want, stream); break;
      case DSB_PREC_FP16: build_plan<__half, false>(ctx, *np, want, stream); break;
      case DSB_PREC_FP16X2: build_plan<__half, true>(ctx, *np, want, stream); break;
      case DSB_PREC_FP16X3: build_plan<__half, true>(ctx, *np, want, stream, true); break;
      default: build_plan<__nv_bfloat16, false>(ctx, *np, want, stream); break;
    }
    slot = std::move(np);
  }
  ForwardPlan& plan = *slot;
  const int KL = ctx->model.kmer_len, S = ctx->model.signal_len;
  for (int64_t off = 0; off < n; off += plan.cap) {
    const int64_t m = std::min<int64_t>(plan.cap, n - off);
    const int32_t* k = kmer ? kmer + off * KL : nullptr;
    const float* mu = means ? means + off * KL : nullptr;
    const float* sd = stds ? stds + off * KL : nullptr;
    const float* ln = lens ? lens + off * KL : nullptr;
    const float* sg = signals + off * S;
    float* pr = probs + off * 2;
    int32_t* pd = pred + off;
    switch (precision) {
      case DSB_PREC_FP32: run_chunk<float, false>(ctx, plan, m, k, mu, sd, ln, sg, pr, pd, stream); break;
      case DSB_PREC_FP16: run_chunk<__half, false>(ctx, plan, m, k, mu, sd, ln, sg, pr, pd, stream); break;
      case DSB_PREC_FP16X2: case DSB_PREC_FP16X3: run_chunk<__half, true>(ctx, plan, m, k, mu, sd, ln, sg, pr, pd, stream); break;
      default: run_chunk<__nv_bfloat16, false>(ctx, plan, m, k, mu, sd, ln, sg, pr, pd, stream); break;
    }
  }
}

}  // namespace dsb
