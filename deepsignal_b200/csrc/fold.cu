// fold.cu — checkpoint tensors (TF-1.x names) -> folded GEMM operands, in float64 on the host.
//
// Exact algebraic folds: BatchNorm(inference) into the preceding bias-free conv
// (reference layers.py:80-84), embedding·K_x into five LUT rows per direction (model.py:61-69),
// dense·dense_1 into one [joint,2] matrix (layers.py:257-264; no bias / activation / dropout at
// inference), and the 1x7 SAME average pool (valid-count divisor, layers.py:233-235) into the CNN
// rows of that matrix.  LSTM kernel columns are re-ordered from TF's [i|j|f|o] blocks to 4*u+g so
// that the four gates of a hidden unit are adjacent in the GEMM output.
#include <cmath>

#include "common.cuh"

namespace dsb {
namespace {

struct TensorMap {
  std::map<std::string, const dsb_tensor*> m;
  const dsb_tensor& get(const std::string& name, std::initializer_list<int64_t> shape) const {
    auto it = m.find(name);
    if (it == m.end()) fail(DSB_ERR_WEIGHTS, "checkpoint tensor '%s' is missing", name.c_str());
    const dsb_tensor& t = *it->second;
    if (t.ndim != (int)shape.size()) fail(DSB_ERR_WEIGHTS, "tensor '%s': rank %d, expected %zu", name.c_str(), t.ndim, shape.size());
    int i = 0;
    for (int64_t s : shape) {
      if (t.shape[i] != s) fail(DSB_ERR_WEIGHTS, "tensor '%s': dim %d is %lld, expected %lld", name.c_str(), i, (long long)t.shape[i], (long long)s);
      ++i;
    }
    if (!t.data) fail(DSB_ERR_WEIGHTS, "tensor '%s' has no data", name.c_str());
    return t;
  }
};

// conv kernel HWIO [1,k,cin,cout] + BN(scope) -> rows n of [cout][k*cin] (tap-major) and bias.
void fold_conv(const TensorMap& tm, const std::string& conv, const std::string& bn, int k, int cin, int cout,
               std::vector<float>& w, std::vector<float>& bias, int row0, int K) {
  const float* kern = tm.get(conv + "/kernel", {1, k, cin, cout}).data;
  const float* g = tm.get(bn + "/gamma", {cout}).data;
  const float* b = tm.get(bn + "/beta", {cout}).data;
  const float* mu = tm.get(bn + "/moving_mean", {cout}).data;
  const float* var = tm.get(bn + "/moving_variance", {cout}).data;
  for (int n = 0; n < cout; ++n) {
    double scale = (double)g[n] / std::sqrt((double)var[n] + 1e-3);
    for (int t = 0; t < k; ++t)
      for (int c = 0; c < cin; ++c)
        w[(size_t)(row0 + n) * K + t * cin + c] = (float)((double)kern[((size_t)t * cin + c) * cout + n] * scale);
    bias[row0 + n] = (float)((double)b[n] - (double)mu[n] * scale);
  }
}

void make_conv(const TensorMap& tm, const std::string& scope, const std::string& conv, const std::string& bn,
               int k, int cin, int cout, FoldedGemm& out) {
  out.N = cout;
  out.K = k * cin;
  out.w.assign((size_t)out.N * out.K, 0.f);
  out.bias.assign(out.N, 0.f);
  fold_conv(tm, scope + conv, scope + bn, k, cin, cout, out.w, out.bias, 0, out.K);
}

}  // namespace

void fold_model(const dsb_tensor* tensors, int n, uint32_t flags, int kmer_len, int signal_len, FoldedModel& fm) {
  if (!(flags & (DSB_FLAG_CNN | DSB_FLAG_RNN))) fail(DSB_ERR_INVALID, "at least one of is_cnn/is_rnn should be True");
  if (kmer_len < 1 || kmer_len > 64 || kmer_len % 2 == 0) fail(DSB_ERR_INVALID, "kmer_len must be odd and in [1,63]");
  if (signal_len < 8 || signal_len > 4096) fail(DSB_ERR_INVALID, "signal_len out of range");
  TensorMap tm;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].name) fail(DSB_ERR_INVALID, "tensor %d has no name", i);
    tm.m[tensors[i].name] = &tensors[i];
  }
  fm = FoldedModel();
  fm.flags = flags;
  fm.kmer_len = kmer_len;
  fm.signal_len = signal_len;
  fm.L1 = (signal_len + 1) / 2;
  fm.L2 = (fm.L1 + 1) / 2;
  fm.L3 = (fm.L2 + 1) / 2;
  fm.L4 = (fm.L3 + 1) / 2;
  const std::string pfx = "model";

  if (fm.has_cnn()) {
    const std::string s = pfx + "signalm";
    make_conv(tm, s + "conv_layer1/", "conv", "bn", 7, 1, 64, fm.conv1);
    make_conv(tm, s + "conv_layer2/", "conv", "bn", 1, 64, 128, fm.conv2);
    make_conv(tm, s + "conv_layer3/", "conv", "bn", 3, 128, 256, fm.conv3);
    for (int i = 1; i <= NBLOCK; ++i) {
      FoldedBlock& fb = fm.blocks[i - 1];
      fb.cin = (i == 1) ? 256 : CH;
      const std::string sc = s + "incp_layer" + std::to_string(i) + "/" + s + std::to_string(i);
      make_conv(tm, sc + "branch1_maxpooling/", "conv1a_1x1", "bn", 1, fb.cin, 48, fb.b1);
      fb.p1.N = 192;
      fb.p1.K = fb.cin;
      fb.p1.w.assign((size_t)192 * fb.cin, 0.f);
      fb.p1.bias.assign(192, 0.f);
      fold_conv(tm, sc + "branch2_1x1/conv0b_1x1", sc + "branch2_1x1/bn", 1, fb.cin, 48, fb.p1.w, fb.p1.bias, 0, fb.cin);
      fold_conv(tm, sc + "branch3_1x3/conv0c_1x1", sc + "branch3_1x3/bn1", 1, fb.cin, 32, fb.p1.w, fb.p1.bias, 48, fb.cin);
      fold_conv(tm, sc + "branch4_1x5/conv0d_1x1", sc + "branch4_1x5/bn1", 1, fb.cin, 32, fb.p1.w, fb.p1.bias, 80, fb.cin);
      fold_conv(tm, sc + "branch5_residual_1x3/conv0e_1x1", sc + "branch5_residual_1x3/bn1", 1, fb.cin, 32, fb.p1.w, fb.p1.bias, 112, fb.cin);
      fold_conv(tm, sc + "branch5_residual_1x3/convstem_1x1", sc + "branch5_residual_1x3/bn0", 1, fb.cin, 48, fb.p1.w, fb.p1.bias, 144, fb.cin);
      make_conv(tm, sc + "branch3_1x3/", "conv1c_1x3", "bn2", 3, 32, 48, fb.b3);
      make_conv(tm, sc + "branch4_1x5/", "conv1d_1x5", "bn2", 5, 32, 48, fb.b4);
      make_conv(tm, sc + "branch5_residual_1x3/", "conv1e_1x3", "bn2", 3, 32, 64, fb.b5b);
      make_conv(tm, sc + "branch5_residual_1x3/", "conv2e_1x1", "bn3", 1, 64, 48, fb.b5c);
    }
  }

  if (fm.has_rnn()) {
    const int in0 = fm.has_base() ? 131 : 3;
    const float* emb = fm.has_base() ? tm.get(pfx + "embedding", {1024, 128}).data : nullptr;
    const char* dirs[2] = {"fw", "bw"};
    for (int d = 0; d < 2; ++d) {
      for (int l = 0; l < NLAYER; ++l) {
        const int cin = (l == 0) ? in0 : HID;
        const std::string base = pfx + "em/" + dirs[d] + "/multi_rnn_cell/cell_" + std::to_string(l) + "/lstm_cell/";
        const float* K = tm.get(base + "kernel", {cin + HID, GATES}).data;
        const float* b = tm.get(base + "bias", {GATES}).data;
        FoldedGemm& g = fm.lstm[d][l].g;
        g.N = GATES;
        g.K = (l == 0) ? HID + 8 : 2 * HID;
        g.w.assign((size_t)g.N * g.K, 0.f);
        g.bias.assign(g.N, 0.f);
        for (int gate = 0; gate < 4; ++gate) {
          for (int u = 0; u < HID; ++u) {
            const int oc = gate * HID + u;   // TF column
            const int nc = 4 * u + gate;     // our column
            g.bias[nc] = b[oc];
            float* row = &g.w[(size_t)nc * g.K];
            if (l == 0) {
              for (int k = 0; k < HID; ++k) row[k] = K[(size_t)(cin + k) * GATES + oc];   // h part first
              const int sc0 = fm.has_base() ? 128 : 0;                                    // rows of mean/std/len
              if (fm.has_base()) {
                for (int code = 0; code < 5; ++code) {
                  double acc = 0.0;
                  for (int e = 0; e < 128; ++e) acc += (double)emb[code * 128 + e] * (double)K[(size_t)e * GATES + oc];
                  row[HID + code] = (float)acc;
                }
              }
              for (int j = 0; j < 3; ++j) row[HID + 5 + j] = K[(size_t)(sc0 + j) * GATES + oc];
            } else {
              for (int k = 0; k < 2 * HID; ++k) row[k] = K[(size_t)k * GATES + oc];       // x rows then h rows
            }
          }
        }
      }
    }
  }

  // head: W = dense/kernel · dense_1/kernel in float64, then split and fold the average pool.
  const int n_rnn = fm.has_rnn() ? 2 * HID : 0;
  const int n_cnn = fm.has_cnn() ? fm.L4 * CH : 0;
  const int J = n_rnn + n_cnn;
  const float* W1 = tm.get("dense/kernel", {J, J}).data;
  const float* W2 = tm.get("dense_1/kernel", {J, 2}).data;
  std::vector<double> Wh((size_t)J * 2, 0.0);
  for (int r = 0; r < J; ++r) {
    double a0 = 0.0, a1 = 0.0;
    const float* w1r = W1 + (size_t)r * J;
    for (int k = 0; k < J; ++k) {
      a0 += (double)w1r[k] * (double)W2[2 * k];
      a1 += (double)w1r[k] * (double)W2[2 * k + 1];
    }
    Wh[2 * r] = a0;
    Wh[2 * r + 1] = a1;
  }
  fm.head_rnn.assign((size_t)n_rnn * 2, 0.f);
  for (int r = 0; r < n_rnn * 2; ++r) fm.head_rnn[r] = (float)Wh[r];
  fm.head_cnn.assign((size_t)n_cnn * 2, 0.f);
  if (fm.has_cnn()) {
    const int L = fm.L4;
    // flat[p'] = (1/cnt[p']) * sum_{p in [p'-3, p'+3] valid} y[p]  ->  weight of y[p] = sum_{p'} Wh[p']/cnt[p']
    for (int p = 0; p < L; ++p) {
      for (int c = 0; c < CH; ++c) {
        double a0 = 0.0, a1 = 0.0;
        for (int q = p - 3; q <= p + 3; ++q) {
          if (q < 0 || q >= L) continue;
          const int lo = q - 3 < 0 ? 0 : q - 3, hi = q + 3 >= L ? L - 1 : q + 3;
          const double inv = 1.0 / (double)(hi - lo + 1);
          const size_t r = (size_t)n_rnn + (size_t)q * CH + c;
          a0 += Wh[2 * r] * inv;
          a1 += Wh[2 * r + 1] * inv;
        }
        fm.head_cnn[((size_t)p * CH + c) * 2] = (float)a0;
        fm.head_cnn[((size_t)p * CH + c) * 2 + 1] = (float)a1;
      }
    }
  }
}

}  // namespace dsb
