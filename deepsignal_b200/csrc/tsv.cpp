// tsv.cpp — host-side codec for the two text formats around the hot path:
//   * the 12-column feature file written by extract_features (reference extract_features.py:289-303) and parsed by
//     call_mods' reader (call_modifications.py:35-91): parsed here straight into the five model-input arrays;
//   * the 10-column call_mods result line (call_modifications.py:183-190): prob_0/(prob_0+prob_1) and
//     prob_1/(prob_0+prob_1) in float32, printed exactly like numpy's str(np.float32) (shortest round-trip digits,
//     positional for 1e-4 <= |x| < 1e16, scientific otherwise).
// Pure host code (no CUDA): the Python reader costs ~50 us per line, which caps the CLI at ~20 k samples/s.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/dsb200.h"

namespace {

inline const char* skip_ws_back(const char* b, const char* e) {
  while (e > b && (e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\n')) --e;
  return e;
}

// Parses `count` comma separated numbers from [p, e) with from_chars; returns false on any mismatch.
template <class T, class Out>
bool parse_list(const char* p, const char* e, int count, Out* out) {
  for (int i = 0; i < count; ++i) {
    T v{};
    auto r = std::from_chars(p, e, v);
    if (r.ec != std::errc()) return false;
    out[i] = static_cast<Out>(v);
    p = r.ptr;
    if (i + 1 < count) {
      if (p >= e || *p != ',') return false;
      ++p;
    }
  }
  return p == e;
}

struct LineJob {
  const char* buf;
  const int64_t* line_start;   // [n]
  const int64_t* line_end;     // [n] (trailing whitespace stripped)
  int64_t lo, hi;
  int K, S;
  int32_t* kmer; float* means; float* stds; float* lens; float* signals; int32_t* labels; int64_t* info_off;
  int64_t bad = -1;
};

void parse_range(LineJob& j) {
  for (int64_t i = j.lo; i < j.hi; ++i) {
    const char* b = j.buf + j.line_start[i];
    const char* e = j.buf + j.line_end[i];
    while (b < e && (*b == ' ' || *b == '\r')) ++b;          // str.strip() of the reference
    const char* col[13];
    int nc = 0;
    col[nc++] = b;
    for (const char* p = b; p < e && nc < 13; ++p)
      if (*p == '\t') col[nc++] = p + 1;
    if (nc != 12) { j.bad = i; return; }
    col[12] = e + 1;
    auto cb = [&](int c) { return col[c]; };
    auto ce = [&](int c) { return col[c + 1] - 1; };
    j.info_off[2 * i] = cb(0) - j.buf;
    j.info_off[2 * i + 1] = ce(5) - j.buf;
    if (ce(6) - cb(6) != j.K) { j.bad = i; return; }
    for (int k = 0; k < j.K; ++k) {
      int code;
      switch (cb(6)[k]) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break;
                          case 'T': code = 3; break; case 'N': code = 4; break; default: j.bad = i; return; }
      j.kmer[i * j.K + k] = code;
    }
    // float(x) then the float32 cast of the feed (call_modifications.py:53-56, model.py:32-36)
    if (!parse_list<double>(cb(7), ce(7), j.K, j.means + i * j.K) || !parse_list<double>(cb(8), ce(8), j.K, j.stds + i * j.K) ||
        !parse_list<long long>(cb(9), ce(9), j.K, j.lens + i * j.K) || !parse_list<double>(cb(10), ce(10), j.S, j.signals + i * j.S)) {
      j.bad = i;
      return;
    }
    long long lab = 0;
    auto r = std::from_chars(cb(11), ce(11), lab);
    if (r.ec != std::errc() || r.ptr != ce(11)) { j.bad = i; return; }
    j.labels[i] = (int32_t)lab;
  }
}

// str(np.float32(x)): shortest round-trip digits; positional when 1e-4 <= |x| < 1e16 (always with a fractional
// digit), otherwise d[.ddd]e±XX.
char* format_np_float32(char* p, char* end, float x) {
  if (std::isnan(x)) { memcpy(p, "nan", 3); return p + 3; }
  if (std::isinf(x)) { if (x < 0) *p++ = '-'; memcpy(p, "inf", 3); return p + 3; }
  if (x == 0.0f) { if (std::signbit(x)) *p++ = '-'; memcpy(p, "0.0", 3); return p + 3; }
  // numpy (scalartypes.c.src) compares the value widened to double: positional iff 1e-4 <= |x| < 1e16
  const double ax = std::fabs((double)x);
  char tmp[48];
  auto rs = std::to_chars(tmp, tmp + sizeof tmp, x, std::chars_format::scientific);
  if (ax >= 1e-4 && ax < 1e16) {
    auto r = std::to_chars(p, end, x, std::chars_format::fixed);
    bool has_dot = false;
    for (char* q = p; q < r.ptr; ++q) has_dot |= (*q == '.');
    char* e = r.ptr;
    if (!has_dot) { *e++ = '.'; *e++ = '0'; }
    return e;
  }
  memcpy(p, tmp, (size_t)(rs.ptr - tmp));
  return p + (rs.ptr - tmp);
}

}  // namespace

extern "C" {

int64_t dsb_tsv_parse_features(const char* buf, int64_t len, int32_t kmer_len, int32_t signal_len, int64_t max_samples,
                               int32_t* kmer, float* means, float* stds, float* lens, float* signals, int32_t* labels,
                               int64_t* info_off, int64_t* consumed, int32_t nthreads) {
  if (!buf || len < 0 || kmer_len < 1 || signal_len < 1 || !kmer || !means || !stds || !lens || !signals || !labels || !info_off)
    return DSB_ERR_INVALID;
  std::vector<int64_t> ls, le;
  ls.reserve(1024);
  le.reserve(1024);
  int64_t pos = 0;
  while (pos < len && (int64_t)ls.size() < max_samples) {
    const char* nl = static_cast<const char*>(memchr(buf + pos, '\n', (size_t)(len - pos)));
    if (!nl) break;                                   // incomplete last line: left for the next block
    const char* e = skip_ws_back(buf + pos, nl);
    if (e > buf + pos) { ls.push_back(pos); le.push_back(e - buf); }   // blank lines are skipped
    pos = (nl - buf) + 1;
  }
  const int64_t n = (int64_t)ls.size();
  if (consumed) *consumed = pos;
  if (n == 0) return 0;
  int nt = nthreads < 1 ? 1 : nthreads;
  if (nt > 64) nt = 64;
  if (n < 256) nt = 1;
  std::vector<LineJob> jobs(nt);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    jobs[t] = LineJob{buf, ls.data(), le.data(), n * t / nt, n * (t + 1) / nt, kmer_len, signal_len, kmer, means, stds, lens, signals, labels, info_off};
    if (nt > 1) th.emplace_back(parse_range, std::ref(jobs[t]));
  }
  if (nt == 1) parse_range(jobs[0]);
  for (auto& x : th) x.join();
  for (auto& j : jobs)
    if (j.bad >= 0) return -(1000 + j.bad);            // -(1000 + index of the first malformed line of a worker)
  return n;
}

int64_t dsb_tsv_format_calls(const char* info_buf, const int64_t* info_off, const int32_t* kmer, int32_t kmer_len,
                             const float* probs, const int32_t* pred, int64_t n, char* out, int64_t out_cap) {
  if (!info_buf || !info_off || !kmer || !probs || !pred || !out) return DSB_ERR_INVALID;
  static const char code2base[5] = {'A', 'C', 'G', 'T', 'N'};
  char* p = out;
  char* end = out + out_cap;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t ilen = info_off[2 * i + 1] - info_off[2 * i];
    if (end - p < ilen + kmer_len + 96) return DSB_ERR_CAPACITY;
    memcpy(p, info_buf + info_off[2 * i], (size_t)ilen);
    p += ilen;
    const float p0 = probs[2 * i], p1 = probs[2 * i + 1];
    const float s = p0 + p1;                           // float32 arithmetic, as np.float32 scalars in the reference
    *p++ = '\t';
    p = format_np_float32(p, end, p0 / s);
    *p++ = '\t';
    p = format_np_float32(p, end, p1 / s);
    *p++ = '\t';
    auto r = std::to_chars(p, end, pred[i]);
    p = r.ptr;
    *p++ = '\t';
    for (int k = 0; k < kmer_len; ++k) {
      const int c = kmer[i * kmer_len + k];
      *p++ = (c >= 0 && c < 5) ? code2base[c] : '?';
    }
    *p++ = '\n';
  }
  return p - out;
}

}  // extern "C"
