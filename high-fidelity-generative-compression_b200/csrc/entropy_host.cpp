// Host half of the compress / decompress path (BASELINE north_star: "the sequential ANS entropy coder stays on the
// host"): probability-table quantisation and the lane-vectorised indexed rANS coder, bit-compatible with the
// reference's Python implementation so that bitstreams and .hfc files interchange.
//
//   hfc_pmf_to_quantized_cdf_host  <- src/helpers/maths.py:5-73 (the reference's own "TODO: port to C++")
//   hfc_rans_encode_host           <- src/compression/entropy_coding.py:251-476 + src/compression/ans.py:50-76,100-107
//   hfc_rans_decode_host           <- src/compression/entropy_coding.py:555-676 + src/compression/ans.py:78-113
//
// The coder keeps one 64-bit rANS state per lane and walks `steps` symbols per lane; 32-bit words spilled by one
// push instruction are grouped in lane order, instruction groups appear in decode order.  Out-of-table symbols are
// escaped as the reference's vectorised coder does it (including its quirk of transmitting only the lowest nibble of
// the escape value, entropy_coding.py:374-386 / :639-651), because that defines the wire format.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "hfc_internal.h"

namespace {

constexpr uint64_t kRansL = 1ull << 31;
constexpr int kOverflowWidth = 4;
constexpr int kMaxNibble = (1 << kOverflowWidth) - 1;

struct Tables {
  const int32_t* cdf;
  int32_t rows, cols;
  const int32_t* length;
  const int32_t* offset;
};

inline int nibble_count(int64_t v) {
  int w = 0;
  while ((v >> (w * kOverflowWidth)) != 0) ++w;
  return w;
}

inline bool rans_push(uint64_t& head, uint32_t start, uint32_t freq, int precision, std::vector<uint32_t>& rev) {
  if (freq == 0) return false;
  const uint64_t x_max = ((kRansL >> precision) << 32) * static_cast<uint64_t>(freq);
  if (head >= x_max) {
    rev.push_back(static_cast<uint32_t>(head));
    head >>= 32;
  }
  head = ((head / freq) << precision) + head % freq + start;
  return true;
}

int check_tables(const Tables& t, const char* who) {
  if (!t.cdf || !t.length || !t.offset || t.rows <= 0 || t.cols < 3)
    return hfc::set_error(HFC_ERR_INVALID, "%s: 'cdf' must be (rows > 0, cols >= 3) with length / offset per row", who);
  for (int r = 0; r < t.rows; ++r)
    if (t.length[r] < 2 || t.length[r] > t.cols)
      return hfc::set_error(HFC_ERR_INVALID, "%s: cdf_length[%d] = %d outside [2, %d]", who, r, t.length[r], t.cols);
  return HFC_OK;
}

}  // namespace

extern "C" int hfc_pmf_to_quantized_cdf_host(const float* pmf_host, int32_t n, int32_t precision, int32_t* cdf_host) {
  if (!pmf_host || !cdf_host || n < 2) return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: need n >= 2");
  if (precision < 8 || precision > 30)
    return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: precision %d outside [8, 30]", precision);
  const int64_t target = 1ll << precision;
  std::vector<float> c(n + 1);
  // torch.cumsum on CPU accumulates float32 inputs in double and rounds every partial sum to float32
  double acc = 0.0;
  c[0] = 0.f;
  for (int i = 0; i < n; ++i) {
    if (!(pmf_host[i] >= 0.f)) return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: pmf[%d] is negative or NaN", i);
    acc += static_cast<double>(pmf_host[i]);
    c[i + 1] = static_cast<float>(acc);
  }
  const float total = c[n];
  if (!(total > 0.f)) return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: pmf sums to zero");
  const float scale = static_cast<float>(target);
  std::vector<int64_t> q(n + 1);
  for (int i = 0; i <= n; ++i) {
    volatile float prod = c[i] * scale;                      // two separately rounded float32 operations, as in torch
    volatile float quot = prod / total;
    q[i] = static_cast<int64_t>(std::nearbyintf(quot));      // torch.round: half to even
  }
  for (int i = 0; i < n; ++i) {
    if (q[i] != q[i + 1]) continue;
    // steal one count from the least frequent symbol that can spare it (first one on ties)
    int64_t best_freq = target + 1;
    int best = -1;
    for (int j = 0; j < n; ++j) {
      const int64_t f = q[j + 1] - q[j];
      if (f > 1 && f < best_freq) { best_freq = f; best = j; }
    }
    if (best < 0) return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: nothing to steal from");
    if (best < i) { for (int j = best + 1; j <= i; ++j) q[j] -= 1; }
    else          { for (int j = i + 1; j <= best; ++j) q[j] += 1; }
  }
  if (q[0] != 0 || q[n] != target) return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: normalisation failed");
  for (int i = 0; i < n; ++i)
    if (q[i + 1] < q[i]) return hfc::set_error(HFC_ERR_INVALID, "pmf_to_quantized_cdf: CDF is not monotonic");
  for (int i = 0; i <= n; ++i) cdf_host[i] = static_cast<int32_t>(q[i]);
  return HFC_OK;
}

extern "C" int64_t hfc_rans_encode_host(const int32_t* symbols_host, const int32_t* indices_host, int64_t steps,
                                        int64_t lanes, const int32_t* cdf_host, int32_t cdf_rows, int32_t cdf_cols,
                                        const int32_t* cdf_length_host, const int32_t* cdf_offset_host,
                                        int32_t precision, uint32_t* out_host, int64_t out_capacity) {
  const Tables t{cdf_host, cdf_rows, cdf_cols, cdf_length_host, cdf_offset_host};
  int rc = check_tables(t, "rans_encode");
  if (rc != HFC_OK) return rc;
  if (!symbols_host || !indices_host || steps <= 0 || lanes <= 0)
    return hfc::set_error(HFC_ERR_INVALID, "rans_encode: empty input");
  if (precision < 8 || precision > 24) return hfc::set_error(HFC_ERR_INVALID, "rans_encode: precision outside [8, 24]");
  std::vector<uint64_t> head(lanes, kRansL);
  std::vector<uint32_t> rev;                       // spilled words in REVERSE stream order
  rev.reserve(static_cast<size_t>(steps * lanes / 2 + 16));
  std::vector<int32_t> value(lanes), low_nibble(lanes), widths(lanes);
  std::vector<int64_t> escaped;                    // lanes of this step that carry the overflow code
  for (int64_t i = steps - 1; i >= 0; --i) {
    const int32_t* sym = symbols_host + i * lanes;
    const int32_t* idx = indices_host + i * lanes;
    escaped.clear();
    int max_w = 0;
    for (int64_t l = 0; l < lanes; ++l) {
      const int32_t r = idx[l];
      if (r < 0 || r >= t.rows) return hfc::set_error(HFC_ERR_INVALID, "rans_encode: index %d out of range", r);
      const int64_t max_value = t.length[r] - 2;
      int64_t v = static_cast<int64_t>(sym[l]) - t.offset[r];
      int64_t ov = 0;
      if (v < 0) { ov = -2 * v - 1; v = max_value; }
      else if (v >= max_value) { ov = 2 * (v - max_value); v = max_value; }
      value[l] = static_cast<int32_t>(v);
      if (v == max_value) {
        const int w = nibble_count(ov);
        if (w >= kMaxNibble)
          return hfc::set_error(HFC_ERR_UNSUPPORTED, "rans_encode: escape value %lld needs >= 15 nibbles", (long long)ov);
        widths[l] = w;
        low_nibble[l] = static_cast<int32_t>(ov & kMaxNibble);
        max_w = std::max(max_w, w);
        escaped.push_back(l);
      }
    }
    // reverse instruction order: escape nibbles (max_w rounds over every escaped lane), their widths, the symbols
    for (int k = 0; k < max_w; ++k)
      for (auto it = escaped.rbegin(); it != escaped.rend(); ++it)
        rans_push(head[*it], static_cast<uint32_t>(low_nibble[*it]), 1u, kOverflowWidth, rev);
    for (auto it = escaped.rbegin(); it != escaped.rend(); ++it)
      rans_push(head[*it], static_cast<uint32_t>(widths[*it]), 1u, kOverflowWidth, rev);
    for (int64_t l = lanes - 1; l >= 0; --l) {
      const int32_t* row = t.cdf + static_cast<int64_t>(idx[l]) * t.cols;
      const uint32_t start = static_cast<uint32_t>(row[value[l]]);
      const uint32_t freq = static_cast<uint32_t>(row[value[l] + 1]) - start;
      if (!rans_push(head[l], start, freq, precision, rev))
        return hfc::set_error(HFC_ERR_INVALID, "rans_encode: zero-frequency symbol (row %d, value %d)", idx[l], value[l]);
    }
  }
  const int64_t total = 2 * lanes + static_cast<int64_t>(rev.size());
  if (!out_host || out_capacity < total) {
    hfc::set_error(HFC_ERR_INVALID, "rans_encode: output buffer holds %lld words, message needs %lld",
                   (long long)out_capacity, (long long)total);
    return out_host ? static_cast<int64_t>(HFC_ERR_INVALID) : total;   // NULL buffer = size query
  }
  for (int64_t l = 0; l < lanes; ++l) {
    out_host[l] = static_cast<uint32_t>(head[l] >> 32);
    out_host[lanes + l] = static_cast<uint32_t>(head[l]);
  }
  std::reverse_copy(rev.begin(), rev.end(), out_host + 2 * lanes);
  return total;
}

extern "C" int hfc_rans_decode_host(const uint32_t* encoded_host, int64_t n_words, const int32_t* indices_host,
                                    int64_t steps, int64_t lanes, const int32_t* cdf_host, int32_t cdf_rows,
                                    int32_t cdf_cols, const int32_t* cdf_length_host, const int32_t* cdf_offset_host,
                                    int32_t precision, int32_t* symbols_host) {
  const Tables t{cdf_host, cdf_rows, cdf_cols, cdf_length_host, cdf_offset_host};
  int rc = check_tables(t, "rans_decode");
  if (rc != HFC_OK) return rc;
  if (!encoded_host || !indices_host || !symbols_host || steps <= 0 || lanes <= 0)
    return hfc::set_error(HFC_ERR_INVALID, "rans_decode: empty input");
  if (precision < 8 || precision > 24) return hfc::set_error(HFC_ERR_INVALID, "rans_decode: precision outside [8, 24]");
  if (n_words < 2 * lanes) return hfc::set_error(HFC_ERR_INVALID, "rans_decode: message shorter than its %lld lane states", (long long)lanes);
  std::vector<uint64_t> head(lanes);
  for (int64_t l = 0; l < lanes; ++l)
    head[l] = (static_cast<uint64_t>(encoded_host[l]) << 32) | encoded_host[lanes + l];
  int64_t pos = 2 * lanes;
  bool exhausted = false;
  auto refill = [&](uint64_t& h) {
    if (h < kRansL) {
      if (pos >= n_words) { exhausted = true; return; }
      h = (h << 32) | encoded_host[pos++];
    }
  };
  auto pop4 = [&](uint64_t& h) -> uint32_t {
    const uint32_t v = static_cast<uint32_t>(h & kMaxNibble);
    h >>= kOverflowWidth;                       // freq 1, start == v
    refill(h);
    return v;
  };
  const uint64_t mask = (1ull << precision) - 1;
  std::vector<int64_t> escaped;
  std::vector<uint32_t> widths, ov;
  for (int64_t i = 0; i < steps; ++i) {
    const int32_t* idx = indices_host + i * lanes;
    int32_t* out = symbols_host + i * lanes;
    escaped.clear();
    for (int64_t l = 0; l < lanes; ++l) {
      const int32_t r = idx[l];
      if (r < 0 || r >= t.rows) return hfc::set_error(HFC_ERR_INVALID, "rans_decode: index %d out of range", r);
      const int32_t* row = t.cdf + static_cast<int64_t>(r) * t.cols;
      const int32_t len = t.length[r];
      const uint32_t cf = static_cast<uint32_t>(head[l] & mask);
      // last s in [0, len) with row[s] <= cf
      const int32_t* ub = std::upper_bound(row, row + len, cf, [](uint32_t a, int32_t b) { return a < static_cast<uint32_t>(b); });
      const int32_t s = static_cast<int32_t>(ub - row) - 1;
      if (s < 0 || s > len - 2) return hfc::set_error(HFC_ERR_INVALID, "rans_decode: corrupt message (step %lld, lane %lld)", (long long)i, (long long)l);
      const uint32_t start = static_cast<uint32_t>(row[s]);
      const uint32_t freq = static_cast<uint32_t>(row[s + 1]) - start;
      head[l] = static_cast<uint64_t>(freq) * (head[l] >> precision) + cf - start;
      refill(head[l]);
      out[l] = s;
      if (s == len - 2) escaped.push_back(l);
    }
    if (!escaped.empty()) {
      widths.resize(escaped.size());
      ov.assign(escaped.size(), 0u);
      uint32_t max_w = 0;
      for (size_t e = 0; e < escaped.size(); ++e) {
        widths[e] = pop4(head[escaped[e]]);
        if (widths[e] == static_cast<uint32_t>(kMaxNibble))
          return hfc::set_error(HFC_ERR_UNSUPPORTED, "rans_decode: escape wider than 14 nibbles");
        max_w = std::max(max_w, widths[e]);
      }
      for (uint32_t k = 0; k < max_w; ++k)
        for (size_t e = 0; e < escaped.size(); ++e) {
          const uint32_t v = pop4(head[escaped[e]]);
          if (k < widths[e]) ov[e] |= v;        // every nibble lands at bit 0 (reference quirk)
        }
      for (size_t e = 0; e < escaped.size(); ++e) {
        const int64_t l = escaped[e];
        const int32_t max_value = t.length[idx[l]] - 2;
        const int32_t half = static_cast<int32_t>(ov[e] >> 1);
        out[l] = (ov[e] & 1u) ? -half - 1 : half + max_value;
      }
    }
    for (int64_t l = 0; l < lanes; ++l) out[l] += t.offset[idx[l]];
    if (exhausted) return hfc::set_error(HFC_ERR_INVALID, "rans_decode: message exhausted at step %lld", (long long)i);
  }
  return HFC_OK;
}
