// GPU half of the compress / decompress path (compress.py -> Model.compress / decompress, src/model.py:262-344;
// Hyperprior.compress_forward / decompress_forward, src/hyperprior.py:195-274): everything either side of the host
// rANS coder that touches a whole latent tensor.
//
//   quantize_symbols  : symbols = floor(x + .5 - mean), table index of every scale, Shannon bit estimate of the
//                       quantised latents, the dequantised latents -- one pass, written in the order the coder walks
//                       (prior_model.py:122-198, hyperprior_model.py:108-197)
//   dequantize_symbols: decoded symbols (+ mean) back to NCHW fp32 (prior_model.py:236-246, entropy_models.py:65-73)
//
// HBM-bound integer / float work: 12 B read + 8..12 B written per latent element.  The "pixel steps" layout (batch 1:
// lanes = channels, steps = pixels) is a (C x HW) -> (HW x C) transpose, done through 32 x 33 shared-memory tiles so
// that both the NCHW reads and the [pixel][channel] writes are coalesced 128-byte segments.
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"
#include "hfc_likelihood.cuh"

namespace hfc {

constexpr int kMaxScales = 256;

struct SymParams {
  int32_t n, c, hw;
  int32_t n_scales;
  float lb;
  int32_t type;
  int32_t sorted;      // scale table is non-decreasing: binary search instead of the linear count
};

struct SymOut {
  int32_t sym, idx;
  float deq, loglik;
};

// One latent element.  `table` is the scale table in shared memory.
__device__ __forceinline__ SymOut sym_one(float x, float mu, float sraw, bool has_x, bool has_mean, bool has_scale,
                                          int ch, const float* table, const SymParams& p, bool want_bits) {
  SymOut o;
  o.sym = 0; o.idx = ch; o.deq = 0.f; o.loglik = 0.f;
  float scale = 1.f;
  if (has_scale) {
    // LowerBoundToward forward = clamp(min) (maths.py:92-96); NaN propagates as in torch.clamp
    scale = (sraw >= p.lb || sraw != sraw) ? sraw : p.lb;
    // compute_indices (prior_model.py:148-156): (n_scales - 1) - #{ s in table[:-1] : scale <= s }
    //   = #{ s in table[:-1] : s < scale } (NaN: n_scales - 1).
    if (p.sorted) {
      // non-decreasing table (the reference's log-spaced one; checked on the host): the count is the lower bound of
      // `scale`, found in ceil(log2(n_scales)) probes instead of the reference's n_scales - 1 full-tensor compares.
      // (first ncu capture of the linear count at the c5 size: 113 M warp instructions, 118 us, sm 83 % -- compute-bound)
      int lo = 0, hi = p.n_scales - 1;                     // answer in [lo, hi]
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (table[mid] < scale) lo = mid + 1; else hi = mid;
      }
      o.idx = (scale != scale) ? p.n_scales - 1 : lo;
    } else {
      int below = p.n_scales - 1;
      for (int k = 0; k < p.n_scales - 1; ++k) below -= (scale <= table[k]) ? 1 : 0;
      o.idx = below;
    }
  }
  if (has_x) {
    const float m = has_mean ? mu : 0.f;
    // torch.floor(bottleneck + 0.5 - means) (prior_model.py:183) / torch.floor(bottleneck + 0.5) (hyperprior_model.py:165)
    const float s = has_mean ? floorf((x + 0.5f) - m) : floorf(x + 0.5f);
    o.sym = static_cast<int32_t>(s);
    o.deq = has_mean ? static_cast<float>(o.sym) + m : static_cast<float>(o.sym);   // entropy_models.py:65-73
    if (want_bits && has_scale) {
      // quantize_st(x, offsets = means) then PriorDensity.likelihood (entropy_models.py:49-63, prior_model.py:292-305)
      const float v = x - m;
      const float q = (v + (floorf(v + 0.5f) - v)) + m;
      const float d = fabsf(q - m);
      const float inv = __fdividef(1.f, scale);
      const float pr = std_cdf((0.5f - d) * inv, p.type) - std_cdf(-(0.5f + d) * inv, p.type);
      o.loglik = __logf(fmaxf(pr, 1e-9f) + 1e-9f);
    }
  }
  return o;
}

__device__ __forceinline__ void block_accumulate(float v, double* target) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float r = lane < nwarps ? red[lane] : 0.f;
    r = warp_sum(r);
    if (lane == 0) atomicAdd(target, static_cast<double>(r));
  }
}

// ---- layout HFC_SYM_BATCH_STEPS: coder order == NCHW order ---------------------------------------------------------
__global__ void __launch_bounds__(256)
quantize_symbols_flat_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ scale,
                             const float* __restrict__ scale_table, const __grid_constant__ SymParams p,
                             int32_t* __restrict__ symbols, int32_t* __restrict__ indices, float* __restrict__ dequant,
                             double* __restrict__ bits_sum) {
  __shared__ float table[kMaxScales];
  for (int k = threadIdx.x; k < p.n_scales; k += blockDim.x) table[k] = scale ? scale_table[k] : 0.f;
  __syncthreads();
  const int64_t count = static_cast<int64_t>(p.n) * p.c * p.hw;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool want_bits = bits_sum != nullptr;
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    // the channel is only the table row of the hyper-latents (no scales); the latent path skips the two 64-bit
    // divisions (second ncu capture: 280 issued instructions per element, most of them this emulated division)
    const int ch = scale ? 0 : static_cast<int>((i / p.hw) % p.c);
    const SymOut o = sym_one(x ? x[i] : 0.f, mean ? mean[i] : 0.f, scale ? scale[i] : 0.f, x != nullptr,
                             mean != nullptr, scale != nullptr, ch, table, p, want_bits);
    if (symbols) symbols[i] = o.sym;
    if (indices) indices[i] = o.idx;
    if (dequant) dequant[i] = o.deq;
    acc += o.loglik;
  }
  if (want_bits) block_accumulate(acc, bits_sum);
}

// ---- layout HFC_SYM_PIXEL_STEPS: out[(img * hw + pix) * c + ch] ----------------------------------------------------
// grid (ceil(hw / 32), ceil(c / 32), n), block (32, 8)
__global__ void __launch_bounds__(256)
quantize_symbols_transposed_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                   const float* __restrict__ scale, const float* __restrict__ scale_table,
                                   const __grid_constant__ SymParams p, int32_t* __restrict__ symbols,
                                   int32_t* __restrict__ indices, float* __restrict__ dequant,
                                   double* __restrict__ bits_sum) {
  __shared__ float table[kMaxScales];
  __shared__ int32_t tile_sym[32][33];
  __shared__ int32_t tile_idx[32][33];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int k = tid; k < p.n_scales; k += 256) table[k] = scale ? scale_table[k] : 0.f;
  __syncthreads();
  const int img = blockIdx.z;
  const int pix0 = blockIdx.x * 32, ch0 = blockIdx.y * 32;
  const bool want_bits = bits_sum != nullptr;
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int chl = threadIdx.y + 8 * r;            // channel within the tile
    const int ch = ch0 + chl, pix = pix0 + threadIdx.x;
    if (ch < p.c && pix < p.hw) {
      const int64_t i = (static_cast<int64_t>(img) * p.c + ch) * p.hw + pix;
      const SymOut o = sym_one(x ? x[i] : 0.f, mean ? mean[i] : 0.f, scale ? scale[i] : 0.f, x != nullptr,
                               mean != nullptr, scale != nullptr, ch, table, p, want_bits);
      tile_sym[chl][threadIdx.x] = o.sym;
      tile_idx[chl][threadIdx.x] = o.idx;
      if (dequant) dequant[i] = o.deq;
      acc += o.loglik;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pl = threadIdx.y + 8 * r;             // pixel within the tile
    const int pix = pix0 + pl, ch = ch0 + threadIdx.x;
    if (ch < p.c && pix < p.hw) {
      const int64_t o = (static_cast<int64_t>(img) * p.hw + pix) * p.c + ch;
      if (symbols) symbols[o] = tile_sym[threadIdx.x][pl];
      if (indices) indices[o] = tile_idx[threadIdx.x][pl];
    }
  }
  if (want_bits) block_accumulate(acc, bits_sum);
}

__global__ void __launch_bounds__(256)
dequantize_symbols_flat_kernel(const int32_t* __restrict__ symbols, const float* __restrict__ mean, int64_t count,
                               float* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float s = static_cast<float>(symbols[i]);
    out[i] = mean ? s + mean[i] : s;
  }
}

// grid (ceil(hw / 32), ceil(c / 32), n), block (32, 8): symbols[(img * hw + pix) * c + ch] -> out[(img * c + ch) * hw + pix]
__global__ void __launch_bounds__(256)
dequantize_symbols_transposed_kernel(const int32_t* __restrict__ symbols, const float* __restrict__ mean, int32_t c,
                                     int32_t hw, float* __restrict__ out) {
  __shared__ int32_t tile[32][33];
  const int img = blockIdx.z;
  const int pix0 = blockIdx.x * 32, ch0 = blockIdx.y * 32;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pl = threadIdx.y + 8 * r;
    const int pix = pix0 + pl, ch = ch0 + threadIdx.x;
    if (ch < c && pix < hw) tile[pl][threadIdx.x] = symbols[(static_cast<int64_t>(img) * hw + pix) * c + ch];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int chl = threadIdx.y + 8 * r;
    const int ch = ch0 + chl, pix = pix0 + threadIdx.x;
    if (ch < c && pix < hw) {
      const int64_t i = (static_cast<int64_t>(img) * c + ch) * hw + pix;
      const float s = static_cast<float>(tile[threadIdx.x][chl]);
      out[i] = mean ? s + mean[i] : s;
    }
  }
}

static int flat_blocks(int64_t count, int sms) {
  const int64_t want = (count + 255) / 256;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(sms) * 16)));
}

static int launch_quantize(const float* x, const float* mean, const float* scale_raw, int32_t n, int32_t c, int32_t hw,
                           const float* scale_table, int32_t n_scales, float lb, int32_t likelihood_type, int32_t layout,
                           int32_t* symbols, int32_t* indices, float* dequant, double* bits_sum, void* stream,
                           int32_t table_sorted, const char* who) {
  if (n <= 0 || c <= 0 || hw <= 0) return set_error(HFC_ERR_INVALID, "%s: empty tensor", who);
  if (!x && !scale_raw) return set_error(HFC_ERR_INVALID, "%s: neither values nor scales given", who);
  if (scale_raw && (!scale_table || n_scales < 2 || n_scales > kMaxScales))
    return set_error(HFC_ERR_INVALID, "%s: scale table of 2..%d entries required with scales", who, kMaxScales);
  if (mean && !x) return set_error(HFC_ERR_INVALID, "%s: mean without values", who);
  if ((symbols || dequant) && !x) return set_error(HFC_ERR_INVALID, "%s: symbols / dequant need the values", who);
  if (bits_sum && (!x || !scale_raw)) return set_error(HFC_ERR_INVALID, "%s: the bit estimate needs values and scales", who);
  if (likelihood_type != 0 && likelihood_type != 1) return set_error(HFC_ERR_INVALID, "%s: likelihood_type must be 0 or 1", who);
  if (layout != HFC_SYM_BATCH_STEPS && layout != HFC_SYM_PIXEL_STEPS) return set_error(HFC_ERR_INVALID, "%s: unknown layout", who);
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  SymParams p;
  p.n = n; p.c = c; p.hw = hw; p.n_scales = scale_raw ? n_scales : 0; p.lb = lb; p.type = likelihood_type;
  p.sorted = table_sorted;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == HFC_SYM_BATCH_STEPS) {
    const int64_t count = static_cast<int64_t>(n) * c * hw;
    quantize_symbols_flat_kernel<<<flat_blocks(count, sms), 256, 0, st>>>(x, mean, scale_raw, scale_table, p, symbols,
                                                                          indices, dequant, bits_sum);
  } else {
    if (n > 65535 || (c + 31) / 32 > 65535) return set_error(HFC_ERR_INVALID, "%s: tensor too large for the tiled layout", who);
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    quantize_symbols_transposed_kernel<<<grid, block, 0, st>>>(x, mean, scale_raw, scale_table, p, symbols, indices,
                                                               dequant, bits_sum);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "%s launch: %s", who, cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

}  // namespace hfc

using namespace hfc;

extern "C" int hfc_quantize_symbols(const float* x, const float* mean, const float* scale_raw, int32_t n, int32_t c,
                                    int32_t hw, const float* scale_table, int32_t n_scales, float scale_lower_bound,
                                    int32_t likelihood_type, int32_t layout, int32_t table_sorted, int32_t* symbols,
                                    int32_t* indices, float* dequant, double* bits_sum, void* stream) {
  if (!x || (!symbols && !bits_sum))
    return set_error(HFC_ERR_INVALID, "quantize_symbols: values and a symbol buffer (or a bit-sum target) required");
  return launch_quantize(x, mean, scale_raw, n, c, hw, scale_table, n_scales, scale_lower_bound, likelihood_type, layout,
                         symbols, indices, dequant, bits_sum, stream, table_sorted, "quantize_symbols");
}

extern "C" int hfc_scale_indices(const float* scale_raw, int32_t n, int32_t c, int32_t hw, const float* scale_table,
                                 int32_t n_scales, float scale_lower_bound, int32_t layout, int32_t table_sorted,
                                 int32_t* indices, void* stream) {
  if (!scale_raw || !indices) return set_error(HFC_ERR_INVALID, "scale_indices: scales and index buffer required");
  return launch_quantize(nullptr, nullptr, scale_raw, n, c, hw, scale_table, n_scales, scale_lower_bound, 0, layout,
                         nullptr, indices, nullptr, nullptr, stream, table_sorted, "scale_indices");
}

extern "C" int hfc_dequantize_symbols(const int32_t* symbols, const float* mean, int32_t n, int32_t c, int32_t hw,
                                      int32_t layout, float* out, void* stream) {
  if (!symbols || !out || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "dequantize_symbols: null pointer or empty tensor");
  if (layout != HFC_SYM_BATCH_STEPS && layout != HFC_SYM_PIXEL_STEPS)
    return set_error(HFC_ERR_INVALID, "dequantize_symbols: unknown layout");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == HFC_SYM_BATCH_STEPS) {
    const int64_t count = static_cast<int64_t>(n) * c * hw;
    dequantize_symbols_flat_kernel<<<flat_blocks(count, sms), 256, 0, st>>>(symbols, mean, count, out);
  } else {
    if (n > 65535 || (c + 31) / 32 > 65535) return set_error(HFC_ERR_INVALID, "dequantize_symbols: tensor too large");
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    dequantize_symbols_transposed_kernel<<<grid, block, 0, st>>>(symbols, mean, c, hw, out);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "dequantize_symbols launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
