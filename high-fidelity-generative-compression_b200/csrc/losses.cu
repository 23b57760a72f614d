// Discriminator-side helpers and loss reductions of the HiFIC training step (all HBM-/latency-bound):
//   disc_input     : cat(x, nearest-upsample x16 of the context features) -> bordered NHWC fp16
//                    (src/network/discriminator.py:75-79)
//   spectral sigma : torch.nn.utils.spectral_norm power iteration (discriminator.py:46-62)
//   gan sums       : BCE-with-logits sums of the non-saturating GAN loss (src/loss/losses.py:30-41)
//   sq-diff sum    : distortion loss sum((255 x_gen - 255 x_real)^2) (src/model.py:190-194)
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

#include <cuda_fp16.h>

namespace hfc {

struct DiscInParams {
  int32_t n, h, w, cx;            // image batch / size / channels (3)
  int32_t ch, cw, cc, ccpad;      // context map size, real / padded channels
  int32_t scale;                  // nearest-neighbour upsampling factor (16)
  int32_t cpad, pt, pl, pb, pr;   // output buffer
};

// one thread per output pixel; 8-channel groups written as 16 B
__global__ void __launch_bounds__(256)
disc_input_kernel(const float* __restrict__ x, const __half* __restrict__ ctx, __half* __restrict__ out,
                  const __grid_constant__ DiscInParams p) {
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(p.n) * plane) return;
  const int ww = static_cast<int>(idx % p.w);
  const int hh = static_cast<int>((idx / p.w) % p.h);
  const int nn = static_cast<int>(idx / plane);
  int rows[3], cols[3];
  const int nr = mirror_targets(hh, p.h, p.pt, p.pb, true, rows);
  const int nc = mirror_targets(ww, p.w, p.pl, p.pr, true, cols);
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
  const __half* crow = ctx + ((static_cast<size_t>(nn) * p.ch + hh / p.scale) * p.cw + ww / p.scale) * p.ccpad;
  for (int g = 0; g < p.cpad / 8; ++g) {
    __half hv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float v = 0.f;
      if (c < p.cx) v = x[(static_cast<size_t>(nn) * p.cx + c) * plane + static_cast<size_t>(hh) * p.w + ww];
      else if (c < p.cx + p.cc) v = __half2float(crow[c - p.cx]);
      hv[j] = __float2half_rn(v);
    }
    const uint4 pk = *reinterpret_cast<const uint4*>(hv);
    for (int ri = 0; ri < nr; ++ri)
      for (int ci = 0; ci < nc; ++ci)
        *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(nn) * Hp + rows[ri]) * Wp + cols[ci]) * p.cpad + g * 8) = pk;
  }
}

// ---- spectral norm: v = normalize(W^T u); u = normalize(W v); sigma = u . (W v) ----------------
__global__ void __launch_bounds__(256) sn_wt_u_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                                      float* __restrict__ v_raw, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;   // coalesced along the columns of row-major W
  if (c >= cols) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc = fmaf(w[static_cast<size_t>(r) * cols + c], u[r], acc);
  v_raw[c] = acc;
}
__global__ void __launch_bounds__(256) sn_w_v_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                     float* __restrict__ u_raw, int rows, int cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float acc = 0.f;
  for (int c = lane; c < cols; c += 32) acc = fmaf(w[static_cast<size_t>(warp) * cols + c], v[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) u_raw[warp] = acc;
}
// single block: dst = src / max(||src||, eps); optionally *dot_out = dot(dst, src2) (src2 may alias src)
__global__ void __launch_bounds__(1024) sn_normalize_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int n, float eps, const float* __restrict__ dot_a,
                                                            const float* __restrict__ dot_b, float* dot_out,
                                                            float* inv_dot_out) {
  __shared__ float red[32];
  __shared__ float s_inv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (dst) {
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) q = fmaf(src[i], src[i], q);
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (warp == 0) {
      float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) s_inv = 1.f / fmaxf(sqrtf(t), eps);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i] * s_inv;
    __syncthreads();
  }
  if (dot_out) {
    float d = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d = fmaf(dot_a[i], dot_b[i], d);
    d = warp_sum(d);
    if (lane == 0) red[warp] = d;
    __syncthreads();
    if (warp == 0) {
      float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) { *dot_out = t; if (inv_dot_out) *inv_dot_out = 1.f / t; }
    }
  }
}

// ---- GAN / distortion reductions -------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float v) { return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v))); }

// logits: [2*half] = real logits then generated logits (torch.chunk(.., 2, dim=0), model.py:185-186)
// sums[0] = sum BCE(real, 1), sums[1] = sum BCE(gen, 0), sums[2] = sum BCE(gen, 1),
// sums[3] = sum sigmoid(real), sums[4] = sum sigmoid(gen)
__global__ void __launch_bounds__(256) gan_sums_kernel(const float* __restrict__ logits, int64_t half,
                                                       double* __restrict__ sums) {
  __shared__ float red[8];
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < half; i += stride) {
    const float r = logits[i], g = logits[half + i];
    a[0] += softplus_f(-r);
    a[1] += softplus_f(g);
    a[2] += softplus_f(-g);
    a[3] += 1.f / (1.f + expf(-r));
    a[4] += 1.f / (1.f + expf(-g));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < 5; ++k) {
    float v = warp_sum(a[k]);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) atomicAdd(&sums[k], static_cast<double>(t));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         int64_t count, float scale, double* __restrict__ sum) {
  __shared__ float red[8];
  float acc = 0.f;
  const int64_t nvec = count / 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float d0 = x.x * scale - y.x * scale, d1 = x.y * scale - y.y * scale;
    const float d2 = x.z * scale - y.z * scale, d3 = x.w * scale - y.w * scale;
    acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  for (int64_t i = nvec * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float d = a[i] * scale - b[i] * scale;
    acc += d * d;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v = warp_sum(acc);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(sum, static_cast<double>(t));
  }
}


// ---- LPIPS feature loss of one trunk layer ------------------------------------------------------------
// networks_basic.py:61-89 + perceptual_loss.py:42-46: per pixel  sum_c w_c (f0_c/|f0| - f1_c/|f1|)^2 with
// |f| = sqrt(sum_c f_c^2 + 1e-10), then the spatial mean, accumulated into out[image].
// One thread per pixel; channel planes are read coalesced along (h, w); two passes (norms, then distance).
__global__ void __launch_bounds__(256) lpips_layer_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                          const float* __restrict__ lin_w, int c, int hw,
                                                          float* __restrict__ out) {
  __shared__ float red[8];
  const int img = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (px < hw) {
    const float* a = f0 + static_cast<size_t>(img) * c * hw + px;
    const float* b = f1 + static_cast<size_t>(img) * c * hw + px;
    float na = 0.f, nb = 0.f;
    for (int k = 0; k < c; ++k) {
      const float x = a[static_cast<size_t>(k) * hw], y = b[static_cast<size_t>(k) * hw];
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
    const float ia = 1.f / sqrtf(na + 1e-10f), ib = 1.f / sqrtf(nb + 1e-10f);
    for (int k = 0; k < c; ++k) {
      const float d = a[static_cast<size_t>(k) * hw] * ia - b[static_cast<size_t>(k) * hw] * ib;
      acc = fmaf(lin_w[k] * d, d, acc);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v = warp_sum(acc);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(&out[img], t / static_cast<float>(hw));
  }
}

}  // namespace hfc

using namespace hfc;

#define HFC_CHECK_LAUNCH(what)                                                              \
  do {                                                                                      \
    cudaError_t e_ = cudaGetLastError();                                                    \
    if (e_ != cudaSuccess) return set_error(HFC_ERR_LAUNCH, what ": %s", cudaGetErrorString(e_)); \
    note_launch();                                                                          \
  } while (0)

extern "C" int hfc_disc_input(const float* x, int32_t x_channels, const void* ctx_act, const hfc_act_geom* ctx,
                              int32_t scale, const hfc_act_geom* out_geom, void* out, void* stream) {
  if (!x || !ctx_act || !ctx || !out_geom || !out) return set_error(HFC_ERR_INVALID, "disc_input: null pointer");
  if (ctx->pt | ctx->pl | ctx->pb | ctx->pr) return set_error(HFC_ERR_INVALID, "disc_input: context buffer must be border-less");
  if (out_geom->n != ctx->n || out_geom->h != ctx->h * scale || out_geom->w != ctx->w * scale ||
      out_geom->c != x_channels + ctx->c || out_geom->cpad % 8 != 0 || out_geom->cpad < out_geom->c)
    return set_error(HFC_ERR_INVALID, "disc_input: inconsistent geometry");
  if (out_geom->pt >= out_geom->h || out_geom->pl >= out_geom->w) return set_error(HFC_ERR_INVALID, "disc_input: border too wide");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  DiscInParams p;
  p.n = out_geom->n; p.h = out_geom->h; p.w = out_geom->w; p.cx = x_channels;
  p.ch = ctx->h; p.cw = ctx->w; p.cc = ctx->c; p.ccpad = ctx->cpad; p.scale = scale;
  p.cpad = out_geom->cpad; p.pt = out_geom->pt; p.pl = out_geom->pl; p.pb = out_geom->pb; p.pr = out_geom->pr;
  const long long npix = static_cast<long long>(p.n) * p.h * p.w;
  disc_input_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const __half*>(ctx_act), reinterpret_cast<__half*>(out), p);
  HFC_CHECK_LAUNCH("disc_input launch");
  return HFC_OK;
}

extern "C" int hfc_spectral_sigma(const float* w, int32_t rows, int32_t cols, float* u, float* v,
                                  int32_t power_iteration, float* workspace, float* sigma, float* inv_sigma,
                                  void* stream) {
  if (!w || !u || !v || !workspace || !sigma || rows <= 0 || cols <= 0)
    return set_error(HFC_ERR_INVALID, "spectral_sigma: null pointer or empty matrix");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* u_raw = workspace;          // [rows]
  float* v_raw = workspace + rows;   // [cols]
  const float eps = 1e-12f;
  if (power_iteration) {
    sn_wt_u_kernel<<<(cols + 255) / 256, 256, 0, st>>>(w, u, v_raw, rows, cols);
    HFC_CHECK_LAUNCH("spectral_sigma launch");
    sn_normalize_kernel<<<1, 1024, 0, st>>>(v_raw, v, cols, eps, nullptr, nullptr, nullptr, nullptr);
    HFC_CHECK_LAUNCH("spectral_sigma launch");
  }
  sn_w_v_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(w, v, u_raw, rows, cols);
  HFC_CHECK_LAUNCH("spectral_sigma launch");
  // training: u = normalize(W v), sigma = u . (W v);  eval: sigma = u_stored . (W v)
  sn_normalize_kernel<<<1, 1024, 0, st>>>(u_raw, power_iteration ? u : nullptr, rows, eps, u, u_raw, sigma, inv_sigma);
  HFC_CHECK_LAUNCH("spectral_sigma launch");
  return HFC_OK;
}

extern "C" int hfc_gan_sums(const float* logits, int64_t half_count, double* sums5, void* stream) {
  if (!logits || !sums5 || half_count <= 0) return set_error(HFC_ERR_INVALID, "gan_sums: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((half_count + 255) / 256, sms * 4LL)));
  gan_sums_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, half_count, sums5);
  HFC_CHECK_LAUNCH("gan_sums launch");
  return HFC_OK;
}

extern "C" int hfc_sqdiff_sum(const float* a, const float* b, int64_t count, float scale, double* sum, void* stream) {
  if (!a || !b || !sum || count <= 0) return set_error(HFC_ERR_INVALID, "sqdiff_sum: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((count / 4 + 255) / 256, sms * 8LL)));
  sqdiff_sum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, count, scale, sum);
  HFC_CHECK_LAUNCH("sqdiff_sum launch");
  return HFC_OK;
}

extern "C" int hfc_lpips_layer(const float* f0, const float* f1, const float* lin_w, int32_t n, int32_t c,
                               int32_t hw, float* out_per_image, void* stream) {
  if (!f0 || !f1 || !lin_w || !out_per_image || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "lpips_layer: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  dim3 grid((hw + 255) / 256, n);
  lpips_layer_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(f0, f1, lin_w, c, hw, out_per_image);
  HFC_CHECK_LAUNCH("lpips_layer launch");
  return HFC_OK;
}
