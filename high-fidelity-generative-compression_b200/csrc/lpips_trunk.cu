// LPIPS (AlexNet) trunk around the tcgen05 conv kernel: everything of
// src/loss/perceptual_similarity/{perceptual_loss.py:26-46, networks_basic.py:61-98, pretrained_networks.py:56-94}
// that is not a convolution, in the internal NHWC fp16 activation format, forward and backward:
//
//   lpips_prep          2x-1 (normalize), ScalingLayer ((x - shift) / scale), zero padding 2 and the 4x4 space-to-depth
//                       that turns AlexNet's 11x11 stride-4 conv into a 3x3 stride-1 conv over 48 channels
//                       (out[i] = sum_k x'[4i + k] w[k], k = 4a + b  ->  sum_a sum_b s2d[i + a][b] w[4a + b])
//   maxpool_3x3s2       nn.MaxPool2d(3, 2) on NHWC fp16
//   lpips_nhwc          per-pixel channel normalisation, squared difference, 1x1 'lin' weights, spatial mean, for the
//                       target half [0, n) and the reconstruction half [n, 2n) of one feature buffer
//   *_bwd               their adjoints (the trunk is frozen: data gradients only, reconstruction half only)
// All HBM-bound; the trunk's convolutions themselves are hfc_conv_forward launches (engine: loss/lpips_trunk.py).
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

#include <cuda_fp16.h>

namespace hfc {

// ---- input preparation ----------------------------------------------------------------------------------------------
// out: (2n, hs, ws, 64) fp16, channel (dy * 4 + dx) * 3 + c of s2d pixel (I, J) = scaled image at (4I + dy - 2, 4J + dx - 2),
// zero outside the image and in channels 48..63.  Images [0, n) come from `target`, [n, 2n) from `pred`.
__global__ void __launch_bounds__(256)
lpips_prep_kernel(const float* __restrict__ target, const float* __restrict__ pred, int n, int h, int w, int hs, int ws,
                  int normalize, const float* __restrict__ shift, const float* __restrict__ scale,
                  __half* __restrict__ out) {
  const long long total = 2LL * n * hs * ws;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int J = static_cast<int>(t % ws);
  const int I = static_cast<int>((t / ws) % hs);
  const int img = static_cast<int>(t / (static_cast<long long>(ws) * hs));
  const float* src = (img < n ? target + static_cast<size_t>(img) * 3 * h * w
                              : pred + static_cast<size_t>(img - n) * 3 * h * w);
  float sh[3], isc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { sh[c] = shift[c]; isc[c] = scale[c]; }
  __align__(16) __half vals[64];
#pragma unroll
  for (int dy = 0; dy < 4; ++dy) {
    const int y = 4 * I + dy - 2;
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
      const int x = 4 * J + dx - 2;
      const bool in = y >= 0 && y < h && x >= 0 && x < w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = 0.f;
        if (in) {
          v = src[(static_cast<size_t>(c) * h + y) * w + x];
          if (normalize) v = 2.f * v - 1.f;                 // perceptual_loss.py:36-38
          v = (v - sh[c]) / isc[c];                         // ScalingLayer, networks_basic.py:91-98
        }
        vals[(dy * 4 + dx) * 3 + c] = __float2half_rn(v);
      }
    }
  }
#pragma unroll
  for (int k = 48; k < 64; ++k) vals[k] = __float2half_rn(0.f);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * 64);
#pragma unroll
  for (int k = 0; k < 8; ++k) dst[k] = reinterpret_cast<const uint4*>(vals)[k];
}

// g: fp32 rows [n * hs * ws][ld] (gradient w.r.t. the s2d buffer of the reconstruction half) -> dpred (n, 3, h, w)
__global__ void __launch_bounds__(256)
lpips_prep_bwd_kernel(const float* __restrict__ g, int ld, int n, int h, int w, int hs, int ws, int normalize,
                      const float* __restrict__ scale, float* __restrict__ dpred) {
  const long long total = static_cast<long long>(n) * 3 * h * w;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int x = static_cast<int>(t % w);
  const int y = static_cast<int>((t / w) % h);
  const int c = static_cast<int>((t / (static_cast<long long>(w) * h)) % 3);
  const int img = static_cast<int>(t / (3LL * w * h));
  const int yy = y + 2, xx = x + 2;
  const int I = yy >> 2, J = xx >> 2;
  float v = 0.f;
  if (I < hs && J < ws) v = g[((static_cast<size_t>(img) * hs + I) * ws + J) * ld + ((yy & 3) * 4 + (xx & 3)) * 3 + c];
  dpred[t] = v * (normalize ? 2.f : 1.f) / scale[c];
}

// ---- MaxPool2d(3, 2) ------------------------------------------------------------------------------------------------
// in: (n, h, w, cpad) fp16 border-less; out: (n, oh, ow, cpad); one thread = 8 channels of one output pixel
__global__ void __launch_bounds__(256)
maxpool_kernel(const __half* __restrict__ in, int n, int h, int w, int cpad, int oh, int ow, __half* __restrict__ out) {
  const int c8 = cpad / 8;
  const long long total = static_cast<long long>(n) * oh * ow * c8;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int ch = static_cast<int>(t % c8) * 8;
  const int ox = static_cast<int>((t / c8) % ow);
  const int oy = static_cast<int>((t / (static_cast<long long>(c8) * ow)) % oh);
  const int img = static_cast<int>(t / (static_cast<long long>(c8) * ow * oh));
  __half2 m[4];
  bool first = true;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int y = 2 * oy + ky, x = 2 * ox + kx;        // always inside: oh = (h - 3) / 2 + 1
      const uint4 v = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(img) * h + y) * w + x) * cpad + ch);
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = first ? hv[k] : __hmax2(m[k], hv[k]);
      first = false;
    }
  uint4 o;
  __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) ho[k] = m[k];
  *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * oh + oy) * ow + ox) * cpad + ch) = o;
}

// Adjoint: the gradient of an output pixel goes to the FIRST maximum of its window in row-major scan order (ATen's
// max_pool2d semantics).  g_out: fp32 rows [n * oh * ow][ld_out]; g_in: fp32 rows [n * h * w][ld_in], ZEROED by the
// caller (windows overlap: atomics).  `in` is the pooled layer's input (the saved feature map of these n images).
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ g_out, int ld_out, const __half* __restrict__ in, int n, int h, int w,
                   int c, int cpad, int oh, int ow, float* __restrict__ g_in, int ld_in) {
  const long long total = static_cast<long long>(n) * oh * ow * c;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int ch = static_cast<int>(t % c);
  const int ox = static_cast<int>((t / c) % ow);
  const int oy = static_cast<int>((t / (static_cast<long long>(c) * ow)) % oh);
  const int img = static_cast<int>(t / (static_cast<long long>(c) * ow * oh));
  const float go = g_out[((static_cast<size_t>(img) * oh + oy) * ow + ox) * ld_out + ch];
  if (go == 0.f) return;
  float best = 0.f;
  int by = 0, bx = 0;
  bool first = true;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int y = 2 * oy + ky, x = 2 * ox + kx;
      const float v = __half2float(in[((static_cast<size_t>(img) * h + y) * w + x) * cpad + ch]);
      if (first || v > best) { best = v; by = y; bx = x; first = false; }
    }
  atomicAdd(&g_in[((static_cast<size_t>(img) * h + by) * w + bx) * ld_in + ch], go);
}

// ---- LPIPS layer on NHWC fp16 features ------------------------------------------------------------------------------
// feat: (2n, hw, cpad); one warp per pixel, lanes over channels.  out[img] += mean_hw sum_c w_c (f0/|f0| - f1/|f1|)^2
template <bool BWD>
__global__ void __launch_bounds__(256)
lpips_nhwc_kernel(const __half* __restrict__ feat, int n, int hw, int c, int cpad, const float* __restrict__ lin_w,
                  float* __restrict__ out, const float* __restrict__ upstream, const float* __restrict__ g_in,
                  int ld_g, float* __restrict__ g_out, int ld_out) {
  const int lane = threadIdx.x & 31;
  const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (pix >= static_cast<long long>(n) * hw) return;
  const int img = static_cast<int>(pix / hw);
  const __half* a = feat + static_cast<size_t>(pix) * cpad;                               // target
  const __half* b = feat + (static_cast<size_t>(n) * hw + static_cast<size_t>(pix)) * cpad;  // reconstruction
  float na = 0.f, nb = 0.f;
  for (int k = lane; k < c; k += 32) {
    const float x = __half2float(a[k]), y = __half2float(b[k]);
    na = fmaf(x, x, na);
    nb = fmaf(y, y, nb);
  }
  na = warp_sum(na);
  nb = warp_sum(nb);
  const float ia = 1.f / sqrtf(na + 1e-10f), ib = 1.f / sqrtf(nb + 1e-10f);             // perceptual_loss.py:42-46
  if (!BWD) {
    float acc = 0.f;
    for (int k = lane; k < c; k += 32) {
      const float d = __half2float(a[k]) * ia - __half2float(b[k]) * ib;
      acc = fmaf(lin_w[k] * d, d, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(&out[img], acc / static_cast<float>(hw));
  } else {
    float dot = 0.f;                                   // b_hat . gb
    for (int k = lane; k < c; k += 32) {
      const float bh = __half2float(b[k]) * ib;
      const float gb = -2.f * lin_w[k] * (__half2float(a[k]) * ia - bh);
      dot = fmaf(bh, gb, dot);
    }
    dot = warp_sum(dot);
    const float up = upstream[img] / static_cast<float>(hw);
    for (int k = lane; k < c; k += 32) {
      const float bv = __half2float(b[k]);
      const float bh = bv * ib;
      const float gb = -2.f * lin_w[k] * (__half2float(a[k]) * ia - bh);
      float gr = up * (gb - bh * dot) * ib;            // d / d f1 of this layer's distance (|f1| has eps inside the sqrt)
      if (g_in) gr += g_in[static_cast<size_t>(pix) * ld_g + k];   // gradient arriving from the deeper layers
      g_out[static_cast<size_t>(pix) * ld_out + k] = bv > 0.f ? gr : 0.f;   // through the ReLU that produced f1
    }
  }
}

}  // namespace hfc

using namespace hfc;

#define HFC_LAUNCH_END(what)                                                                          \
  do {                                                                                                \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess) return set_error(HFC_ERR_LAUNCH, what " launch: %s", cudaGetErrorString(e_)); \
    note_launch();                                                                                    \
    return HFC_OK;                                                                                    \
  } while (0)

static inline unsigned blocks_for(long long threads) { return static_cast<unsigned>((threads + 255) / 256); }

extern "C" int hfc_lpips_prep(const float* target, const float* pred, int32_t n, int32_t h, int32_t w, int32_t hs,
                              int32_t ws, int32_t normalize, const float* shift3, const float* scale3, void* out_act,
                              void* stream) {
  if (!target || !pred || !shift3 || !scale3 || !out_act || n <= 0 || h <= 0 || w <= 0 || hs <= 0 || ws <= 0)
    return set_error(HFC_ERR_INVALID, "lpips_prep: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  lpips_prep_kernel<<<blocks_for(2LL * n * hs * ws), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      target, pred, n, h, w, hs, ws, normalize, shift3, scale3, static_cast<__half*>(out_act));
  HFC_LAUNCH_END("lpips_prep");
}

extern "C" int hfc_lpips_prep_bwd(const float* g_rows, int32_t ld, int32_t n, int32_t h, int32_t w, int32_t hs,
                                  int32_t ws, int32_t normalize, const float* scale3, float* dpred, void* stream) {
  if (!g_rows || !scale3 || !dpred || n <= 0 || h <= 0 || w <= 0 || ld < 48)
    return set_error(HFC_ERR_INVALID, "lpips_prep_bwd: null pointer, empty input or ld < 48");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  lpips_prep_bwd_kernel<<<blocks_for(3LL * n * h * w), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g_rows, ld, n, h, w, hs, ws, normalize, scale3, dpred);
  HFC_LAUNCH_END("lpips_prep_bwd");
}

extern "C" int hfc_maxpool3s2(const void* in_act, const hfc_act_geom* g, void* out_act, void* stream) {
  if (!in_act || !g || !out_act || g->n <= 0 || g->h < 3 || g->w < 3 || g->cpad % 8 || g->pt || g->pl || g->pb || g->pr)
    return set_error(HFC_ERR_INVALID, "maxpool3s2: needs a border-less buffer of at least 3x3 pixels, cpad % 8 == 0");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int oh = (g->h - 3) / 2 + 1, ow = (g->w - 3) / 2 + 1;
  maxpool_kernel<<<blocks_for(static_cast<long long>(g->n) * oh * ow * (g->cpad / 8)), 256, 0,
                   static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(in_act), g->n, g->h, g->w, g->cpad,
                                                        oh, ow, static_cast<__half*>(out_act));
  HFC_LAUNCH_END("maxpool3s2");
}

extern "C" int hfc_maxpool3s2_bwd(const float* g_out_rows, int32_t ld_out, const void* in_act, const hfc_act_geom* g,
                                  float* g_in_rows, int32_t ld_in, void* stream) {
  if (!g_out_rows || !in_act || !g || !g_in_rows || g->n <= 0 || g->h < 3 || g->w < 3 || ld_out < g->c || ld_in < g->c ||
      g->pt || g->pl || g->pb || g->pr)
    return set_error(HFC_ERR_INVALID, "maxpool3s2_bwd: bad geometry or row pitch");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int oh = (g->h - 3) / 2 + 1, ow = (g->w - 3) / 2 + 1;
  maxpool_bwd_kernel<<<blocks_for(static_cast<long long>(g->n) * oh * ow * g->c), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(g_out_rows, ld_out, static_cast<const __half*>(in_act), g->n,
                                                            g->h, g->w, g->c, g->cpad, oh, ow, g_in_rows, ld_in);
  HFC_LAUNCH_END("maxpool3s2_bwd");
}

extern "C" int hfc_lpips_nhwc(const void* feat_act, int32_t n, int32_t hw, int32_t c, int32_t cpad, const float* lin_w,
                              float* out_per_image, void* stream) {
  if (!feat_act || !lin_w || !out_per_image || n <= 0 || hw <= 0 || c <= 0 || cpad < c)
    return set_error(HFC_ERR_INVALID, "lpips_nhwc: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  lpips_nhwc_kernel<false><<<blocks_for(32LL * n * hw), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(feat_act), n, hw, c, cpad, lin_w, out_per_image, nullptr, nullptr, 0, nullptr, 0);
  HFC_LAUNCH_END("lpips_nhwc");
}

extern "C" int hfc_lpips_nhwc_bwd(const void* feat_act, int32_t n, int32_t hw, int32_t c, int32_t cpad,
                                  const float* lin_w, const float* upstream, const float* g_in_rows, int32_t ld_g,
                                  float* g_out_rows, int32_t ld_out, void* stream) {
  if (!feat_act || !lin_w || !upstream || !g_out_rows || n <= 0 || hw <= 0 || c <= 0 || cpad < c || ld_out < c ||
      (g_in_rows && ld_g < c))
    return set_error(HFC_ERR_INVALID, "lpips_nhwc_bwd: null pointer, empty input or row pitch < c");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  lpips_nhwc_kernel<true><<<blocks_for(32LL * n * hw), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(feat_act), n, hw, c, cpad, lin_w, nullptr, upstream, g_in_rows, ld_g, g_out_rows, ld_out);
  HFC_LAUNCH_END("lpips_nhwc_bwd");
}
