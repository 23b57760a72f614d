// Discretised mixture likelihood of the latents (`-LMM`, train.py:227): HyperpriorDLMM.latent_log_likelihood_DLMM
// src/hyperprior.py:379-401 with unpack_likelihood_params src/network/hyper.py:19-35, the noise / round quantisation
// (hyperprior.py:57-78), the straight-through latents (:108-122, :443-446) and _estimate_entropy_log (:95-106), fused
// into one pass forward and one pass backward.
//
//   params (n, 3*c*k, hw): plane (s*c + ch)*k + j, s = 0 mixture logits, 1 means, 2 log-scales (lower-bounded at -3)
//   L(v) = logsumexp_j [ log_softmax(logit)_j + log max(Phi(inv_j (.5 - |v - mu_j|)) - Phi(inv_j (-.5 - |v - mu_j|)), 1e-9) ]
//   forward : sums[0] += sum L(x + noise), sums[1] += sum L(floor(x + .5)), decoded
//   backward: d / d x and d / d params of  g * coef * sum L(x + noise)  (+ the straight-through gradient of `decoded`)
//
// HBM-bound (4 + 4 + 12k B read, 4 B written per element forward); non-default variant, written for accuracy first:
// erfcf / expf / logf, one thread per latent element, parameter planes coalesced along hw.
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

namespace hfc {

constexpr int kMaxMix = 8;
constexpr float kLogScalesMin = -3.f;      // LOG_SCALES_MIN, src/hyperprior.py:31
constexpr float kMinLik = 1e-9f;

__device__ __forceinline__ float dlmm_cdf(float t, int type) {
  return type == 0 ? 0.5f * erfcf(t * -0.70710678118654752440f) : 1.f / (1.f + expf(-t));   // maths.py:102-109
}
__device__ __forceinline__ float dlmm_pdf(float t, int type) {
  if (type == 0) return 0.39894228040143267794f * expf(-0.5f * t * t);
  const float s = 1.f / (1.f + expf(-t));
  return s * (1.f - s);
}

struct Mix {
  float logit[kMaxMix], mu[kMaxMix], ls_raw[kMaxMix];
};

__device__ __forceinline__ void load_mix(const float* __restrict__ params, int64_t img_base, int c, int k, int hw, int ch,
                                         int px, Mix& m) {
  const int64_t plane = static_cast<int64_t>(c) * k * hw;
#pragma unroll
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) {
      const int64_t o = img_base + (static_cast<int64_t>(ch) * k + j) * hw + px;
      m.logit[j] = params[o];
      m.mu[j] = params[o + plane];
      m.ls_raw[j] = params[o + 2 * plane];
    }
}

// log-likelihood of value v; optionally the softmax weights w_j = dL/da_j and the per-component raw pmf
__device__ __forceinline__ float dlmm_loglik(float v, const Mix& m, int k, int type, float lse_logit) {
  float a[kMaxMix];
  float amax = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) {
      const float inv = expf(-fmaxf(m.ls_raw[j], kLogScalesMin));
      const float d = fabsf(v - m.mu[j]);
      const float p = fmaxf(dlmm_cdf(inv * (0.5f - d), type) - dlmm_cdf(inv * (-0.5f - d), type), kMinLik);
      a[j] = (m.logit[j] - lse_logit) + logf(p);
      amax = fmaxf(amax, a[j]);
    }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) s += expf(a[j] - amax);
  return amax + logf(s);
}

__device__ __forceinline__ float logsumexp_logits(const Mix& m, int k) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) mx = fmaxf(mx, m.logit[j]);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) s += expf(m.logit[j] - mx);
  return mx + logf(s);
}

__global__ void __launch_bounds__(256)
dlmm_likelihood_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ params,
                       int n, int c, int k, int hw, int type, int straight_through, float* __restrict__ decoded,
                       double* __restrict__ sums) {
  __shared__ float red[2][8];
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float acc_n = 0.f, acc_q = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int px = static_cast<int>(i % hw);
    const int ch = static_cast<int>((i / hw) % c);
    const int64_t img = i / (static_cast<int64_t>(hw) * c);
    Mix m;
    load_mix(params, img * 3 * c * k * hw, c, k, hw, ch, px, m);
    const float lse = logsumexp_logits(m, k);
    const float xv = x[i];
    const float q = floorf(xv + 0.5f);
    acc_q += dlmm_loglik(q, m, k, type, lse);
    if (noise) acc_n += dlmm_loglik(xv + noise[i], m, k, type, lse);
    if (decoded) decoded[i] = straight_through ? xv + (q - xv) : q;
  }
  acc_n = warp_sum(acc_n);
  acc_q = warp_sum(acc_q);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = acc_n; red[1][warp] = acc_q; }
  __syncthreads();
  if (warp == 0) {
    float rn = lane < 8 ? red[0][lane] : 0.f, rq = lane < 8 ? red[1][lane] : 0.f;
    rn = warp_sum(rn);
    rq = warp_sum(rq);
    if (lane == 0) {
      if (noise) atomicAdd(&sums[0], static_cast<double>(rn));
      atomicAdd(&sums[1], static_cast<double>(rq));
    }
  }
}

__global__ void __launch_bounds__(256)
dlmm_likelihood_bwd_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ params,
                           const float* __restrict__ d_decoded, const float* __restrict__ g_nbpp, float coef, int n,
                           int c, int k, int hw, int type, float* __restrict__ dx, float* __restrict__ dparams) {
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float g = (g_nbpp ? g_nbpp[0] : 0.f) * coef;                // d loss / d (sum of log-likelihoods)
  const int64_t plane = static_cast<int64_t>(c) * k * hw;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int px = static_cast<int>(i % hw);
    const int ch = static_cast<int>((i / hw) % c);
    const int64_t img = i / (static_cast<int64_t>(hw) * c);
    const int64_t base = img * 3 * plane;
    Mix m;
    load_mix(params, base, c, k, hw, ch, px, m);
    const float lse = logsumexp_logits(m, k);
    const float v = x[i] + noise[i];
    float a[kMaxMix], praw[kMaxMix], inv[kMaxMix], tu[kMaxMix], tl[kMaxMix];
    float amax = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxMix; ++j)
      if (j < k) {
        inv[j] = expf(-fmaxf(m.ls_raw[j], kLogScalesMin));
        const float d = fabsf(v - m.mu[j]);
        tu[j] = inv[j] * (0.5f - d);
        tl[j] = inv[j] * (-0.5f - d);
        praw[j] = dlmm_cdf(tu[j], type) - dlmm_cdf(tl[j], type);
        a[j] = (m.logit[j] - lse) + logf(fmaxf(praw[j], kMinLik));
        amax = fmaxf(amax, a[j]);
      }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxMix; ++j)
      if (j < k) s += expf(a[j] - amax);
    float dv = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxMix; ++j)
      if (j < k) {
        const float w = expf(a[j] - amax) / s;                       // d L / d a_j
        const float pi = expf(m.logit[j] - lse);                     // softmax(logit)_j
        const int64_t o = base + (static_cast<int64_t>(ch) * k + j) * hw + px;
        dparams[o] = g * (w - pi);                                   // a_j = logit_j - lse(logit) + log p_j
        // log p_j -> p_j: LowerBoundToward(p, 1e-9) passes where p >= bound or the gradient is negative (maths.py:96-100)
        float gp = g * w / fmaxf(praw[j], kMinLik);
        if (!(praw[j] >= kMinLik || gp < 0.f)) gp = 0.f;
        const float d = fabsf(v - m.mu[j]);
        const float fu = dlmm_pdf(tu[j], type), fl = dlmm_pdf(tl[j], type);
        const float dp_dd = -inv[j] * (fu - fl);
        const float dp_dinv = fu * (0.5f - d) - fl * (-0.5f - d);
        const float sgn = v > m.mu[j] ? 1.f : (v < m.mu[j] ? -1.f : 0.f);
        dv += gp * dp_dd * sgn;
        dparams[o + plane] = -gp * dp_dd * sgn;                      // d / d mean_j
        float gls = gp * dp_dinv * -inv[j];                          // inv = exp(-ls)
        if (!(m.ls_raw[j] >= kLogScalesMin || gls < 0.f)) gls = 0.f; // LowerBoundToward(log_scales, -3)
        dparams[o + 2 * plane] = gls;
      }
    dx[i] = dv + (d_decoded ? d_decoded[i] : 0.f);                   // straight-through latents: d decoded / d x = 1
  }
}

}  // namespace hfc

using namespace hfc;

static int dlmm_blocks(int64_t count, int sms) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, static_cast<int64_t>(sms) * 8)));
}

extern "C" int hfc_dlmm_likelihood(const float* x, const float* noise, const float* dlmm_params, int32_t n, int32_t c,
                                   int32_t k, int32_t hw, int32_t likelihood_type, int32_t straight_through,
                                   float* decoded, double* sums, void* stream) {
  if (!x || !dlmm_params || !sums || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "dlmm_likelihood: null pointer or empty input");
  if (k < 1 || k > kMaxMix) return set_error(HFC_ERR_UNSUPPORTED, "dlmm_likelihood: 1..%d mixture components", kMaxMix);
  if (likelihood_type != 0 && likelihood_type != 1) return set_error(HFC_ERR_INVALID, "dlmm_likelihood: likelihood_type must be 0 or 1");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  dlmm_likelihood_kernel<<<dlmm_blocks(count, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, noise, dlmm_params, n, c, k, hw, likelihood_type, straight_through, decoded, sums);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "dlmm_likelihood launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

extern "C" int hfc_dlmm_likelihood_bwd(const float* x, const float* noise, const float* dlmm_params,
                                       const float* d_decoded, const float* g_nbpp, float coef, int32_t n, int32_t c,
                                       int32_t k, int32_t hw, int32_t likelihood_type, float* dx, float* dparams,
                                       void* stream) {
  if (!x || !noise || !dlmm_params || !dx || !dparams || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "dlmm_likelihood_bwd: null pointer or empty input");
  if (k < 1 || k > kMaxMix) return set_error(HFC_ERR_UNSUPPORTED, "dlmm_likelihood_bwd: 1..%d mixture components", kMaxMix);
  if (likelihood_type != 0 && likelihood_type != 1) return set_error(HFC_ERR_INVALID, "dlmm_likelihood_bwd: likelihood_type must be 0 or 1");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  dlmm_likelihood_bwd_kernel<<<dlmm_blocks(count, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, noise, dlmm_params, d_decoded, g_nbpp, coef, n, c, k, hw, likelihood_type, dx, dparams);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "dlmm_likelihood_bwd launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
