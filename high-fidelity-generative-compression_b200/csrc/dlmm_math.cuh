// Per-element arithmetic of the discretised mixture likelihood (`-LMM`): shared by the CUDA kernels (csrc/dlmm.cu) and by
// the host harness tests/dlmm_math_host.cpp, which compiles THIS file with g++ so that the formulas the GPU runs are checked
// against the oracle and torch autograd on the CPU.  Reference: HyperpriorDLMM.latent_log_likelihood_DLMM
// src/hyperprior.py:379-401, unpack_likelihood_params src/network/hyper.py:19-35, LowerBoundToward src/helpers/maths.py:87-100.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define HFC_HD __host__ __device__ __forceinline__
#else
#define HFC_HD inline
#endif

namespace hfc {

constexpr int kMaxMix = 8;
constexpr float kLogScalesMin = -3.f;      // LOG_SCALES_MIN, src/hyperprior.py:31
constexpr float kMinLik = 1e-9f;

HFC_HD float dlmm_cdf(float t, int type) {
  return type == 0 ? 0.5f * erfcf(t * -0.70710678118654752440f) : 1.f / (1.f + expf(-t));   // maths.py:102-109
}
HFC_HD float dlmm_pdf(float t, int type) {
  if (type == 0) return 0.39894228040143267794f * expf(-0.5f * t * t);
  const float s = 1.f / (1.f + expf(-t));
  return s * (1.f - s);
}

struct Mix {
  float logit[kMaxMix], mu[kMaxMix], ls_raw[kMaxMix];
};

HFC_HD float logsumexp_logits(const Mix& m, int k) {
  float mx = -INFINITY;
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) mx = fmaxf(mx, m.logit[j]);
  float s = 0.f;
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) s += expf(m.logit[j] - mx);
  return mx + logf(s);
}

// L(v) = logsumexp_j [ (logit_j - lse) + log max(Phi(inv_j (.5 - |v - mu_j|)) - Phi(inv_j (-.5 - |v - mu_j|)), 1e-9) ]
HFC_HD float dlmm_loglik(float v, const Mix& m, int k, int type, float lse_logit) {
  float a[kMaxMix];
  float amax = -INFINITY;
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) {
      const float inv = expf(-fmaxf(m.ls_raw[j], kLogScalesMin));
      const float d = fabsf(v - m.mu[j]);
      const float p = fmaxf(dlmm_cdf(inv * (0.5f - d), type) - dlmm_cdf(inv * (-0.5f - d), type), kMinLik);
      a[j] = (m.logit[j] - lse_logit) + logf(p);
      amax = fmaxf(amax, a[j]);
    }
  float s = 0.f;
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) s += expf(a[j] - amax);
  return amax + logf(s);
}

// Gradient of g * L(v): returns d / d v; d / d (logit_j, mean_j, log-scale_j) into the three arrays.  Both
// LowerBoundToward gates: pass where the input was >= the bound or the incoming gradient is negative.
HFC_HD float dlmm_grad(float v, const Mix& m, int k, int type, float lse, float g, float* d_logit, float* d_mu,
                       float* d_ls) {
  float a[kMaxMix], praw[kMaxMix], inv[kMaxMix], tu[kMaxMix], tl[kMaxMix];
  float amax = -INFINITY;
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
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) s += expf(a[j] - amax);
  float dv = 0.f;
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) {
      const float w = expf(a[j] - amax) / s;                       // d L / d a_j
      const float pi = expf(m.logit[j] - lse);                     // softmax(logit)_j
      d_logit[j] = g * (w - pi);                                   // a_j = logit_j - lse(logit) + log p_j
      float gp = g * w / fmaxf(praw[j], kMinLik);                  // through log, then LowerBoundToward(p, 1e-9)
      if (!(praw[j] >= kMinLik || gp < 0.f)) gp = 0.f;
      const float d = fabsf(v - m.mu[j]);
      const float fu = dlmm_pdf(tu[j], type), fl = dlmm_pdf(tl[j], type);
      const float dp_dd = -inv[j] * (fu - fl);
      const float dp_dinv = fu * (0.5f - d) - fl * (-0.5f - d);
      const float sgn = v > m.mu[j] ? 1.f : (v < m.mu[j] ? -1.f : 0.f);
      dv += gp * dp_dd * sgn;
      d_mu[j] = -gp * dp_dd * sgn;
      float gls = gp * dp_dinv * -inv[j];                          // inv = exp(-ls)
      if (!(m.ls_raw[j] >= kLogScalesMin || gls < 0.f)) gls = 0.f; // LowerBoundToward(log_scales, -3)
      d_ls[j] = gls;
    }
  return dv;
}

// ---- whole-element bodies (index decomposition + loads + stores), shared by the kernels and the host harness ----------
HFC_HD void load_mix(const float* params, int64_t img_base, int c, int k, int hw, int ch, int px, Mix& m) {
  const int64_t plane = static_cast<int64_t>(c) * k * hw;
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) {
      const int64_t o = img_base + (static_cast<int64_t>(ch) * k + j) * hw + px;
      m.logit[j] = params[o];
      m.mu[j] = params[o + plane];
      m.ls_raw[j] = params[o + 2 * plane];
    }
}

// element i of x (n, c, hw): adds L(x + noise) to *acc_n (if noise) and L(floor(x + .5)) to *acc_q, writes decoded[i]
HFC_HD void dlmm_element_fwd(int64_t i, const float* x, const float* noise, const float* params, int c, int k, int hw,
                             int type, int straight_through, float* decoded, float* acc_n, float* acc_q) {
  const int px = static_cast<int>(i % hw);
  const int ch = static_cast<int>((i / hw) % c);
  const int64_t img = i / (static_cast<int64_t>(hw) * c);
  Mix m;
  load_mix(params, img * 3 * c * k * hw, c, k, hw, ch, px, m);
  const float lse = logsumexp_logits(m, k);
  const float xv = x[i];
  const float q = floorf(xv + 0.5f);
  *acc_q += dlmm_loglik(q, m, k, type, lse);
  if (noise) *acc_n += dlmm_loglik(xv + noise[i], m, k, type, lse);
  if (decoded) decoded[i] = straight_through ? xv + (q - xv) : q;
}

// gradient of g * L(x + noise) for element i: dx[i] (+ d_decoded[i]) and the 3k parameter gradients
HFC_HD void dlmm_element_bwd(int64_t i, const float* x, const float* noise, const float* params, const float* d_decoded,
                             float g, int c, int k, int hw, int type, float* dx, float* dparams) {
  const int64_t plane = static_cast<int64_t>(c) * k * hw;
  const int px = static_cast<int>(i % hw);
  const int ch = static_cast<int>((i / hw) % c);
  const int64_t img = i / (static_cast<int64_t>(hw) * c);
  const int64_t base = img * 3 * plane;
  Mix m;
  load_mix(params, base, c, k, hw, ch, px, m);
  const float lse = logsumexp_logits(m, k);
  const float v = x[i] + noise[i];
  float d_logit[kMaxMix], d_mu[kMaxMix], d_ls[kMaxMix];
  const float dv = dlmm_grad(v, m, k, type, lse, g, d_logit, d_mu, d_ls);
  for (int j = 0; j < kMaxMix; ++j)
    if (j < k) {
      const int64_t o = base + (static_cast<int64_t>(ch) * k + j) * hw + px;
      dparams[o] = d_logit[j];
      dparams[o + plane] = d_mu[j];
      dparams[o + 2 * plane] = d_ls[j];
    }
  dx[i] = dv + (d_decoded ? d_decoded[i] : 0.f);                   // straight-through latents: d decoded / d x = 1
}

}  // namespace hfc
