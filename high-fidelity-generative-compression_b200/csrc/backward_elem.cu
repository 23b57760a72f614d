// Elementwise / row-wise backward kernels of the HiFIC training step (all HBM- or latency-bound):
//   channelnorm_bwd        ChannelNorm2D (+ReLU) backward: dz, dgamma, dbeta     (normalisation/channel.py:48-59)
//   relu_mask              g * (y > 0) for the bias+ReLU layers of the hyper networks (network/hyper.py:56-63,90-97)
//   latent_likelihood_bwd  d n_bpp / d(y, mean, scale_raw) incl. both LowerBoundToward gates (hyperprior.py:124-139,
//                          helpers/maths.py:87-100)
//   hyperlatent_likelihood_bwd  d n_bpp / d(z, density parameters)                (compression/hyperprior_model.py)
//   lpips_layer_bwd        gradient of the LPIPS feature loss w.r.t. the features of the reconstruction
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace hfc {

// ------------------------------------------------------------------------------------------------
// ChannelNorm backward.  y = act(gamma * xhat + beta), xhat = (z - mu) * r, r = rsqrt(var_unbiased + eps).
// With G = g * act'(.) * gamma:   dz = r * (G - mean(G) - xhat * sum(G * xhat) / (C - 1)).
// One warp per pixel (row in registers), persistent blocks, per-lane dgamma / dbeta accumulators.
// ------------------------------------------------------------------------------------------------
static constexpr int kCnbVec = 8;  // float4 per lane -> up to 1024 channels

template <int GROUP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GROUP lanes per pixel and VEC float4 per lane (capacity GROUP * VEC * 4 channels), persistent blocks.  Narrow layers
// put several pixels in one warp and run many blocks per SM: the kernel is a pure HBM stream (read z, g; write dz)
// and needs tens of KB in flight per SM.  dbias (optional) receives sum_pixels dz = the gradient of the bias of the
// convolution in front of the norm, which saves a separate pass over dz.
template <int VEC, int GROUP, int MINB>
__global__ void __launch_bounds__(256, MINB)
channelnorm_bwd_kernel(const float* __restrict__ z, int ld_z, const float* __restrict__ g, int ld_g,
                       const float* __restrict__ gamma, const float* __restrict__ beta, int c, long long npix,
                       float eps, int act, float* __restrict__ dz, int ld_dz, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dbias, uint16_t* __restrict__ dz_act, int act_cpad,
                       int act_bf16) {
  constexpr int kPix = 32 / GROUP;
  constexpr int kCap = GROUP * VEC * 4;
  __shared__ float s_acc[3][kCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane % GROUP;
  for (int i = threadIdx.x; i < 3 * kCap; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  // gamma / beta live in registers for the narrow variants; the wide ones re-read them (L1 hits) to stay spill-free
  constexpr bool kCacheGB = VEC <= 1;
  constexpr int kGB = kCacheGB ? VEC : 1;
  float4 ag[VEC], ab[VEC], ad[VEC], gmc[kGB], btc[kGB];
#pragma unroll
  for (int i = 0; i < VEC; ++i) ag[i] = ab[i] = ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kGB; ++i) {
    const int ch = (i * GROUP + gl) * 4;
    gmc[i] = btc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kCacheGB && ch < c) {
      gmc[i] = *reinterpret_cast<const float4*>(gamma + ch);
      btc[i] = *reinterpret_cast<const float4*>(beta + ch);
    }
  }
  const float inv_c = 1.f / static_cast<float>(c), inv_c1 = 1.f / static_cast<float>(c - 1);
  const long long stride = static_cast<long long>(gridDim.x) * 8 * kPix;
  // the loop bound is warp-uniform (first pixel of the warp): all lanes take part in the shuffles
  for (long long pix0 = (static_cast<long long>(blockIdx.x) * 8 + warp) * kPix; pix0 < npix; pix0 += stride) {
    const long long pix = pix0 + lane / GROUP;
    const bool live = pix < npix;
    const float* zr = z + pix * ld_z;
    const float* gr = g + pix * ld_g;
    float4 v[VEC], gg[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int ch = (i * GROUP + gl) * 4;
      v[i] = gg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && ch < c) {
        v[i] = *reinterpret_cast<const float4*>(zr + ch);
        gg[i] = *reinterpret_cast<const float4*>(gr + ch);     // issued early: both streams in flight together
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    s = group_sum<GROUP>(s);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int ch = (i * GROUP + gl) * 4;
      if (ch < c) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    q = group_sum<GROUP>(q);
    const float r = rsqrtf(q * inv_c1 + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int ch = (i * GROUP + gl) * 4;
      if (live && ch < c) {
        float4 go = gg[i];
        const float4 gmi = kCacheGB ? gmc[kCacheGB ? i : 0] : __ldg(reinterpret_cast<const float4*>(gamma + ch));
        const float4 bti = kCacheGB ? btc[kCacheGB ? i : 0] : __ldg(reinterpret_cast<const float4*>(beta + ch));
        // xhat overwrites v
        v[i].x = (v[i].x - mean) * r; v[i].y = (v[i].y - mean) * r;
        v[i].z = (v[i].z - mean) * r; v[i].w = (v[i].w - mean) * r;
        if (act == 1) {  // ReLU: the gradient passes where the forward output was positive
          if (!(fmaf(gmi.x, v[i].x, bti.x) > 0.f)) go.x = 0.f;
          if (!(fmaf(gmi.y, v[i].y, bti.y) > 0.f)) go.y = 0.f;
          if (!(fmaf(gmi.z, v[i].z, bti.z) > 0.f)) go.z = 0.f;
          if (!(fmaf(gmi.w, v[i].w, bti.w) > 0.f)) go.w = 0.f;
        }
        ab[i].x += go.x; ab[i].y += go.y; ab[i].z += go.z; ab[i].w += go.w;
        ag[i].x = fmaf(go.x, v[i].x, ag[i].x); ag[i].y = fmaf(go.y, v[i].y, ag[i].y);
        ag[i].z = fmaf(go.z, v[i].z, ag[i].z); ag[i].w = fmaf(go.w, v[i].w, ag[i].w);
        gg[i] = make_float4(go.x * gmi.x, go.y * gmi.y, go.z * gmi.z, go.w * gmi.w);
        s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
        s2 += (gg[i].x * v[i].x + gg[i].y * v[i].y) + (gg[i].z * v[i].z + gg[i].w * v[i].w);
      } else {
        gg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s1 = group_sum<GROUP>(s1) * inv_c;
    s2 = group_sum<GROUP>(s2) * inv_c1;
    float* dr = dz + pix * ld_dz;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int ch = (i * GROUP + gl) * 4;
      if (live && ch < c) {
        float4 o;
        o.x = r * (gg[i].x - s1 - v[i].x * s2); o.y = r * (gg[i].y - s1 - v[i].y * s2);
        o.z = r * (gg[i].z - s1 - v[i].z * s2); o.w = r * (gg[i].w - s1 - v[i].w * s2);
        if (dz) *reinterpret_cast<float4*>(dr + ch) = o;
        if (dz_act) {     // 16-bit copy in the operand layout of the backward GEMMs (border-less NHWC, pitch act_cpad)
          uint2 pk;
          if (act_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          } else {        // fp16 (loss-scaled gradients, see grad.py): overflow becomes inf and is caught by the optimizer
            const float kMax = 65504.f;   // saturate: see grad.py
            const __half2 lo = __floats2half2_rn(fminf(fmaxf(o.x, -kMax), kMax), fminf(fmaxf(o.y, -kMax), kMax));
            const __half2 hi = __floats2half2_rn(fminf(fmaxf(o.z, -kMax), kMax), fminf(fmaxf(o.w, -kMax), kMax));
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          }
          *reinterpret_cast<uint2*>(dz_act + pix * act_cpad + ch) = pk;
        }
        ad[i].x += o.x; ad[i].y += o.y; ad[i].z += o.z; ad[i].w += o.w;
      } else if (live && dz_act && ch < act_cpad) {
        *reinterpret_cast<uint2*>(dz_act + pix * act_cpad + ch) = make_uint2(0u, 0u);   // channel padding
      }
    }
  }
  // block reduction of the per-lane parameter gradients, then one global atomic per channel per block
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int ch = (i * GROUP + gl) * 4;
    if (ch < c) {
      atomicAdd(&s_acc[0][ch + 0], ag[i].x); atomicAdd(&s_acc[0][ch + 1], ag[i].y);
      atomicAdd(&s_acc[0][ch + 2], ag[i].z); atomicAdd(&s_acc[0][ch + 3], ag[i].w);
      atomicAdd(&s_acc[1][ch + 0], ab[i].x); atomicAdd(&s_acc[1][ch + 1], ab[i].y);
      atomicAdd(&s_acc[1][ch + 2], ab[i].z); atomicAdd(&s_acc[1][ch + 3], ab[i].w);
      atomicAdd(&s_acc[2][ch + 0], ad[i].x); atomicAdd(&s_acc[2][ch + 1], ad[i].y);
      atomicAdd(&s_acc[2][ch + 2], ad[i].z); atomicAdd(&s_acc[2][ch + 3], ad[i].w);
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    atomicAdd(dgamma + ch, s_acc[0][ch]);
    atomicAdd(dbeta + ch, s_acc[1][ch]);
    if (dbias) atomicAdd(dbias + ch, s_acc[2][ch]);
  }
}

template <int VEC, int GROUP, int MINB>
static void launch_channelnorm_bwd(const float* z, int ld_z, const float* g, int ld_g, const float* gamma,
                                   const float* beta, int c, long long npix, float eps, int act, float* dz, int ld_dz,
                                   float* dgamma, float* dbeta, float* dbias, uint16_t* dz_act, int act_cpad, int act_bf16, int sms,
                                   cudaStream_t st) {
  const long long per_block = 8LL * (32 / GROUP);
  const long long blocks = std::max<long long>(1, std::min<long long>((npix + per_block - 1) / per_block,
                                                                       static_cast<long long>(sms) * MINB));
  channelnorm_bwd_kernel<VEC, GROUP, MINB><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
      z, ld_z, g, ld_g, gamma, beta, c, npix, eps, act, dz, ld_dz, dgamma, dbeta, dbias, dz_act, act_cpad, act_bf16);
}

// g_out[p][c] = g[p][c] * (y_act[p][c] > 0)   (y_act: bordered NHWC fp16 output of a bias+ReLU conv)
struct ReluMaskParams {
  int32_t n, h, w, c, cpad, pt, pl, pb, pr, ld_g, ld_out;
  float slope;   // 0: ReLU ; 0.2: LeakyReLU(0.2) of the discriminator (sign(y) == sign(pre-activation) for slope > 0)
};
__global__ void __launch_bounds__(256)
relu_mask_kernel(const float* __restrict__ g, const __half* __restrict__ y, float* __restrict__ out,
                 const __grid_constant__ ReluMaskParams p) {
  const long long total = static_cast<long long>(p.n) * p.h * p.w * p.c;
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % p.c);
    const long long pix = i / p.c;
    const int ww = static_cast<int>(pix % p.w);
    const int hh = static_cast<int>((pix / p.w) % p.h);
    const int nn = static_cast<int>(pix / (static_cast<long long>(p.w) * p.h));
    const __half yv = y[((static_cast<size_t>(nn) * Hp + hh + p.pt) * Wp + ww + p.pl) * p.cpad + ch];
    const float gv = g[pix * p.ld_g + ch];
    out[pix * p.ld_out + ch] = __half2float(yv) > 0.f ? gv : gv * p.slope;
  }
}

// ------------------------------------------------------------------------------------------------
// Conditional likelihood backward (noisy branch only: q_bpp is consumed through .item(), losses.py:21).
//   L = coef * sum ln(p + 1e-9),  p = max(p_raw, 1e-9),  p_raw = Phi(u) - Phi(l),
//   u = (.5 - d)/s, l = -(.5 + d)/s, d = |y + noise - mu|, s = max(s_raw, lb)
// dy_out = dyhat_in (straight-through: d yhat / d y = 1, d yhat / d mu = 0) + dL/dy ; dmu = -dL/d(y+noise) ;
// LowerBoundToward gate (maths.py:97-100): the gradient passes where x >= bound or grad < 0.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cdf_f(float v, int type) {
  if (type == 0) return 0.5f * erfcf(v * -0.70710678118654752440f);
  return 1.f / (1.f + expf(-v));
}
__device__ __forceinline__ float pdf_f(float v, int type) {
  if (type == 0) return 0.3989422804014327f * expf(-0.5f * v * v);
  const float sg = 1.f / (1.f + expf(-v));
  return sg * (1.f - sg);
}

__global__ void __launch_bounds__(256)
latent_likelihood_bwd_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                             const float* __restrict__ scale_raw, const float* __restrict__ noise,
                             const float* __restrict__ dyhat, const float* __restrict__ g_nbpp, float coef,
                             int64_t count, float lb, int type, float* __restrict__ dy, float* __restrict__ dmean,
                             float* __restrict__ dscale) {
  const float G = coef * (g_nbpp ? *g_nbpp : 1.f);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float sr = scale_raw[i];
    const float s = fmaxf(sr, lb);
    const float t = (y[i] + noise[i]) - mean[i];
    const float d = fabsf(t);
    const float sgn = t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f);
    const float u = (0.5f - d) / s, l = -(0.5f + d) / s;
    const float p_raw = cdf_f(u, type) - cdf_f(l, type);
    const float p = fmaxf(p_raw, 1e-9f);
    float dp = G / (p + 1e-9f);
    if (!(p_raw >= 1e-9f || dp < 0.f)) dp = 0.f;          // LowerBoundToward on the likelihood
    const float fu = pdf_f(u, type), fl = pdf_f(l, type);
    const float dd = dp * (fl - fu) / s;                   // dL/dd
    float ds = dp * (fl * l - fu * u) / s;                 // dL/ds
    if (!(sr >= lb || ds < 0.f)) ds = 0.f;                 // LowerBoundToward on the scale
    const float dt = dd * sgn;
    dy[i] = (dyhat ? dyhat[i] : 0.f) + dt;
    dmean[i] = -dt;
    dscale[i] = ds;
  }
}

// ------------------------------------------------------------------------------------------------
// Factorized density backward: one block per channel; reverse mode through the 1-3-3-3-1 monotone MLP for
// both CDF evaluations of the noisy hyper-latent; parameter gradients w.r.t. the PACKED (softplus / tanh
// pre-applied) parameters are reduced in shared memory (the chain rule back to H, a, b is 14 080 scalars of
// torch autograd on the host side).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void density_fwd_bwd(float x, const float* __restrict__ pr, float seed, float* gp,
                                                float& dx) {
  // forward with saved pre-activations
  float t0[3], h0[3], t1[3], h1[3], t2[3], h2[3], t3, th;
#pragma unroll
  for (int j = 0; j < 3; ++j) { t0[j] = pr[j] * x + pr[3 + j]; h0[j] = t0[j] + pr[6 + j] * tanhf(t0[j]); }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t1[i] = pr[9 + 3 * i] * h0[0] + pr[9 + 3 * i + 1] * h0[1] + pr[9 + 3 * i + 2] * h0[2] + pr[18 + i];
    h1[i] = t1[i] + pr[21 + i] * tanhf(t1[i]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t2[i] = pr[24 + 3 * i] * h1[0] + pr[24 + 3 * i + 1] * h1[1] + pr[24 + 3 * i + 2] * h1[2] + pr[33 + i];
    h2[i] = t2[i] + pr[36 + i] * tanhf(t2[i]);
  }
  t3 = pr[39] * h2[0] + pr[40] * h2[1] + pr[41] * h2[2] + pr[42];
  th = tanhf(t3);
  // backward: seed = dL/d(out), out = t3 + pr[43] * tanh(t3)
  gp[43] += seed * th;
  float dt3 = seed * (1.f + pr[43] * (1.f - th * th));
  gp[42] += dt3;
  float dh2[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { gp[39 + j] += dt3 * h2[j]; dh2[j] = dt3 * pr[39 + j]; }
  float dh1[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float tt = tanhf(t2[i]);
    gp[36 + i] += dh2[i] * tt;
    const float dt = dh2[i] * (1.f + pr[36 + i] * (1.f - tt * tt));
    gp[33 + i] += dt;
#pragma unroll
    for (int j = 0; j < 3; ++j) { gp[24 + 3 * i + j] += dt * h1[j]; dh1[j] += dt * pr[24 + 3 * i + j]; }
  }
  float dh0[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float tt = tanhf(t1[i]);
    gp[21 + i] += dh1[i] * tt;
    const float dt = dh1[i] * (1.f + pr[21 + i] * (1.f - tt * tt));
    gp[18 + i] += dt;
#pragma unroll
    for (int j = 0; j < 3; ++j) { gp[9 + 3 * i + j] += dt * h0[j]; dh0[j] += dt * pr[9 + 3 * i + j]; }
  }
  dx = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float tt = tanhf(t0[j]);
    gp[6 + j] += dh0[j] * tt;
    const float dt = dh0[j] * (1.f + pr[6 + j] * (1.f - tt * tt));
    gp[3 + j] += dt;
    gp[j] += dt * x;
    dx += dt * pr[j];
  }
}

__device__ __forceinline__ float density_logits_f(float x, const float* __restrict__ pr) {
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { float t = pr[j] * x + pr[3 + j]; h[j] = t + pr[6 + j] * tanhf(t); }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = pr[9 + 3 * i] * h[0] + pr[9 + 3 * i + 1] * h[1] + pr[9 + 3 * i + 2] * h[2] + pr[18 + i];
    g[i] = t + pr[21 + i] * tanhf(t);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = pr[24 + 3 * i] * g[0] + pr[24 + 3 * i + 1] * g[1] + pr[24 + 3 * i + 2] * g[2] + pr[33 + i];
    h[i] = t + pr[36 + i] * tanhf(t);
  }
  float t = pr[39] * h[0] + pr[40] * h[1] + pr[41] * h[2] + pr[42];
  return t + pr[43] * tanhf(t);
}

__global__ void __launch_bounds__(256)
hyperlatent_likelihood_bwd_kernel(const float* __restrict__ z_noisy, const float* __restrict__ dz_in,
                                  const float* __restrict__ params, const float* __restrict__ g_nbpp, float coef,
                                  int32_t n, int32_t c, int32_t hw, float* __restrict__ dz,
                                  float* __restrict__ dparams) {
  __shared__ float s_par[64];
  __shared__ float s_gp[44];
  const int ch = blockIdx.x;
  if (threadIdx.x < 64) s_par[threadIdx.x] = params[ch * 64 + threadIdx.x];
  if (threadIdx.x < 44) s_gp[threadIdx.x] = 0.f;
  __syncthreads();
  const float G = coef * (g_nbpp ? *g_nbpp : 1.f);
  float gp[44];
#pragma unroll
  for (int k = 0; k < 44; ++k) gp[k] = 0.f;
  const int per = n * hw;
  for (int e = threadIdx.x; e < per; e += blockDim.x) {
    const int img = e / hw, px = e % hw;
    const size_t idx = (static_cast<size_t>(img) * c + ch) * hw + px;
    const float x = z_noisy[idx];
    const float u = density_logits_f(x + 0.5f, s_par), l = density_logits_f(x - 0.5f, s_par);
    const float sum = u + l;
    const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);     // detached in the reference
    const float su = 1.f / (1.f + expf(-sgn * u)), sl = 1.f / (1.f + expf(-sgn * l));
    const float diff = su - sl;
    const float p_raw = fabsf(diff);
    const float p = fmaxf(p_raw, 1e-9f);
    float dp = G / (p + 1e-9f);
    if (!(p_raw >= 1e-9f || dp < 0.f)) dp = 0.f;
    const float dabs = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const float du = dp * dabs * su * (1.f - su) * sgn;
    const float dl = -dp * dabs * sl * (1.f - sl) * sgn;
    float dxu, dxl;
    density_fwd_bwd(x + 0.5f, s_par, du, gp, dxu);
    density_fwd_bwd(x - 0.5f, s_par, dl, gp, dxl);
    dz[idx] = (dz_in ? dz_in[idx] : 0.f) + dxu + dxl;
  }
#pragma unroll
  for (int k = 0; k < 44; ++k) {
    const float v = warp_sum(gp[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_gp[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 44) dparams[ch * 64 + threadIdx.x] = s_gp[threadIdx.x];
  else if (threadIdx.x < 64) dparams[ch * 64 + threadIdx.x] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// LPIPS layer backward w.r.t. f1 (features of the reconstruction); f0 (target) carries no gradient.
//   dist = sum_c w_c (a_c - b_c)^2, a = f0/|f0|, b = f1/|f1|;  d dist/d b_c = -2 w_c (a_c - b_c) =: gb_c
//   d dist/d f1 = (gb - b (b . gb)) / |f1| ; scaled by upstream[image] / hw.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lpips_layer_bwd_kernel(const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ lin_w,
                       const float* __restrict__ upstream, int c, int hw, float* __restrict__ df1) {
  const int img = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= hw) return;
  const float* a = f0 + static_cast<size_t>(img) * c * hw + px;
  const float* b = f1 + static_cast<size_t>(img) * c * hw + px;
  float* o = df1 + static_cast<size_t>(img) * c * hw + px;
  float na = 0.f, nb = 0.f;
  for (int k = 0; k < c; ++k) {
    const float x = a[static_cast<size_t>(k) * hw], y = b[static_cast<size_t>(k) * hw];
    na = fmaf(x, x, na);
    nb = fmaf(y, y, nb);
  }
  const float ia = 1.f / sqrtf(na + 1e-10f), ib = 1.f / sqrtf(nb + 1e-10f);
  float dot = 0.f;   // b_hat . gb
  for (int k = 0; k < c; ++k) {
    const float bh = b[static_cast<size_t>(k) * hw] * ib;
    const float gb = -2.f * lin_w[k] * (a[static_cast<size_t>(k) * hw] * ia - bh);
    dot = fmaf(bh, gb, dot);
  }
  const float up = upstream[img] / static_cast<float>(hw);
  // note: |f1| = sqrt(sum f1^2 + eps): d b/d f1 = (I - b_hat b_hat^T) * ib exactly (eps is inside the sqrt)
  for (int k = 0; k < c; ++k) {
    const float bh = b[static_cast<size_t>(k) * hw] * ib;
    const float gb = -2.f * lin_w[k] * (a[static_cast<size_t>(k) * hw] * ia - bh);
    o[static_cast<size_t>(k) * hw] = up * (gb - bh * dot) * ib;
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

extern "C" int hfc_channelnorm_bwd(const float* z, int32_t ld_z, const float* g, int32_t ld_g, const float* gamma,
                                   const float* beta, int32_t c, int64_t npix, float eps, int32_t act, float* dz,
                                   int32_t ld_dz, float* dgamma, float* dbeta, float* dbias, void* dz_act,
                                   int32_t act_cpad, int32_t act_bf16, void* stream) {
  if (!z || !g || !gamma || !beta || (!dz && !dz_act) || !dgamma || !dbeta || npix <= 0)
    return set_error(HFC_ERR_INVALID, "channelnorm_bwd: null pointer or empty input");
  if (dz_act && (act_cpad % 4 != 0 || act_cpad < c || act_cpad > kCnbVec * 128))
    return set_error(HFC_ERR_INVALID, "channelnorm_bwd: act_cpad (%d) must be a multiple of 4 in [c, %d]", act_cpad, kCnbVec * 128);
  const int width = dz_act ? std::max(c, act_cpad) : c;     // the lanes past c zero the operand's channel padding
  if (c % 4 != 0 || c < 4 || c > kCnbVec * 128 || ld_z % 4 != 0 || ld_g % 4 != 0 || ld_dz % 4 != 0)
    return set_error(HFC_ERR_INVALID, "channelnorm_bwd: needs c %% 4 == 0, c <= %d and pitches %% 4 == 0", kCnbVec * 128);
  if (act != HFC_ACT_NONE && act != HFC_ACT_RELU) return set_error(HFC_ERR_INVALID, "channelnorm_bwd: act must be none or relu");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define HFC_CNB(V, G, B) launch_channelnorm_bwd<V, G, B>(z, ld_z, g, ld_g, gamma, beta, c, npix, eps, act, dz, ld_dz, dgamma, \
                                                        dbeta, dbias, reinterpret_cast<uint16_t*>(dz_act), act_cpad, act_bf16, sms, st)
  if (width <= 32) HFC_CNB(1, 8, 4);
  else if (width <= 64) HFC_CNB(1, 16, 4);
  else if (width <= 128) HFC_CNB(1, 32, 4);
  else if (width <= 256) HFC_CNB(2, 32, 3);
  else if (width <= 512) HFC_CNB(4, 32, 2);
  else HFC_CNB(8, 32, 1);
#undef HFC_CNB
  HFC_CHECK_LAUNCH("channelnorm_bwd launch");
  return HFC_OK;
}

extern "C" int hfc_relu_mask(const float* g, int32_t ld_g, const void* y_act, const hfc_act_geom* geom, float slope,
                             float* out, int32_t ld_out, void* stream) {
  if (!g || !y_act || !geom || !out) return set_error(HFC_ERR_INVALID, "relu_mask: null pointer");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  ReluMaskParams p;
  p.n = geom->n; p.h = geom->h; p.w = geom->w; p.c = geom->c; p.cpad = geom->cpad;
  p.pt = geom->pt; p.pl = geom->pl; p.pb = geom->pb; p.pr = geom->pr; p.ld_g = ld_g; p.ld_out = ld_out; p.slope = slope;
  const long long total = static_cast<long long>(p.n) * p.h * p.w * p.c;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, sms * 8LL)));
  relu_mask_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, reinterpret_cast<const __half*>(y_act), out, p);
  HFC_CHECK_LAUNCH("relu_mask launch");
  return HFC_OK;
}

extern "C" int hfc_latent_likelihood_bwd(const float* y, const float* mean, const float* scale_raw, const float* noise,
                                         const float* dyhat, const float* g_nbpp, float coef, int64_t count,
                                         float scale_lower_bound, int32_t likelihood_type, float* dy, float* dmean,
                                         float* dscale, void* stream) {
  if (!y || !mean || !scale_raw || !noise || !dy || !dmean || !dscale || count <= 0)
    return set_error(HFC_ERR_INVALID, "latent_likelihood_bwd: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((count + 255) / 256, sms * 8LL)));
  latent_likelihood_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, mean, scale_raw, noise, dyhat, g_nbpp, coef, count, scale_lower_bound, likelihood_type, dy, dmean, dscale);
  HFC_CHECK_LAUNCH("latent_likelihood_bwd launch");
  return HFC_OK;
}

extern "C" int hfc_hyperlatent_likelihood_bwd(const float* z_noisy, const float* dz_in, const float* params64,
                                              const float* g_nbpp, float coef, int32_t n, int32_t c, int32_t hw,
                                              float* dz, float* dparams64, void* stream) {
  if (!z_noisy || !params64 || !dz || !dparams64 || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "hyperlatent_likelihood_bwd: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  hyperlatent_likelihood_bwd_kernel<<<c, 256, 0, static_cast<cudaStream_t>(stream)>>>(z_noisy, dz_in, params64, g_nbpp,
                                                                                   coef, n, c, hw, dz, dparams64);
  HFC_CHECK_LAUNCH("hyperlatent_likelihood_bwd launch");
  return HFC_OK;
}

extern "C" int hfc_lpips_layer_bwd(const float* f0, const float* f1, const float* lin_w, const float* upstream,
                                   int32_t n, int32_t c, int32_t hw, float* df1, void* stream) {
  if (!f0 || !f1 || !lin_w || !upstream || !df1 || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "lpips_layer_bwd: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  dim3 grid((hw + 255) / 256, n);
  lpips_layer_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(f0, f1, lin_w, upstream, c, hw, df1);
  HFC_CHECK_LAUNCH("lpips_layer_bwd launch");
  return HFC_OK;
}
