// HBM-bound kernels of the HiFIC hot path: module-boundary layout conversion (+ChannelNorm),
// the stand-alone ChannelNorm2D for 480/960 channels, and the two hyperprior likelihood kernels.
// All of them are single-pass, vectorised and coalesced; none goes near the tensor cores.
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"
#include "hfc_likelihood.cuh"

#include <cuda_fp16.h>
#include <cstdlib>

#ifndef HFC_LIKELIHOOD_DEFAULT_VARIANT
#define HFC_LIKELIHOOD_DEFAULT_VARIANT 3
#endif

namespace hfc {

// ------------------------------------------------------------------------------------------------
// NCHW fp32 -> bordered NHWC fp16 act buffer (optionally ChannelNorm2D first)
//   reference: module boundaries + Generator.conv_block_init[0] (src/network/generator.py:98-103,
//   src/normalisation/channel.py:48-59) + the ReflectionPad2d of the first conv of each module.
// One block = one image row segment of 32 pixels, all channels staged through shared memory
// (coalesced 128 B reads along W, 16 B writes along C).
// ------------------------------------------------------------------------------------------------
struct ToActParams {
  int32_t n, c, h, w, cpad;
  int32_t pt, pl, pb, pr;
  int32_t reflect, norm;
  float eps;
};

__global__ void __launch_bounds__(256)
nchw_to_act_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, __half* __restrict__ out,
                   const __grid_constant__ ToActParams p) {
  extern __shared__ float tile[];  // [c][33]
  __shared__ float s_mean[32], s_rstd[32];
  const int segs = (p.w + 31) / 32;
  const int seg = blockIdx.x % segs;
  const int hh = (blockIdx.x / segs) % p.h;
  const int nn = blockIdx.x / (segs * p.h);
  const int w0 = seg * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npx = min(32, p.w - w0);

  for (int c = warp; c < p.c; c += 8) {
    float v = 0.f;
    if (lane < npx) v = x[((static_cast<size_t>(nn) * p.c + c) * p.h + hh) * p.w + w0 + lane];
    tile[c * 33 + lane] = v;
  }
  __syncthreads();
  if (p.norm) {
    for (int px = warp; px < 32; px += 8) {
      float s = 0.f;
      for (int c = lane; c < p.c; c += 32) s += tile[c * 33 + px];
      s = warp_sum(s);
      const float mean = s / static_cast<float>(p.c);
      float q = 0.f;
      for (int c = lane; c < p.c; c += 32) {
        const float d = tile[c * 33 + px] - mean;
        q += d * d;
      }
      q = warp_sum(q);
      if (lane == 0) {
        s_mean[px] = mean;
        s_rstd[px] = rsqrtf(q / static_cast<float>(p.c - 1) + p.eps);
      }
    }
    __syncthreads();
  }
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
  const int groups = p.cpad / 8;
  int rows[3];
  const int nr = mirror_targets(hh, p.h, p.pt, p.pb, p.reflect != 0, rows);
  for (int item = threadIdx.x; item < npx * groups; item += blockDim.x) {
    const int px = item / groups;
    const int g = item % groups;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float v = 0.f;
      if (c < p.c) {
        v = tile[c * 33 + px];
        if (p.norm) v = gamma[c] * ((v - s_mean[px]) * s_rstd[px]) + beta[c];
      }
      f[j] = v;
    }
    uint4 pk;
    __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
    __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    int cols[3];
    const int nc = mirror_targets(w0 + px, p.w, p.pl, p.pr, p.reflect != 0, cols);
    for (int ri = 0; ri < nr; ++ri)
      for (int ci = 0; ci < nc; ++ci) {
        __half* dst = out + ((static_cast<size_t>(nn) * Hp + rows[ri]) * Wp + cols[ci]) * p.cpad + g * 8;
        *reinterpret_cast<uint4*>(dst) = pk;
      }
  }
}

// Few-channel variant (the RGB image: c <= 8, cpad == 8): one thread per pixel, channel planes read
// coalesced along W, one 16 B store per (mirrored) target pixel.
__global__ void __launch_bounds__(256)
nchw_small_to_act_kernel(const float* __restrict__ x, __half* __restrict__ out,
                         const __grid_constant__ ToActParams p) {
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  const size_t total = static_cast<size_t>(p.n) * plane;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ww = static_cast<int>(idx % p.w);
  const int hh = static_cast<int>((idx / p.w) % p.h);
  const int nn = static_cast<int>(idx / plane);
  float f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    f[c] = (c < p.c) ? x[(static_cast<size_t>(nn) * p.c + c) * plane + static_cast<size_t>(hh) * p.w + ww] : 0.f;
  uint4 pk;
  __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
  __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
  pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
  int rows[3], cols[3];
  const int nr = mirror_targets(hh, p.h, p.pt, p.pb, p.reflect != 0, rows);
  const int nc = mirror_targets(ww, p.w, p.pl, p.pr, p.reflect != 0, cols);
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
  for (int ri = 0; ri < nr; ++ri)
    for (int ci = 0; ci < nc; ++ci)
      *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(nn) * Hp + rows[ri]) * Wp + cols[ci]) * 8) = pk;
}

// ------------------------------------------------------------------------------------------------
// Stand-alone ChannelNorm2D over NHWC fp32 rows (one warp per pixel, row held in registers)
// ------------------------------------------------------------------------------------------------
struct CnParams {
  int32_t n, c, h, w, cpad, ld;
  int32_t pt, pl, pb, pr;
  int32_t reflect, act;
  float eps;
};

static constexpr int kCnMaxVec = 8;  // 8 float4 per lane -> up to 1024 channels

// sum over the GROUP lanes (a power of two <= 32) that share a pixel; every lane of the warp takes part
template <int GROUP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GROUP lanes per pixel, VEC float4 per lane (capacity GROUP * VEC * 4 channels): narrow layers put several pixels
// in a warp so that every lane carries data and enough bytes are in flight per SM to cover the HBM latency.
// ITER: pixel groups per warp slot, all loaded before the first is reduced (ITER x 16 B in flight per lane).  With one
// group per slot the 60-channel 256 x 256 x 32 layers of the training forward ran 131 072 blocks of one load each at
// ~2 TB/s; the wide variants (VEC > 1) already carry enough bytes per lane and keep ITER = 1.
template <int VEC, int GROUP, int ITER>
__global__ void __launch_bounds__(256)
channelnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ res1,
                   const float* __restrict__ res2, float* __restrict__ out_f32,
                   __half* __restrict__ out_act, const __grid_constant__ CnParams p) {
  constexpr int kPix = 32 / GROUP;                 // pixels per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane % GROUP;
  const size_t npix = static_cast<size_t>(p.n) * p.h * p.w;
  float4 v[ITER][VEC];
  size_t pixs[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const size_t pix = ((static_cast<size_t>(blockIdx.x) * ITER + it) * 8 + warp) * kPix + lane / GROUP;
    pixs[it] = pix;
    const float* row = x + pix * p.ld;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (i * GROUP + gl) * 4;
      v[it][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix < npix && c < p.c) v[it][i] = *reinterpret_cast<const float4*>(row + c);
    }
  }
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const size_t pix = pixs[it];
    const bool live = pix < npix;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += (v[it][i].x + v[it][i].y) + (v[it][i].z + v[it][i].w);   // lanes past c hold zeros
    s = group_sum<GROUP>(s);
    const float mean = s / static_cast<float>(p.c);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (i * GROUP + gl) * 4;
      if (c < p.c) {
        const float a = v[it][i].x - mean, b = v[it][i].y - mean, cc = v[it][i].z - mean, d = v[it][i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    q = group_sum<GROUP>(q);
    if (!live) continue;                           // after the shuffles: every lane of the warp takes part in them
    const float rstd = rsqrtf(q / static_cast<float>(p.c - 1) + p.eps);

    // 32-bit pixel decode (n*h*w < 2^31 is checked by the launcher): 64-bit div / mod cost more than the rest of the
    // thread's work on the narrow layers
    const unsigned pix32 = static_cast<unsigned>(pix);
    const unsigned row32 = pix32 / static_cast<unsigned>(p.w);
    const int ww = static_cast<int>(pix32 - row32 * static_cast<unsigned>(p.w));
    const int nn = static_cast<int>(row32 / static_cast<unsigned>(p.h));
    const int hh = static_cast<int>(row32 - static_cast<unsigned>(nn) * static_cast<unsigned>(p.h));
    int rows[3], cols[3];
    const int nr = mirror_targets(hh, p.h, p.pt, p.pb, p.reflect != 0, rows);
    const int nc = mirror_targets(ww, p.w, p.pl, p.pr, p.reflect != 0, cols);

#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (i * GROUP + gl) * 4;
      if (c >= p.cpad && c >= p.c) continue;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < p.c) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float4 b = *reinterpret_cast<const float4*>(beta + c);
        y.x = act_apply(g.x * ((v[it][i].x - mean) * rstd) + b.x, p.act);
        y.y = act_apply(g.y * ((v[it][i].y - mean) * rstd) + b.y, p.act);
        y.z = act_apply(g.z * ((v[it][i].z - mean) * rstd) + b.z, p.act);
        y.w = act_apply(g.w * ((v[it][i].w - mean) * rstd) + b.w, p.act);
        if (res1) {
          const float4 r = *reinterpret_cast<const float4*>(res1 + pix * p.c + c);
          y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
        }
        if (res2) {
          const float4 r = *reinterpret_cast<const float4*>(res2 + pix * p.c + c);
          y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
        }
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * p.c + c) = y;
      }
      if (out_act && c < p.cpad) {
        __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        for (int ri = 0; ri < nr; ++ri)
          for (int ci = 0; ci < nc; ++ci) {
            __half* dst = out_act + ((static_cast<size_t>(nn) * Hp + rows[ri]) * Wp + cols[ci]) * p.cpad + c;
            *reinterpret_cast<uint2*>(dst) = pk;
          }
      }
    }
  }
}

template <int VEC, int GROUP>
static void launch_channelnorm(const float* x, const float* gamma, const float* beta, const float* res1,
                               const float* res2, float* out_f32, __half* out_act, const CnParams& p,
                               cudaStream_t st) {
  const long long npix = static_cast<long long>(p.n) * p.h * p.w;
  const long long per_block = 8LL * (32 / GROUP);
  static const bool iter_on = [] { const char* e = getenv("HFC_CN_ITER"); return !(e && e[0] == '0'); }();
  if (VEC == 1 && iter_on && npix >= (1LL << 18)) {
    constexpr int kIter = VEC == 1 ? 4 : 1;
    channelnorm_kernel<VEC, GROUP, kIter><<<static_cast<unsigned>((npix + per_block * kIter - 1) / (per_block * kIter)), 256, 0, st>>>(
        x, gamma, beta, res1, res2, out_f32, out_act, p);
    return;
  }
  channelnorm_kernel<VEC, GROUP, 1><<<static_cast<unsigned>((npix + per_block - 1) / per_block), 256, 0, st>>>(
      x, gamma, beta, res1, res2, out_f32, out_act, p);
}

// ------------------------------------------------------------------------------------------------
// Conditional (mean-scale) likelihood of the latents: hyperprior.py:57-139, maths.py:87-109
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_to(float v, float* smem8) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem8[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? smem8[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

__device__ __forceinline__ void latent_one(float y, float mu, float sraw, float nz, bool has_noise,
                                           float lb, int type, float& dec, float& ln, float& lq) {
  const float inv = __fdividef(1.f, fmaxf(sraw, lb));      // 1 / LowerBoundToward(scale)
  // quantised branch: floor(y - mu + .5) + mu  (hyperprior.py:68-71)
  const float v = y - mu;
  const float vq = floorf(v + 0.5f);
  float d = fabsf((vq + mu) - mu);
  float pq = std_cdf((0.5f - d) * inv, type) - std_cdf(-(0.5f + d) * inv, type);
  lq = __logf(fmaxf(pq, 1e-9f) + 1e-9f);
  // straight-through value, same op order as quantize_latents_st (hyperprior.py:108-122)
  dec = (v + (vq - v)) + mu;
  ln = 0.f;
  if (has_noise) {
    d = fabsf((y + nz) - mu);
    float pn = std_cdf((0.5f - d) * inv, type) - std_cdf(-(0.5f + d) * inv, type);
    ln = __logf(fmaxf(pn, 1e-9f) + 1e-9f);
  }
}

__global__ void __launch_bounds__(128)
latent_likelihood_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                         const float* __restrict__ scale, const float* __restrict__ noise,
                         int64_t count, float lb, int type, float* __restrict__ decoded,
                         double* __restrict__ sums) {
  __shared__ float red[8];
  float acc_n = 0.f, acc_q = 0.f;
  const bool has_noise = noise != nullptr;
  const int64_t nvec = count / 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 yy = reinterpret_cast<const float4*>(y)[i];
    const float4 mm = reinterpret_cast<const float4*>(mean)[i];
    const float4 ss = reinterpret_cast<const float4*>(scale)[i];
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_noise) nz = reinterpret_cast<const float4*>(noise)[i];
    float4 dd;
    float ln, lq;
    latent_one(yy.x, mm.x, ss.x, nz.x, has_noise, lb, type, dd.x, ln, lq); acc_n += ln; acc_q += lq;
    latent_one(yy.y, mm.y, ss.y, nz.y, has_noise, lb, type, dd.y, ln, lq); acc_n += ln; acc_q += lq;
    latent_one(yy.z, mm.z, ss.z, nz.z, has_noise, lb, type, dd.z, ln, lq); acc_n += ln; acc_q += lq;
    latent_one(yy.w, mm.w, ss.w, nz.w, has_noise, lb, type, dd.w, ln, lq); acc_n += ln; acc_q += lq;
    if (decoded) reinterpret_cast<float4*>(decoded)[i] = dd;
  }
  // tail
  for (int64_t i = nvec * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += stride) {
    float dd, ln, lq;
    latent_one(y[i], mean[i], scale[i], has_noise ? noise[i] : 0.f, has_noise, lb, type, dd, ln, lq);
    acc_n += ln; acc_q += lq;
    if (decoded) decoded[i] = dd;
  }
  const float bn = block_sum_to(acc_n, red);
  const float bq = block_sum_to(acc_q, red);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], static_cast<double>(bn));
    atomicAdd(&sums[1], static_cast<double>(bq));
  }
}

// ------------------------------------------------------------------------------------------------
// Factorized density of the hyper-latents: hyperprior_model.py:305-326 (cdf_logits), :349-384
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float density_logits(float x, const float* __restrict__ pr) {
  // layer 0: 1 -> 3
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = pr[j] * x + pr[3 + j];
    h[j] = t + pr[6 + j] * tanhf(t);
  }
  // layer 1: 3 -> 3
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = pr[9 + 3 * i] * h[0];
    t += pr[9 + 3 * i + 1] * h[1];
    t += pr[9 + 3 * i + 2] * h[2];
    t += pr[18 + i];
    g[i] = t + pr[21 + i] * tanhf(t);
  }
  // layer 2: 3 -> 3
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = pr[24 + 3 * i] * g[0];
    t += pr[24 + 3 * i + 1] * g[1];
    t += pr[24 + 3 * i + 2] * g[2];
    t += pr[33 + i];
    h[i] = t + pr[36 + i] * tanhf(t);
  }
  // layer 3: 3 -> 1
  float t = pr[39] * h[0];
  t += pr[40] * h[1];
  t += pr[41] * h[2];
  t += pr[42];
  return t + pr[43] * tanhf(t);
}

__device__ __forceinline__ float density_loglik(float x, const float* __restrict__ pr) {
  const float u = density_logits(x + 0.5f, pr);
  const float l = density_logits(x - 0.5f, pr);
  const float sum = u + l;
  const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
  float pv = fabsf(1.f / (1.f + expf(-sgn * u)) - 1.f / (1.f + expf(-sgn * l)));
  pv = fmaxf(pv, 1e-9f);
  return logf(pv + 1e-9f);
}

__global__ void __launch_bounds__(256)
hyperlatent_likelihood_kernel(const float* __restrict__ z, const float* __restrict__ noise,
                              const float* __restrict__ params, int32_t n, int32_t c, int32_t hw,
                              float* __restrict__ z_noisy, float* __restrict__ z_quant,
                              double* __restrict__ sums) {
  __shared__ float red[8];
  float acc_n = 0.f, acc_q = 0.f;
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int ch = static_cast<int>((i / hw) % c);
    const float* pr = params + ch * 64;
    const float zz = z[i];
    const float zq = floorf(zz + 0.5f);
    acc_q += density_loglik(zq, pr);
    if (z_quant) z_quant[i] = zq;
    if (noise) {
      const float zn = zz + noise[i];
      acc_n += density_loglik(zn, pr);
      if (z_noisy) z_noisy[i] = zn;
    }
  }
  const float bn = block_sum_to(acc_n, red);
  const float bq = block_sum_to(acc_q, red);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], static_cast<double>(bn));
    atomicAdd(&sums[1], static_cast<double>(bq));
  }
}

static int check_geom(const hfc_act_geom* g, const char* who) {
  if (!g) return set_error(HFC_ERR_INVALID, "%s: null geometry", who);
  if (g->n <= 0 || g->h <= 0 || g->w <= 0 || g->c <= 0 || g->cpad < g->c || g->cpad % 8 != 0)
    return set_error(HFC_ERR_INVALID, "%s: bad geometry n=%d h=%d w=%d c=%d cpad=%d", who, g->n,
                     g->h, g->w, g->c, g->cpad);
  if (g->pt < 0 || g->pl < 0 || g->pb < 0 || g->pr < 0)
    return set_error(HFC_ERR_INVALID, "%s: negative border", who);
  return HFC_OK;
}

}  // namespace hfc

using namespace hfc;

extern "C" int hfc_nchw_to_act(const float* x, const hfc_act_geom* g, int32_t reflect, int32_t norm,
                               const float* gamma, const float* beta, float eps, void* out,
                               void* stream) {
  int rc = check_geom(g, "nchw_to_act");
  if (rc != HFC_OK) return rc;
  if (!x || !out) return set_error(HFC_ERR_INVALID, "nchw_to_act: null pointer");
  if (!reflect && (g->pt | g->pl | g->pb | g->pr))
    return set_error(HFC_ERR_INVALID, "nchw_to_act: a border needs reflect=1 (zero pad is TMA OOB fill)");
  if (reflect && (g->pt >= g->h || g->pb >= g->h || g->pl >= g->w || g->pr >= g->w))
    return set_error(HFC_ERR_INVALID, "nchw_to_act: reflected border wider than the image");
  if (norm && (!gamma || !beta || g->c < 2))
    return set_error(HFC_ERR_INVALID, "nchw_to_act: norm needs gamma, beta and c >= 2");
  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  ToActParams p;
  p.n = g->n; p.c = g->c; p.h = g->h; p.w = g->w; p.cpad = g->cpad;
  p.pt = g->pt; p.pl = g->pl; p.pb = g->pb; p.pr = g->pr;
  p.reflect = reflect; p.norm = norm; p.eps = eps;
  if (g->cpad == 8 && !norm) {
    const long long npix = static_cast<long long>(g->n) * g->h * g->w;
    nchw_small_to_act_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0,
                               static_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<__half*>(out), p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "nchw_to_act launch: %s", cudaGetErrorString(e));
    note_launch();
    return HFC_OK;
  }
  const size_t smem = static_cast<size_t>(g->c) * 33 * sizeof(float);
  if (smem > 200 * 1024) return set_error(HFC_ERR_INVALID, "nchw_to_act: too many channels (%d)", g->c);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nchw_to_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int segs = (g->w + 31) / 32;
  const long long blocks = static_cast<long long>(g->n) * g->h * segs;
  nchw_to_act_kernel<<<static_cast<unsigned>(blocks), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      x, gamma, beta, reinterpret_cast<__half*>(out), p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "nchw_to_act launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

extern "C" int hfc_channelnorm(const float* x, int32_t ld, const hfc_act_geom* g, int32_t reflect,
                               const float* gamma, const float* beta, float eps, int32_t act,
                               const float* res1, const float* res2, float* out_f32, void* out_act,
                               void* stream) {
  int rc = check_geom(g, "channelnorm");
  if (rc != HFC_OK) return rc;
  if (!x || !gamma || !beta) return set_error(HFC_ERR_INVALID, "channelnorm: null pointer");
  if (res2 && !res1) { res1 = res2; res2 = nullptr; }
  if (g->c % 4 != 0 || g->c < 4 || g->cpad > kCnMaxVec * 128 || ld % 4 != 0 || ld < g->c)
    return set_error(HFC_ERR_INVALID, "channelnorm: needs c %% 4 == 0, cpad <= %d, ld %% 4 == 0",
                     kCnMaxVec * 128);
  if (!reflect && (g->pt | g->pl | g->pb | g->pr) && out_act)
    return set_error(HFC_ERR_INVALID, "channelnorm: a border needs reflect=1");
  if (reflect && (g->pt >= g->h || g->pb >= g->h || g->pl >= g->w || g->pr >= g->w))
    return set_error(HFC_ERR_INVALID, "channelnorm: reflected border wider than the image");
  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  CnParams p;
  p.n = g->n; p.c = g->c; p.h = g->h; p.w = g->w; p.cpad = g->cpad; p.ld = ld;
  p.pt = g->pt; p.pl = g->pl; p.pb = g->pb; p.pr = g->pr;
  p.reflect = reflect; p.act = act; p.eps = eps;
  if (static_cast<long long>(g->n) * g->h * g->w >= (1LL << 31))
    return set_error(HFC_ERR_UNSUPPORTED, "channelnorm: more than 2^31 pixels");
  // the padded channels (cpad > c) are zero-filled by the lanes past c: capacity must cover cpad
  const int width = std::max(g->c, out_act ? g->cpad : g->c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* oa = reinterpret_cast<__half*>(out_act);
  if (width <= 32) launch_channelnorm<1, 8>(x, gamma, beta, res1, res2, out_f32, oa, p, st);
  else if (width <= 64) launch_channelnorm<1, 16>(x, gamma, beta, res1, res2, out_f32, oa, p, st);
  else if (width <= 128) launch_channelnorm<1, 32>(x, gamma, beta, res1, res2, out_f32, oa, p, st);
  else if (width <= 256) launch_channelnorm<2, 32>(x, gamma, beta, res1, res2, out_f32, oa, p, st);
  else if (width <= 512) launch_channelnorm<4, 32>(x, gamma, beta, res1, res2, out_f32, oa, p, st);
  else launch_channelnorm<8, 32>(x, gamma, beta, res1, res2, out_f32, oa, p, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "channelnorm launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

static int likelihood_variant() {
  const char* v = getenv("HFC_LIKELIHOOD_V");
  return (v && v[0] >= '1' && v[0] <= '4') ? v[0] - '0' : HFC_LIKELIHOOD_DEFAULT_VARIANT;
}

extern "C" int hfc_latent_likelihood(const float* y, const float* mean, const float* scale_raw,
                                     const float* noise, int64_t count, float scale_lower_bound,
                                     int32_t likelihood_type, float* decoded, double* sums,
                                     void* stream) {
  if (!y || !mean || !scale_raw || !sums || count <= 0)
    return set_error(HFC_ERR_INVALID, "latent_likelihood: null pointer or empty input");
  if (likelihood_type != 0 && likelihood_type != 1)
    return set_error(HFC_ERR_INVALID, "latent_likelihood: likelihood_type must be 0 or 1");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  // Gaussian likelihood: schedule 3 by default (likelihood_v2.cu: packed fp32 + balanced persistent grid + register-
  // double-buffered loads) -- measured at c2 / c5 sizes with a cold L2 (profiles/r01_likelihood_ab.json): schedule 1
  // 20.5 / 49.2 us, schedule 2 16.4 / 36.9 us, schedule 3 14.3 / 35.8 us.  HFC_LIKELIHOOD_V=1|2|3 selects one explicitly.
  const int variant = likelihood_variant();
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(scale_raw) |
                           reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(decoded)) & 15) == 0;
  if (likelihood_type == 0 && variant == 4 && aligned16) {
    rc = launch_latent_likelihood_bulk(y, mean, scale_raw, noise, count, scale_lower_bound, decoded, sums,
                                       static_cast<cudaStream_t>(stream));
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "latent_likelihood (bulk) launch: %s", cudaGetErrorString(e2));
    note_launch();
    return rc;
  }
  if (likelihood_type == 0 && variant >= 2) {
    rc = launch_latent_likelihood_v2(y, mean, scale_raw, noise, count, scale_lower_bound, decoded, sums, sms,
                                     variant >= 3, static_cast<cudaStream_t>(stream));
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "latent_likelihood (v2) launch: %s", cudaGetErrorString(e2));
    note_launch();
    return rc;
  }
  // one float4 per thread up to 16 resident blocks of 128 threads per SM, grid-stride beyond that
  const long long want = (count / 4 + 127) / 128;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(want, sms * 32LL)));
  latent_likelihood_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      y, mean, scale_raw, noise, count, scale_lower_bound, likelihood_type, decoded, sums);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "latent_likelihood launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

extern "C" int hfc_hyperlatent_likelihood(const float* z, const float* noise, const float* params64,
                                          int32_t n, int32_t c, int32_t hw, float* z_noisy,
                                          float* z_quant, double* sums, void* stream) {
  if (!z || !params64 || !sums || n <= 0 || c <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "hyperlatent_likelihood: null pointer or empty input");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const long long count = static_cast<long long>(n) * c * hw;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((count + 255) / 256, sms * 8LL)));
  hyperlatent_likelihood_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, noise, params64, n, c, hw, z_noisy, z_quant, sums);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "hyperlatent_likelihood launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
