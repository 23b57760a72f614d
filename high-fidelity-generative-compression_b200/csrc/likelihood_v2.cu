// Conditional (mean-scale) Gaussian likelihood of the latents, second schedule of `hfc_latent_likelihood`
// (src/hyperprior.py:57-139, src/helpers/maths.py:87-109) -- same arithmetic as latent_likelihood_kernel in
// elementwise.cu, re-organised around what ncu showed about that kernel at the c2 size (1.8 M elements, 17.3 us,
// profiles/r1_ncu_prof_likelihood.json): 199 issued instructions per element = 18.9 k issue cycles per SM sub-partition
// against 33.4 k elapsed, DRAM at 20 % -- issue-bound with a 1.49-wave tail, not HBM-bound.  Two changes:
//   * packed fp32 (FFMA2 / FMUL2 / FADD2, `fma.rn.f32x2`, sm_100+): the four erfc evaluations of an element are two
//     packed pairs (upper / lower CDF argument, each for the quantised and the noisy value), the exponent runs in base 2
//     with log2(e) folded into the polynomial coefficients, and the log-likelihoods are accumulated in log2 units
//     (one multiplication by ln 2 per block);
//   * a balanced persistent grid: 5 blocks of 256 threads per SM, every block owns an equal contiguous slice of the
//     float4 vectors, so all SMs finish together instead of running a half-empty second wave.
// Algorithmic bytes: 16 B read + 4 B written per element (20 B/element).
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"
#include "hfc_ptx.cuh"

namespace hfc {

namespace {

struct f2 { unsigned long long v; };

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 pk1(float x) { return pk(x, x); }
__device__ __forceinline__ void unpk(f2 a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Chebyshev fit of erfc (Numerical Recipes' erfcc, fractional error < 1.2e-7): erfc(z) = t * exp(-z^2 + P(t)),
// t = 1 / (1 + z / 2), z >= 0.  Coefficients of P pre-multiplied by log2(e) so that the exponential is one ex2.
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kC0 = -1.26551223f * kLog2e, kC1 = 1.00002368f * kLog2e, kC2 = 0.37409196f * kLog2e,
                kC3 = 0.09678418f * kLog2e, kC4 = -0.18628806f * kLog2e, kC5 = 0.27886807f * kLog2e,
                kC6 = -1.13520398f * kLog2e, kC7 = 1.48851587f * kLog2e, kC8 = -0.82215223f * kLog2e,
                kC9 = 0.17087277f * kLog2e;

// erfc of two NON-NEGATIVE arguments.
__device__ __forceinline__ f2 erfc2_pos(f2 z) {
  const f2 den = fma2(pk1(0.5f), z, pk1(1.f));
  float d0, d1;
  unpk(den, d0, d1);
  const f2 t = pk(rcp(d0), rcp(d1));
  f2 p = pk1(kC9);
  p = fma2(p, t, pk1(kC8));
  p = fma2(p, t, pk1(kC7));
  p = fma2(p, t, pk1(kC6));
  p = fma2(p, t, pk1(kC5));
  p = fma2(p, t, pk1(kC4));
  p = fma2(p, t, pk1(kC3));
  p = fma2(p, t, pk1(kC2));
  p = fma2(p, t, pk1(kC1));
  p = fma2(p, t, pk1(kC0));
  const f2 e = fma2(mul2(z, pk1(-kLog2e)), z, p);      // (-z^2 + P(t)) * log2(e)
  float e0, e1;
  unpk(e, e0, e1);
  return mul2(t, pk(ex2(e0), ex2(e1)));
}

// One latent element: quantised + noisy likelihood (lo lane = quantised, hi lane = noisy), returns
// (log2(p_q + 1e-9), log2(p_n + 1e-9)) packed and the straight-through value.
__device__ __forceinline__ f2 element(float y, float mu, float sraw, float nz, float lb, float& dec) {
  const float k = rcp(fmaxf(sraw, lb)) * 0.70710678118654752440f;     // 1 / (sqrt(2) * LowerBoundToward(scale))
  const float v = y - mu;
  const float vq = floorf(v + 0.5f);
  dec = (v + (vq - v)) + mu;                                          // quantize_latents_st, hyperprior.py:108-122
  const float dq = fabsf((vq + mu) - mu);
  const float dn = fabsf((y + nz) - mu);
  const f2 d = pk(dq, dn), kk = pk1(k), hk = pk1(0.5f * k);
  // Phi(v) = erfc(-v / sqrt 2) / 2: upper argument (d - .5) k (either sign), lower argument (d + .5) k (>= 0)
  const f2 xu = fma2(d, kk, pk1(-0.5f * k));
  const f2 xl = fma2(d, kk, hk);
  float xu0, xu1;
  unpk(xu, xu0, xu1);
  f2 eu = erfc2_pos(pk(fabsf(xu0), fabsf(xu1)));
  float eu0, eu1;
  unpk(eu, eu0, eu1);
  eu0 = xu0 >= 0.f ? eu0 : 2.f - eu0;
  eu1 = xu1 >= 0.f ? eu1 : 2.f - eu1;
  const f2 el = erfc2_pos(xl);
  // p = (erfc(xu) - erfc(xl)) / 2, LowerBoundToward(p, 1e-9), log(p + 1e-9)
  const f2 p = fma2(el, pk1(-0.5f), mul2(pk(eu0, eu1), pk1(0.5f)));
  float p0, p1;
  unpk(p, p0, p1);
  return pk(lg2(fmaxf(p0, 1e-9f) + 1e-9f), lg2(fmaxf(p1, 1e-9f) + 1e-9f));
}

template <bool HAS_NOISE, bool PREFETCH>
__global__ void __launch_bounds__(256, PREFETCH ? 4 : 5)
latent_likelihood_v2_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                            const float* __restrict__ scale, const float* __restrict__ noise, int64_t count, float lb,
                            float* __restrict__ decoded, double* __restrict__ sums) {
  __shared__ float red[2][8];
  const int64_t nvec = count / 4;
  // equal contiguous slices of the float4 vectors per block
  const int64_t v0 = nvec * blockIdx.x / gridDim.x, v1 = nvec * (blockIdx.x + 1) / gridDim.x;
  f2 acc = pk1(0.f);                                                   // (quantised, noisy) in log2 units
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (PREFETCH) {
    // schedule 3: the loads of vector i + 256 are in flight while vector i is evaluated (register double buffer), so
    // that DRAM does not idle during the arithmetic phase of a thread's 1-3 iterations
    int64_t i = v0 + threadIdx.x;
    float4 yy = zero4, mm = zero4, ss = zero4, nz = zero4;
    if (i < v1) {
      yy = reinterpret_cast<const float4*>(y)[i];
      mm = reinterpret_cast<const float4*>(mean)[i];
      ss = reinterpret_cast<const float4*>(scale)[i];
      if (HAS_NOISE) nz = reinterpret_cast<const float4*>(noise)[i];
    }
    while (i < v1) {
      const int64_t in = i + 256;
      float4 yn = zero4, mn = zero4, sn = zero4, nn = zero4;
      if (in < v1) {
        yn = reinterpret_cast<const float4*>(y)[in];
        mn = reinterpret_cast<const float4*>(mean)[in];
        sn = reinterpret_cast<const float4*>(scale)[in];
        if (HAS_NOISE) nn = reinterpret_cast<const float4*>(noise)[in];
      }
      float4 dd;
      acc = add2(acc, element(yy.x, mm.x, ss.x, nz.x, lb, dd.x));
      acc = add2(acc, element(yy.y, mm.y, ss.y, nz.y, lb, dd.y));
      acc = add2(acc, element(yy.z, mm.z, ss.z, nz.z, lb, dd.z));
      acc = add2(acc, element(yy.w, mm.w, ss.w, nz.w, lb, dd.w));
      if (decoded) reinterpret_cast<float4*>(decoded)[i] = dd;
      yy = yn; mm = mn; ss = sn; nz = nn;
      i = in;
    }
  } else {
    for (int64_t i = v0 + threadIdx.x; i < v1; i += 256) {
      const float4 yy = reinterpret_cast<const float4*>(y)[i];
      const float4 mm = reinterpret_cast<const float4*>(mean)[i];
      const float4 ss = reinterpret_cast<const float4*>(scale)[i];
      float4 nz = zero4;
      if (HAS_NOISE) nz = reinterpret_cast<const float4*>(noise)[i];
      float4 dd;
      acc = add2(acc, element(yy.x, mm.x, ss.x, nz.x, lb, dd.x));
      acc = add2(acc, element(yy.y, mm.y, ss.y, nz.y, lb, dd.y));
      acc = add2(acc, element(yy.z, mm.z, ss.z, nz.z, lb, dd.z));
      acc = add2(acc, element(yy.w, mm.w, ss.w, nz.w, lb, dd.w));
      if (decoded) reinterpret_cast<float4*>(decoded)[i] = dd;
    }
  }
  if (blockIdx.x == 0) {                                              // ragged tail (count % 4 elements)
    const int64_t i = nvec * 4 + threadIdx.x;
    if (i < count) {
      float dd;
      acc = add2(acc, element(y[i], mean[i], scale[i], HAS_NOISE ? noise[i] : 0.f, lb, dd));
      if (decoded) decoded[i] = dd;
    }
  }
  float aq, an;
  unpk(acc, aq, an);
  aq = warp_sum(aq);
  an = warp_sum(an);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = aq; red[1][warp] = an; }
  __syncthreads();
  if (warp == 0) {
    float rq = lane < 8 ? red[0][lane] : 0.f, rn = lane < 8 ? red[1][lane] : 0.f;
    rq = warp_sum(rq);
    rn = warp_sum(rn);
    if (lane == 0) {
      constexpr double kLn2 = 0.69314718055994530942;
      if (HAS_NOISE) atomicAdd(&sums[0], static_cast<double>(rn) * kLn2);
      atomicAdd(&sums[1], static_cast<double>(rq) * kLn2);
    }
  }
}

// Schedule 4: the block's slice of y / mean / scale / noise is fetched by FOUR bulk copies (cp.async.bulk global -> shared,
// completion on one mbarrier) issued by one thread the moment the block starts, instead of 3 x 4 float4 loads per thread
// spread over the thread's iterations: every byte a resident block will ever need is in flight from its first
// instruction (4 blocks x 48 KB per SM), which is what a 36 MB pass that lasts only a few microseconds needs -- at that
// size the kernel is DRAM ramp + drain, and schedule 3 still alternated load bursts and arithmetic (12.6 us in ncu for
// 5.5 us of traffic at peak).  One chunk of kBulkVec float4 per array and block; non-persistent grid.
constexpr int kBulkVec = 768;                       // float4 per array per block: 12 KB x 4 arrays = 48 KB of smem

template <bool HAS_NOISE>
__global__ void __launch_bounds__(256, 4)
latent_likelihood_bulk_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                              const float* __restrict__ scale, const float* __restrict__ noise, int64_t count, float lb,
                              float* __restrict__ decoded, double* __restrict__ sums) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ uint64_t bar;
  __shared__ float red[2][8];
  float4* sy = reinterpret_cast<float4*>(bulk_smem);
  float4* sm = sy + kBulkVec;
  float4* ss = sm + kBulkVec;
  float4* sn = ss + kBulkVec;
  const int64_t nvec = count / 4;
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * kBulkVec;
  const int nv = static_cast<int>(nvec - v0 < kBulkVec ? nvec - v0 : kBulkVec);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && nv > 0) {
    const uint32_t bytes = static_cast<uint32_t>(nv) * 16u;
    mbar_arrive_expect_tx(&bar, bytes * (HAS_NOISE ? 4u : 3u));
    bulk_g2s(sy, reinterpret_cast<const float4*>(y) + v0, bytes, &bar);
    bulk_g2s(sm, reinterpret_cast<const float4*>(mean) + v0, bytes, &bar);
    bulk_g2s(ss, reinterpret_cast<const float4*>(scale) + v0, bytes, &bar);
    if (HAS_NOISE) bulk_g2s(sn, reinterpret_cast<const float4*>(noise) + v0, bytes, &bar);
  }
  f2 acc = pk1(0.f);
  if (nv > 0) {
    mbar_wait(&bar, 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < nv; i += 256) {
      const float4 yy = sy[i], mm = sm[i], sc = ss[i];
      const float4 nz = HAS_NOISE ? sn[i] : zero4;
      float4 dd;
      acc = add2(acc, element(yy.x, mm.x, sc.x, nz.x, lb, dd.x));
      acc = add2(acc, element(yy.y, mm.y, sc.y, nz.y, lb, dd.y));
      acc = add2(acc, element(yy.z, mm.z, sc.z, nz.z, lb, dd.z));
      acc = add2(acc, element(yy.w, mm.w, sc.w, nz.w, lb, dd.w));
      if (decoded) reinterpret_cast<float4*>(decoded)[v0 + i] = dd;
    }
  }
  if (blockIdx.x == 0) {                                              // ragged tail (count % 4 elements)
    const int64_t i = nvec * 4 + threadIdx.x;
    if (i < count) {
      float dd;
      acc = add2(acc, element(y[i], mean[i], scale[i], HAS_NOISE ? noise[i] : 0.f, lb, dd));
      if (decoded) decoded[i] = dd;
    }
  }
  float aq, an;
  unpk(acc, aq, an);
  aq = warp_sum(aq);
  an = warp_sum(an);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = aq; red[1][warp] = an; }
  __syncthreads();
  if (warp == 0) {
    float rq = lane < 8 ? red[0][lane] : 0.f, rn = lane < 8 ? red[1][lane] : 0.f;
    rq = warp_sum(rq);
    rn = warp_sum(rn);
    if (lane == 0) {
      constexpr double kLn2 = 0.69314718055994530942;
      if (HAS_NOISE) atomicAdd(&sums[0], static_cast<double>(rn) * kLn2);
      atomicAdd(&sums[1], static_cast<double>(rq) * kLn2);
    }
  }
}

}  // namespace

// schedule 4 (bulk-staged); needs 16-byte aligned inputs (the caller falls back to schedule 3 otherwise)
int launch_latent_likelihood_bulk(const float* y, const float* mean, const float* scale_raw, const float* noise,
                                  int64_t count, float lb, float* decoded, double* sums, cudaStream_t st) {
  const int64_t nvec = count / 4;
  const int blocks = static_cast<int>(std::max<int64_t>(1, (nvec + kBulkVec - 1) / kBulkVec));
  const size_t smem = static_cast<size_t>(kBulkVec) * 16 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(latent_likelihood_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(latent_likelihood_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    attr_set = true;
  }
  if (noise)
    latent_likelihood_bulk_kernel<true><<<blocks, 256, smem, st>>>(y, mean, scale_raw, noise, count, lb, decoded, sums);
  else
    latent_likelihood_bulk_kernel<false><<<blocks, 256, smem, st>>>(y, mean, scale_raw, noise, count, lb, decoded, sums);
  return HFC_OK;
}

// Gaussian likelihood only (the logistic variant stays on latent_likelihood_kernel).  prefetch: schedule 3.
int launch_latent_likelihood_v2(const float* y, const float* mean, const float* scale_raw, const float* noise,
                                int64_t count, float lb, float* decoded, double* sums, int sms, bool prefetch,
                                cudaStream_t st) {
  const int64_t nvec = count / 4;
  const int per_sm = prefetch ? 4 : 5;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((nvec + 255) / 256, static_cast<int64_t>(sms) * per_sm)));
  if (prefetch) {
    if (noise)
      latent_likelihood_v2_kernel<true, true><<<blocks, 256, 0, st>>>(y, mean, scale_raw, noise, count, lb, decoded, sums);
    else
      latent_likelihood_v2_kernel<false, true><<<blocks, 256, 0, st>>>(y, mean, scale_raw, noise, count, lb, decoded, sums);
  } else {
    if (noise)
      latent_likelihood_v2_kernel<true, false><<<blocks, 256, 0, st>>>(y, mean, scale_raw, noise, count, lb, decoded, sums);
    else
      latent_likelihood_v2_kernel<false, false><<<blocks, 256, 0, st>>>(y, mean, scale_raw, noise, count, lb, decoded, sums);
  }
  return HFC_OK;
}

}  // namespace hfc
