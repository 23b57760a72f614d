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
#include "dlmm_math.cuh"

namespace hfc {

__global__ void __launch_bounds__(256)
dlmm_likelihood_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ params,
                       int n, int c, int k, int hw, int type, int straight_through, float* __restrict__ decoded,
                       double* __restrict__ sums) {
  __shared__ float red[2][8];
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float acc_n = 0.f, acc_q = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    dlmm_element_fwd(i, x, noise, params, c, k, hw, type, straight_through, decoded, &acc_n, &acc_q);
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
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
    dlmm_element_bwd(i, x, noise, params, d_decoded, g, c, k, hw, type, dx, dparams);
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
