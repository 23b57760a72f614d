// Adam step over a list of tensors in ONE launch (train.py:287-300 builds torch.optim.Adam(lr) optimizers for the
// amortization parameters, the hyper-latent density and the discriminator; their update is part of every training
// step, T1 in SURVEY.md section 8).  Pure HBM stream: reads p, g, m, v and writes p, m, v = 28 B per parameter.
//   m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// (torch.optim.Adam, amsgrad = False, maximize = False; weight decay adds wd * p to g first.)
#include "hfc_internal.h"

namespace hfc {

static constexpr int kAdamChunk = 256 * 4 * 8;   // elements per block

struct AdamHyper {
  float lr_over_bc1, b1, b2, omb1, omb2, eps, inv_sqrt_bc2, weight_decay;   // omb = 1 - beta, rounded from double like torch
};

// table: 5 x int64 per tensor {p, g, m, v, numel}; blockmap: 2 x int32 per block {tensor, chunk}
__global__ void __launch_bounds__(256)
adam_multi_kernel(const long long* __restrict__ table, const int* __restrict__ blockmap, const AdamHyper h) {
  const int t = blockmap[2 * blockIdx.x], chunk = blockmap[2 * blockIdx.x + 1];
  float* __restrict__ p = reinterpret_cast<float*>(table[5 * t + 0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(table[5 * t + 1]);
  float* __restrict__ m = reinterpret_cast<float*>(table[5 * t + 2]);
  float* __restrict__ v = reinterpret_cast<float*>(table[5 * t + 3]);
  const long long n = table[5 * t + 4];
  const long long base = static_cast<long long>(chunk) * kAdamChunk;
  const long long end = base + kAdamChunk < n ? base + kAdamChunk : n;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](float& pw, float gw, float& mw, float& vw) {
    if (h.weight_decay != 0.f) gw = fmaf(h.weight_decay, pw, gw);
    mw = h.b1 * mw + h.omb1 * gw;              // torch: lerp / mul_ + addcmul_, each rounded separately
    vw = h.b2 * vw + h.omb2 * gw * gw;
    pw -= h.lr_over_bc1 * mw / (sqrtf(vw) * h.inv_sqrt_bc2 + h.eps);
  };
  if (vec) {
    const long long end4 = base + ((end - base) & ~3LL);
    for (long long i = base + threadIdx.x * 4LL; i < end4; i += 256 * 4) {
      float4 pw = *reinterpret_cast<float4*>(p + i), mw = *reinterpret_cast<float4*>(m + i), vw = *reinterpret_cast<float4*>(v + i);
      const float4 gw = *reinterpret_cast<const float4*>(g + i);
      upd(pw.x, gw.x, mw.x, vw.x); upd(pw.y, gw.y, mw.y, vw.y); upd(pw.z, gw.z, mw.z, vw.z); upd(pw.w, gw.w, mw.w, vw.w);
      *reinterpret_cast<float4*>(p + i) = pw; *reinterpret_cast<float4*>(m + i) = mw; *reinterpret_cast<float4*>(v + i) = vw;
    }
    for (long long i = end4 + threadIdx.x; i < end; i += 256) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (long long i = base + threadIdx.x; i < end; i += 256) upd(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace hfc

using namespace hfc;

extern "C" int32_t hfc_adam_chunk(void) { return kAdamChunk; }

extern "C" int hfc_adam_multi(const int64_t* table_dev, const int32_t* blockmap_dev, int32_t n_blocks, double lr, double beta1,
                              double beta2, double eps, double weight_decay, int64_t step, void* stream) {
  if (!table_dev || !blockmap_dev || n_blocks <= 0 || step <= 0)
    return set_error(HFC_ERR_INVALID, "adam_multi: null pointer, no blocks or step < 1");
  AdamHyper h;
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  h.lr_over_bc1 = static_cast<float>(lr / bc1);
  h.b1 = static_cast<float>(beta1); h.b2 = static_cast<float>(beta2); h.eps = static_cast<float>(eps);
  h.omb1 = static_cast<float>(1.0 - beta1);
  h.omb2 = static_cast<float>(1.0 - beta2);
  h.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  h.weight_decay = static_cast<float>(weight_decay);
  adam_multi_kernel<<<n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(table_dev), blockmap_dev, h);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "adam_multi launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
