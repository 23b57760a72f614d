// Backward-side helpers of the patch discriminator and the GAN loss (HBM-/latency-bound, no GEMMs):
//   disc_input_bwd  adjoint of cat(x, nearest-upsample(ctx)) (src/network/discriminator.py:75-79): splits the
//                   gradient rows of the 15-channel input into d x (NCHW fp32) and the block sums d ctx
//   spectral_bwd    gradient of W = W_orig / sigma(W_orig), sigma = u^T W_orig v with u, v constants
//                   (torch.nn.utils.spectral_norm as used at discriminator.py:46-62):
//                   dW_orig = (dW - <dW, W> u v^T) / sigma
//   gan_grad        d loss / d logits of the non-saturating losses (src/loss/losses.py:30-41)
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

namespace hfc {

struct DiscInBwdParams {
  int32_t n, h, w, cx, cc, scale, ld_g, ld_ctx;
};

// d x: one thread per (n, c, h, w) element, coalesced along w on the store side
__global__ void __launch_bounds__(256)
disc_input_bwd_x_kernel(const float* __restrict__ g, float* __restrict__ dx, const __grid_constant__ DiscInBwdParams p) {
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  const size_t total = static_cast<size_t>(p.n) * p.cx * plane;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i % plane;
    const int c = static_cast<int>((i / plane) % p.cx);
    const size_t nn = i / (plane * p.cx);
    dx[i] = g[(nn * plane + pix) * p.ld_g + c];
  }
}

// d ctx: one block per context pixel; thread t sums channel (t % 16) over a slice of the scale x scale patch
__global__ void __launch_bounds__(256)
disc_input_bwd_ctx_kernel(const float* __restrict__ g, float* __restrict__ dctx, const __grid_constant__ DiscInBwdParams p) {
  __shared__ float s_part[16][17];
  const int cw = p.w / p.scale, chh = p.h / p.scale;
  const int cell = blockIdx.x;
  const int cx0 = cell % cw, cy0 = (cell / cw) % chh, nn = cell / (cw * chh);
  const int ch = threadIdx.x & 15, slice = threadIdx.x >> 4;   // 16 slices
  float acc = 0.f;
  if (ch < p.cc) {
    const int npatch = p.scale * p.scale;
    for (int q = slice; q < npatch; q += 16) {
      const int hh = cy0 * p.scale + q / p.scale, ww = cx0 * p.scale + q % p.scale;
      acc += g[((static_cast<size_t>(nn) * p.h + hh) * p.w + ww) * p.ld_g + p.cx + ch];
    }
  }
  s_part[slice][ch] = acc;
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int s = 0; s < 16; ++s) t += s_part[s][threadIdx.x];
    if (threadIdx.x < p.ld_ctx) dctx[static_cast<size_t>(cell) * p.ld_ctx + threadIdx.x] = threadIdx.x < p.cc ? t : 0.f;
  }
}

// <dW, W_orig> (block partial sums -> atomicAdd into *dot, caller zeroes)
__global__ void __launch_bounds__(256)
sn_dot_kernel(const float* __restrict__ dw, const float* __restrict__ w, long long count, float* __restrict__ dot) {
  __shared__ float s_w[8];
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc = fmaf(dw[i], w[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_w[i];
    atomicAdd(dot, t);
  }
}

// dW_orig[r][c] (+)= (dW[r][c] - <dW, W> u[r] v[c]) * inv_sigma with W = W_orig * inv_sigma, i.e. <dW, W> = dot * inv_sigma
__global__ void __launch_bounds__(256)
sn_apply_kernel(const float* __restrict__ dw, const float* __restrict__ u, const float* __restrict__ v,
                const float* __restrict__ inv_sigma, const float* __restrict__ dot, int rows, int cols, int accumulate,
                float* __restrict__ out) {
  const float is = *inv_sigma;
  const float coef = *dot * is;              // <dW, W>
  const long long count = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    const float val = (dw[i] - coef * u[r] * v[c]) * is;
    out[i] = accumulate ? out[i] + val : val;
  }
}

// mode 0 (generator loss, mean BCE(gen, 1)):       d/d real = 0,                 d/d gen = (sigmoid - 1) * s
// mode 1 (discriminator loss, BCE(real,1)+BCE(gen,0)): d/d real = (sigmoid - 1) * s, d/d gen = sigmoid * s
__global__ void __launch_bounds__(256)
gan_grad_kernel(const float* __restrict__ logits, long long half, int mode, const float* __restrict__ upstream,
                float inv_n, float* __restrict__ out) {
  const float s = (upstream ? *upstream : 1.f) * inv_n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < 2 * half;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float sg = 1.f / (1.f + __expf(-logits[i]));
    float gval;
    if (i < half) gval = mode == 1 ? (sg - 1.f) * s : 0.f;
    else gval = mode == 1 ? sg * s : (sg - 1.f) * s;
    out[i] = gval;
  }
}

static int grid_for(long long work, int sms, int per_sm = 8) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((work + 255) / 256, static_cast<long long>(sms) * per_sm)));
}

}  // namespace hfc

using namespace hfc;

#define HFC_CHECK_LAUNCH(what)                                                              \
  do {                                                                                      \
    cudaError_t e_ = cudaGetLastError();                                                    \
    if (e_ != cudaSuccess) return set_error(HFC_ERR_LAUNCH, what ": %s", cudaGetErrorString(e_)); \
    note_launch();                                                                          \
  } while (0)

extern "C" int hfc_disc_input_bwd(const float* g, int32_t ld_g, int32_t n, int32_t h, int32_t w, int32_t x_channels,
                                  int32_t ctx_channels, int32_t scale, float* dx, float* dctx, int32_t ld_ctx,
                                  void* stream) {
  if (!g || !dctx) return set_error(HFC_ERR_INVALID, "disc_input_bwd: null pointer");
  if (n <= 0 || h <= 0 || w <= 0 || scale <= 0 || h % scale || w % scale || ctx_channels > 16 || ctx_channels <= 0 ||
      x_channels < 0 || ld_g < x_channels + ctx_channels || ld_ctx < ctx_channels || ld_ctx > 16)
    return set_error(HFC_ERR_INVALID, "disc_input_bwd: inconsistent geometry");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  DiscInBwdParams p;
  p.n = n; p.h = h; p.w = w; p.cx = x_channels; p.cc = ctx_channels; p.scale = scale; p.ld_g = ld_g; p.ld_ctx = ld_ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx && x_channels > 0) {
    const long long total = static_cast<long long>(n) * x_channels * h * w;
    disc_input_bwd_x_kernel<<<grid_for(total, sms), 256, 0, st>>>(g, dx, p);
    HFC_CHECK_LAUNCH("disc_input_bwd launch");
  }
  const int cells = n * (h / scale) * (w / scale);
  disc_input_bwd_ctx_kernel<<<cells, 256, 0, st>>>(g, dctx, p);
  HFC_CHECK_LAUNCH("disc_input_bwd launch");
  return HFC_OK;
}

extern "C" int hfc_spectral_bwd(const float* dw, const float* w_orig, const float* u, const float* v,
                                const float* inv_sigma, int32_t rows, int32_t cols, float* workspace1,
                                int32_t accumulate, float* dw_orig, void* stream) {
  if (!dw || !w_orig || !u || !v || !inv_sigma || !workspace1 || !dw_orig || rows <= 0 || cols <= 0)
    return set_error(HFC_ERR_INVALID, "spectral_bwd: null pointer or empty matrix");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(workspace1, 0, sizeof(float), st) != cudaSuccess)
    return set_error(HFC_ERR_LAUNCH, "spectral_bwd: memset failed");
  const long long count = static_cast<long long>(rows) * cols;
  sn_dot_kernel<<<grid_for(count, sms, 2), 256, 0, st>>>(dw, w_orig, count, workspace1);
  HFC_CHECK_LAUNCH("spectral_bwd launch");
  sn_apply_kernel<<<grid_for(count, sms), 256, 0, st>>>(dw, u, v, inv_sigma, workspace1, rows, cols, accumulate, dw_orig);
  HFC_CHECK_LAUNCH("spectral_bwd launch");
  return HFC_OK;
}

extern "C" int hfc_gan_grad(const float* logits, int64_t half_count, int32_t mode, const float* upstream, float* dlogits,
                            void* stream) {
  if (!logits || !dlogits || half_count <= 0 || (mode != 0 && mode != 1))
    return set_error(HFC_ERR_INVALID, "gan_grad: null pointer, empty input or bad mode");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  gan_grad_kernel<<<grid_for(2 * half_count, sms, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, half_count, mode, upstream, 1.f / static_cast<float>(half_count), dlogits);
  HFC_CHECK_LAUNCH("gan_grad launch");
  return HFC_OK;
}
