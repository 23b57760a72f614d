// InstanceNorm2d variant of the inter-layer normalisation (use_channel_norm = False):
//   src/normalisation/instance.py:7-15 wraps torch.nn.InstanceNorm2d(affine=True, track_running_stats=False), used by
//   src/network/encoder.py:41-44 and src/network/generator.py:21-24, 81-84 in place of ChannelNorm2D.
// Statistics are per (image, channel) over the h*w pixels (biased variance, eps 1e-5), so unlike ChannelNorm the
// reduction runs ACROSS the rows of the NHWC activation: every pass below walks the rows with 32 x 8 thread tiles
// (32 float4 lanes = 128 consecutive channels of a row, 8 rows at a time: 512 B coalesced segments), reduces the 8 row
// lanes through shared memory and adds the tile's partial sums into double-precision accumulators.
//   forward : shifted sums of z -> (mean, rstd) -> y = act(gamma * (z - mean) * rstd + beta) [+ res1] [+ res2]
//   backward: (mean, rstd) again from the saved z; s1 = sum G, s2 = sum G * xhat with G = g * act'(.);
//             dz = rstd * gamma * (G - s1 / HW - xhat * s2 / HW), dgamma = sum_n s2, dbeta = sum_n s1, dbias = sum dz
// HBM-bound: forward reads z twice and writes y once; backward reads z three times and g twice.
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace hfc {

struct InParams {
  int32_t n, hw, w, h, c, cpad, ld;     // ld: pitch of z (floats)
  int32_t strip;                        // rows (pixels) per block, a multiple of 8
  int32_t pt, pl, pb, pr, reflect, act;
  float eps, inv_hw;
};

static constexpr int kInLanes = 32;     // float4 lanes: 128 channels per block column
static constexpr int kInRows = 8;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// Adds the block's per-channel partial sums (reduced over the 8 row lanes through shared memory) into acc0 / acc1.
__device__ __forceinline__ void tile_reduce_to(float4 a, float4 b, double* acc0, double* acc1, int ch, int c) {
  __shared__ float4 s_red[2][kInRows][kInLanes];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  s_red[0][ty][tx] = a;
  s_red[1][ty][tx] = b;
  __syncthreads();
  if (ty < 2 && ch < c) {                // row lane 0 finishes sum a, row lane 1 sum b
    float4 t = s_red[ty][0][tx];
#pragma unroll
    for (int r = 1; r < kInRows; ++r) {
      const float4 u = s_red[ty][r][tx];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    double* dst = (ty == 0 ? acc0 : acc1) + ch;
    atomicAdd(dst + 0, static_cast<double>(t.x)); atomicAdd(dst + 1, static_cast<double>(t.y));
    atomicAdd(dst + 2, static_cast<double>(t.z)); atomicAdd(dst + 3, static_cast<double>(t.w));
  }
}

// acc[0][n][c] += sum (z - K), acc[1][n][c] += sum (z - K)^2 with K = z at the image's first pixel: the shift keeps
// the one-pass variance free of cancellation when |mean| >> std.
__global__ void __launch_bounds__(256)
instnorm_stats_kernel(const float* __restrict__ z, double* __restrict__ acc, const __grid_constant__ InParams p) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nn = blockIdx.z, ch = (blockIdx.y * kInLanes + tx) * 4;
  const float* img = z + static_cast<size_t>(nn) * p.hw * p.ld;
  float4 a = zero4(), b = zero4();
  if (ch < p.c) {
    const float4 k = ld4(img + ch);
    const int p0 = blockIdx.x * p.strip, p1 = min(p0 + p.strip, p.hw);
    for (int px = p0 + ty; px < p1; px += kInRows) {
      const float4 v = ld4(img + static_cast<size_t>(px) * p.ld + ch);
      const float dx = v.x - k.x, dy = v.y - k.y, dz = v.z - k.z, dw = v.w - k.w;
      a.x += dx; a.y += dy; a.z += dz; a.w += dw;
      b.x = fmaf(dx, dx, b.x); b.y = fmaf(dy, dy, b.y); b.z = fmaf(dz, dz, b.z); b.w = fmaf(dw, dw, b.w);
    }
  }
  const size_t nc = static_cast<size_t>(p.n) * p.c;
  tile_reduce_to(a, b, acc + static_cast<size_t>(nn) * p.c, acc + nc + static_cast<size_t>(nn) * p.c, ch, p.c);
}

// (mean, rstd)[n][c] from the shifted sums (biased variance, as torch.nn.functional.instance_norm)
__global__ void __launch_bounds__(256)
instnorm_finalize_stats_kernel(const float* __restrict__ z, const double* __restrict__ acc, float* __restrict__ mean,
                               float* __restrict__ rstd, const __grid_constant__ InParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = p.n * p.c;
  if (i >= nc) return;
  const int nn = i / p.c, ch = i - nn * p.c;
  const double k = static_cast<double>(z[static_cast<size_t>(nn) * p.hw * p.ld + ch]);
  const double m = acc[i] / p.hw;
  const double var = fmax(acc[nc + i] / p.hw - m * m, 0.0);
  mean[i] = static_cast<float>(k + m);
  rstd[i] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
}

__global__ void __launch_bounds__(256)
instnorm_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ res1,
                      const float* __restrict__ res2, float* __restrict__ out_f32, __half* __restrict__ out_act,
                      const __grid_constant__ InParams p) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nn = blockIdx.z, ch = (blockIdx.y * kInLanes + tx) * 4;
  const bool real = ch < p.c;
  if (!real && !(out_act && ch < p.cpad)) return;
  float4 m = zero4(), r = zero4(), g = zero4(), b = zero4();
  if (real) {
    m = ld4(mean + static_cast<size_t>(nn) * p.c + ch);
    r = ld4(rstd + static_cast<size_t>(nn) * p.c + ch);
    g = ld4(gamma + ch);
    b = ld4(beta + ch);
  }
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
  const int p0 = blockIdx.x * p.strip, p1 = min(p0 + p.strip, p.hw);
  for (int px = p0 + ty; px < p1; px += kInRows) {
    const size_t pix = static_cast<size_t>(nn) * p.hw + px;
    float4 y = zero4();
    if (real) {
      const float4 v = ld4(z + pix * p.ld + ch);
      y.x = act_apply(g.x * ((v.x - m.x) * r.x) + b.x, p.act); y.y = act_apply(g.y * ((v.y - m.y) * r.y) + b.y, p.act);
      y.z = act_apply(g.z * ((v.z - m.z) * r.z) + b.z, p.act); y.w = act_apply(g.w * ((v.w - m.w) * r.w) + b.w, p.act);
      if (res1) { const float4 r = ld4(res1 + pix * p.c + ch); y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w; }
      if (res2) { const float4 r = ld4(res2 + pix * p.c + ch); y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w; }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * p.c + ch) = y;
    }
    if (out_act) {
      const __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      const int hh = px / p.w, ww = px - hh * p.w;
      int rows[3], cols[3];
      const int nr = mirror_targets(hh, p.h, p.pt, p.pb, p.reflect != 0, rows);
      const int ncol = mirror_targets(ww, p.w, p.pl, p.pr, p.reflect != 0, cols);
      for (int ri = 0; ri < nr; ++ri)
        for (int ci = 0; ci < ncol; ++ci)
          *reinterpret_cast<uint2*>(out_act + ((static_cast<size_t>(nn) * Hp + rows[ri]) * Wp + cols[ci]) * p.cpad + ch) = pk;
    }
  }
}

// G = g * act'(gamma * xhat + beta): the gradient passes where the forward output was positive (same rule as
// channelnorm_bwd_kernel)
__device__ __forceinline__ float4 masked_grad(float4 g, float4 xh, float4 gm, float4 bt, int act) {
  if (act == 1) {
    if (!(fmaf(gm.x, xh.x, bt.x) > 0.f)) g.x = 0.f;
    if (!(fmaf(gm.y, xh.y, bt.y) > 0.f)) g.y = 0.f;
    if (!(fmaf(gm.z, xh.z, bt.z) > 0.f)) g.z = 0.f;
    if (!(fmaf(gm.w, xh.w, bt.w) > 0.f)) g.w = 0.f;
  }
  return g;
}

// acc[0][n][c] += sum G, acc[1][n][c] += sum G * xhat
__global__ void __launch_bounds__(256)
instnorm_bwd_sums_kernel(const float* __restrict__ z, const float* __restrict__ g, int ld_g, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                         double* __restrict__ acc, const __grid_constant__ InParams p) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nn = blockIdx.z, ch = (blockIdx.y * kInLanes + tx) * 4;
  float4 a = zero4(), b = zero4();
  if (ch < p.c) {
    const float4 m = ld4(mean + static_cast<size_t>(nn) * p.c + ch), r = ld4(rstd + static_cast<size_t>(nn) * p.c + ch);
    const float4 gm = ld4(gamma + ch), bt = ld4(beta + ch);
    const int p0 = blockIdx.x * p.strip, p1 = min(p0 + p.strip, p.hw);
    for (int px = p0 + ty; px < p1; px += kInRows) {
      const size_t pix = static_cast<size_t>(nn) * p.hw + px;
      const float4 v = ld4(z + pix * p.ld + ch);
      const float4 xh = make_float4((v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w);
      const float4 go = masked_grad(ld4(g + pix * ld_g + ch), xh, gm, bt, p.act);
      a.x += go.x; a.y += go.y; a.z += go.z; a.w += go.w;
      b.x = fmaf(go.x, xh.x, b.x); b.y = fmaf(go.y, xh.y, b.y); b.z = fmaf(go.z, xh.z, b.z); b.w = fmaf(go.w, xh.w, b.w);
    }
  }
  const size_t nc = static_cast<size_t>(p.n) * p.c;
  tile_reduce_to(a, b, acc + static_cast<size_t>(nn) * p.c, acc + nc + static_cast<size_t>(nn) * p.c, ch, p.c);
}

// m1 = s1 / HW, m2 = s2 / HW per (n, c); dgamma[c] += sum_n s2, dbeta[c] += sum_n s1 (one thread per channel)
__global__ void __launch_bounds__(256)
instnorm_bwd_finalize_kernel(const double* __restrict__ acc, float* __restrict__ m1, float* __restrict__ m2,
                             float* __restrict__ dgamma, float* __restrict__ dbeta, const __grid_constant__ InParams p) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= p.c) return;
  const int nc = p.n * p.c;
  double sg = 0.0, sb = 0.0;
  for (int nn = 0; nn < p.n; ++nn) {
    const double s1 = acc[nn * p.c + ch], s2 = acc[nc + nn * p.c + ch];
    sb += s1;
    sg += s2;
    m1[nn * p.c + ch] = static_cast<float>(s1 / p.hw);
    m2[nn * p.c + ch] = static_cast<float>(s2 / p.hw);
  }
  dgamma[ch] += static_cast<float>(sg);
  dbeta[ch] += static_cast<float>(sb);
}

__global__ void __launch_bounds__(256)
instnorm_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ g, int ld_g, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ m1, const float* __restrict__ m2,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dz, int ld_dz,
                          float* __restrict__ dbias, uint16_t* __restrict__ dz_act, int act_cpad, int act_bf16,
                          const __grid_constant__ InParams p) {
  __shared__ float4 s_db[kInRows][kInLanes];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nn = blockIdx.z, ch = (blockIdx.y * kInLanes + tx) * 4;
  const bool real = ch < p.c;
  const bool padlane = !real && dz_act && ch < act_cpad;
  float4 ad = zero4();
  const int p0 = blockIdx.x * p.strip, p1 = min(p0 + p.strip, p.hw);
  if (real) {
    const size_t o = static_cast<size_t>(nn) * p.c + ch;
    const float4 m = ld4(mean + o), r = ld4(rstd + o), a1 = ld4(m1 + o), a2 = ld4(m2 + o);
    const float4 gm = ld4(gamma + ch), bt = ld4(beta + ch);
    const float4 rg = make_float4(r.x * gm.x, r.y * gm.y, r.z * gm.z, r.w * gm.w);
    for (int px = p0 + ty; px < p1; px += kInRows) {
      const size_t pix = static_cast<size_t>(nn) * p.hw + px;
      const float4 v = ld4(z + pix * p.ld + ch);
      const float4 xh = make_float4((v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w);
      const float4 go = masked_grad(ld4(g + pix * ld_g + ch), xh, gm, bt, p.act);
      float4 o4;
      o4.x = rg.x * (go.x - a1.x - xh.x * a2.x); o4.y = rg.y * (go.y - a1.y - xh.y * a2.y);
      o4.z = rg.z * (go.z - a1.z - xh.z * a2.z); o4.w = rg.w * (go.w - a1.w - xh.w * a2.w);
      if (dz) *reinterpret_cast<float4*>(dz + pix * ld_dz + ch) = o4;
      if (dz_act) {     // 16-bit operand of the backward GEMMs (border-less NHWC, pitch act_cpad), as channelnorm_bwd
        uint2 pk;
        if (act_bf16) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(o4.x, o4.y), hi = __floats2bfloat162_rn(o4.z, o4.w);
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        } else {
          const float kMax = 65504.f;   // saturate: see grad.py
          const __half2 lo = __floats2half2_rn(fminf(fmaxf(o4.x, -kMax), kMax), fminf(fmaxf(o4.y, -kMax), kMax));
          const __half2 hi = __floats2half2_rn(fminf(fmaxf(o4.z, -kMax), kMax), fminf(fmaxf(o4.w, -kMax), kMax));
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        }
        *reinterpret_cast<uint2*>(dz_act + pix * act_cpad + ch) = pk;
      }
      ad.x += o4.x; ad.y += o4.y; ad.z += o4.z; ad.w += o4.w;
    }
  } else if (padlane) {
    for (int px = p0 + ty; px < p1; px += kInRows)
      *reinterpret_cast<uint2*>(dz_act + (static_cast<size_t>(nn) * p.hw + px) * act_cpad + ch) = make_uint2(0u, 0u);
  }
  if (dbias) {          // block-uniform
    s_db[ty][tx] = ad;
    __syncthreads();
    if (ty == 0 && real) {
      float4 t = s_db[0][tx];
#pragma unroll
      for (int r = 1; r < kInRows; ++r) { const float4 u = s_db[r][tx]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
      atomicAdd(dbias + ch + 0, t.x); atomicAdd(dbias + ch + 1, t.y);
      atomicAdd(dbias + ch + 2, t.z); atomicAdd(dbias + ch + 3, t.w);
    }
  }
}

static int strip_for(int n, int hw, int chunks, int sms) {
  // enough blocks to fill the machine a few times over, at least 8 rows and at most 512 rows per block
  int strip = 512;
  while (strip > 8 && static_cast<long long>(n) * chunks * ((hw + strip - 1) / strip) < 4LL * sms) strip >>= 1;
  return strip;
}

struct InWs {
  double* acc_a;   // [2][n][c]
  double* acc_b;   // [2][n][c]
  float *mean, *rstd, *m1, *m2;   // [n][c] each
};

static InWs carve(void* ws, int n, int c) {
  const size_t nc = static_cast<size_t>(n) * c;
  InWs w;
  w.acc_a = reinterpret_cast<double*>(ws);
  w.acc_b = w.acc_a + 2 * nc;
  w.mean = reinterpret_cast<float*>(w.acc_b + 2 * nc);
  w.rstd = w.mean + nc;
  w.m1 = w.rstd + nc;
  w.m2 = w.m1 + nc;
  return w;
}

static int launch_stats(const float* z, const InParams& p, const InWs& w, int chunks_c, cudaStream_t st) {
  const size_t nc = static_cast<size_t>(p.n) * p.c;
  cudaError_t e = cudaMemsetAsync(w.acc_a, 0, 4 * nc * sizeof(double), st);     // acc_a and acc_b are adjacent
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "instancenorm: memset: %s", cudaGetErrorString(e));
  const dim3 grid((p.hw + p.strip - 1) / p.strip, chunks_c, p.n);
  instnorm_stats_kernel<<<grid, 256, 0, st>>>(z, w.acc_a, p);
  note_launch();
  instnorm_finalize_stats_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, st>>>(z, w.acc_a, w.mean, w.rstd, p);
  note_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "instancenorm stats launch: %s", cudaGetErrorString(e));
  return HFC_OK;
}

}  // namespace hfc

using namespace hfc;

extern "C" int64_t hfc_instancenorm_ws_bytes(int32_t n, int32_t c) {
  if (n <= 0 || c <= 0) return 0;
  return static_cast<int64_t>(n) * c * (4 * sizeof(double) + 4 * sizeof(float));
}

extern "C" int hfc_instancenorm(const float* x, int32_t ld, const hfc_act_geom* g, int32_t reflect, const float* gamma,
                                const float* beta, float eps, int32_t act, const float* res1, const float* res2,
                                float* out_f32, void* out_act, void* ws, int64_t ws_bytes, void* stream) {
  if (!g) return set_error(HFC_ERR_INVALID, "instancenorm: null geometry");
  if (g->n <= 0 || g->h <= 0 || g->w <= 0 || g->c <= 0 || g->cpad < g->c || g->cpad % 8 != 0 || g->pt < 0 || g->pl < 0 ||
      g->pb < 0 || g->pr < 0)
    return set_error(HFC_ERR_INVALID, "instancenorm: bad geometry n=%d h=%d w=%d c=%d cpad=%d", g->n, g->h, g->w, g->c, g->cpad);
  if (!x || !gamma || !beta || (!out_f32 && !out_act)) return set_error(HFC_ERR_INVALID, "instancenorm: null pointer");
  if (g->c % 4 != 0 || ld % 4 != 0 || ld < g->c) return set_error(HFC_ERR_INVALID, "instancenorm: needs c %% 4 == 0 and ld %% 4 == 0, ld >= c");
  if (!reflect && (g->pt | g->pl | g->pb | g->pr) && out_act)
    return set_error(HFC_ERR_INVALID, "instancenorm: a border needs reflect=1");
  if (reflect && (g->pt >= g->h || g->pb >= g->h || g->pl >= g->w || g->pr >= g->w))
    return set_error(HFC_ERR_INVALID, "instancenorm: reflected border wider than the image");
  if (static_cast<long long>(g->h) * g->w >= (1LL << 31) || g->n > 65535)
    return set_error(HFC_ERR_UNSUPPORTED, "instancenorm: more than 2^31 pixels per image or more than 65535 images");
  if (!ws || ws_bytes < hfc_instancenorm_ws_bytes(g->n, g->c))
    return set_error(HFC_ERR_INVALID, "instancenorm: workspace too small (hfc_instancenorm_ws_bytes)");
  if (res2 && !res1) { res1 = res2; res2 = nullptr; }
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  InParams p{};
  p.n = g->n; p.h = g->h; p.w = g->w; p.hw = g->h * g->w; p.c = g->c; p.cpad = g->cpad; p.ld = ld;
  p.pt = g->pt; p.pl = g->pl; p.pb = g->pb; p.pr = g->pr; p.reflect = reflect; p.act = act;
  p.eps = eps; p.inv_hw = 1.f / static_cast<float>(p.hw);
  const int chunks_c = (p.c + 127) / 128;
  const int chunks_w = (std::max(p.c, out_act ? p.cpad : p.c) + 127) / 128;
  p.strip = strip_for(p.n, p.hw, chunks_c, sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const InWs w = carve(ws, p.n, p.c);
  rc = launch_stats(x, p, w, chunks_c, st);
  if (rc != HFC_OK) return rc;
  const dim3 grid((p.hw + p.strip - 1) / p.strip, chunks_w, p.n);
  instnorm_apply_kernel<<<grid, 256, 0, st>>>(x, w.mean, w.rstd, gamma, beta, res1, res2, out_f32,
                                               reinterpret_cast<__half*>(out_act), p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "instancenorm launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

extern "C" int hfc_instancenorm_bwd(const float* z, int32_t ld_z, const float* g, int32_t ld_g, const float* gamma,
                                    const float* beta, int32_t c, int32_t n, int32_t hw, float eps, int32_t act, float* dz,
                                    int32_t ld_dz, float* dgamma, float* dbeta, float* dbias, void* dz_act, int32_t act_cpad,
                                    int32_t act_bf16, void* ws, int64_t ws_bytes, void* stream) {
  if (!z || !g || !gamma || !beta || (!dz && !dz_act) || !dgamma || !dbeta || n <= 0 || hw <= 0)
    return set_error(HFC_ERR_INVALID, "instancenorm_bwd: null pointer or empty input");
  if (c % 4 != 0 || c < 4 || ld_z % 4 != 0 || ld_g % 4 != 0 || ld_z < c || ld_g < c || (dz && (ld_dz % 4 != 0 || ld_dz < c)))
    return set_error(HFC_ERR_INVALID, "instancenorm_bwd: needs c %% 4 == 0 and pitches %% 4 == 0, >= c");
  if (dz_act && (act_cpad % 4 != 0 || act_cpad < c))
    return set_error(HFC_ERR_INVALID, "instancenorm_bwd: act_cpad (%d) must be a multiple of 4 >= c", act_cpad);
  if (act != HFC_ACT_NONE && act != HFC_ACT_RELU) return set_error(HFC_ERR_INVALID, "instancenorm_bwd: act must be none or relu");
  if (n > 65535) return set_error(HFC_ERR_UNSUPPORTED, "instancenorm_bwd: more than 65535 images");
  if (!ws || ws_bytes < hfc_instancenorm_ws_bytes(n, c))
    return set_error(HFC_ERR_INVALID, "instancenorm_bwd: workspace too small (hfc_instancenorm_ws_bytes)");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  InParams p{};
  p.n = n; p.hw = hw; p.h = 1; p.w = hw; p.c = c; p.cpad = dz_act ? act_cpad : c; p.ld = ld_z;
  p.act = act; p.eps = eps; p.inv_hw = 1.f / static_cast<float>(hw);
  const int chunks_c = (c + 127) / 128;
  const int chunks_w = (std::max(c, dz_act ? act_cpad : c) + 127) / 128;
  p.strip = strip_for(n, hw, chunks_c, sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const InWs w = carve(ws, n, c);
  rc = launch_stats(z, p, w, chunks_c, st);
  if (rc != HFC_OK) return rc;
  const dim3 grid_c((hw + p.strip - 1) / p.strip, chunks_c, n);
  instnorm_bwd_sums_kernel<<<grid_c, 256, 0, st>>>(z, g, ld_g, w.mean, w.rstd, gamma, beta, w.acc_b, p);
  note_launch();
  instnorm_bwd_finalize_kernel<<<(c + 255) / 256, 256, 0, st>>>(w.acc_b, w.m1, w.m2, dgamma, dbeta, p);
  note_launch();
  const dim3 grid_w((hw + p.strip - 1) / p.strip, chunks_w, n);
  instnorm_bwd_apply_kernel<<<grid_w, 256, 0, st>>>(z, g, ld_g, w.mean, w.rstd, w.m1, w.m2, gamma, beta, dz, ld_dz, dbias,
                                                     reinterpret_cast<uint16_t*>(dz_act), act_cpad, act_bf16, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "instancenorm_bwd launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
