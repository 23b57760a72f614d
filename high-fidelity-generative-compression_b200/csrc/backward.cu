// Data-movement and reduction kernels of the backward pass (all HBM-bound).
//
// The weight gradient of every conv / transposed conv is computed as ONE tcgen05 GEMM (hfc_gemm_nt):
//     dW[c1][(tap, c2)] = sum_p  A1T[c1][p] * COLT[(tap, c2)][p]
// where A1T is one operand transposed to [channels][pixels] and COLT is the transposed im2col matrix of the
// other operand; both are K-major (pixels contiguous) 16-bit matrices built here by a tiled transpose.
// Gradients travel between layers as fp32 rows [pixels][channels]; GEMM operands made from them are bf16.
#include "hfc_internal.h"
#include "hfc_device_utils.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace hfc {

__device__ __forceinline__ uint16_t f32_to_bf16_bits(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// fp16 gradient operands saturate (a stale loss scale clips outliers instead of turning the step into inf / nan)
__device__ __forceinline__ uint16_t f32_to_f16_bits(float v) {
  __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  return *reinterpret_cast<uint16_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// fp32 rows [P][ld] -> border-less 16-bit act buffer [P][cpad] (bf16 or fp16), padding channels = 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_to_act_kernel(const float* __restrict__ rows, int ld, int c, int cpad, long long npix, int to_bf16,
                   uint16_t* __restrict__ out) {
  const long long groups = npix * (cpad / 8);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / (cpad / 8);
    const int g = static_cast<int>(i % (cpad / 8));
    uint16_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      const float f = ch < c ? rows[pix * ld + ch] : 0.f;
      if (to_bf16) v[j] = f32_to_bf16_bits(f);
      else v[j] = f32_to_f16_bits(f);
    }
    *reinterpret_cast<uint4*>(out + pix * cpad + g * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// same, into the interior of a bordered buffer (the border is left alone: the caller zeroes it once)
struct RowsGeomParams { int32_t n, h, w, c, cpad, pt, pl, pb, pr, ld, to_bf16; };
__global__ void __launch_bounds__(256)
rows_to_act_geom_kernel(const float* __restrict__ rows, uint16_t* __restrict__ out, const __grid_constant__ RowsGeomParams p) {
  const long long groups = static_cast<long long>(p.n) * p.h * p.w * (p.cpad / 8);
  const int Hp = p.h + p.pt + p.pb, Wp = p.w + p.pl + p.pr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / (p.cpad / 8);
    const int g = static_cast<int>(i % (p.cpad / 8));
    const int ww = static_cast<int>(pix % p.w);
    const int hh = static_cast<int>((pix / p.w) % p.h);
    const int nn = static_cast<int>(pix / (static_cast<long long>(p.w) * p.h));
    uint16_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      const float f = ch < p.c ? rows[pix * p.ld + ch] : 0.f;
      if (p.to_bf16) v[j] = f32_to_bf16_bits(f);
      else v[j] = f32_to_f16_bits(f);
    }
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(nn) * Hp + hh + p.pt) * Wp + ww + p.pl) * p.cpad + g * 8) =
        *reinterpret_cast<const uint4*>(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Transposed im2col:  out[(tap * c_rows + ch)][p] = src[n, gh*S + dh[tap] + oh0, gw*S + dw[tap] + ow0, ch]
// for p = (n, gh, gw) over the pixel grid, 0 outside the physical source buffer / beyond P.
// ntaps = 1, dh = dw = 0, S = 1 is the plain transpose.  src is a 16-bit act buffer or fp32 rows (-> bf16).
// One block = 64 pixels x 64 channels of one tap, transposed through shared memory.
// ------------------------------------------------------------------------------------------------
struct ColTParams {
  int32_t n, hp, wp, cpad;          // physical source buffer: n x hp x wp pixels, cpad channels (ld for fp32 rows)
  int32_t gh, gw;                   // pixel grid
  int32_t stride, oh0, ow0;         // source coordinate = g*stride + d + o0
  int32_t ntaps, c_rows, c_src;     // channels emitted per tap (multiple of 64), real source channels
  int32_t src_f32;                  // 0: 16-bit copied verbatim, 1: fp32 -> bf16, 2: fp16 -> bf16, 3: fp32 -> fp16
  long long p_total, p_pad;
  int8_t dh[64], dw[64];
};

__global__ void __launch_bounds__(256)
im2col_t_kernel(const void* __restrict__ src, uint16_t* __restrict__ out, const __grid_constant__ ColTParams p) {
  __shared__ uint16_t tile[64][66];
  const int tap = blockIdx.z;
  const int c0 = blockIdx.y * 64;
  const long long p0 = static_cast<long long>(blockIdx.x) * 64;
  // load: thread -> (pixel = tid / 4, 16-channel group = tid % 4)
  {
    const int px = threadIdx.x >> 2, cg = (threadIdx.x & 3) * 16;
    const long long pp = p0 + px;
    bool ok = pp < p.p_total;
    int n = 0, sh = 0, sw = 0;
    if (ok) {
      const int gw_i = static_cast<int>(pp % p.gw);
      const int gh_i = static_cast<int>((pp / p.gw) % p.gh);
      n = static_cast<int>(pp / (static_cast<long long>(p.gw) * p.gh));
      sh = gh_i * p.stride + p.dh[tap] + p.oh0;
      sw = gw_i * p.stride + p.dw[tap] + p.ow0;
      ok = sh >= 0 && sh < p.hp && sw >= 0 && sw < p.wp;
    }
    const size_t base = ((static_cast<size_t>(n) * p.hp + sh) * p.wp + sw) * p.cpad;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ch = c0 + cg + j;
      uint16_t v = 0;
      if (ok && ch < p.c_src) {
        if (p.src_f32 == 1) v = f32_to_bf16_bits(reinterpret_cast<const float*>(src)[base + ch]);
        else if (p.src_f32 == 2) v = f32_to_bf16_bits(__half2float(reinterpret_cast<const __half*>(src)[base + ch]));
        else if (p.src_f32 == 3) v = f32_to_f16_bits(reinterpret_cast<const float*>(src)[base + ch]);
        else v = reinterpret_cast<const uint16_t*>(src)[base + ch];
      }
      tile[cg + j][px] = v;
    }
  }
  __syncthreads();
  // store: thread -> (channel = tid / 4, 16-pixel group = tid % 4): 32 B contiguous per thread
  {
    const int ch = threadIdx.x >> 2, pg = (threadIdx.x & 3) * 16;
    uint16_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = tile[ch][pg + j];
    uint16_t* dst = out + (static_cast<size_t>(tap) * p.c_rows + c0 + ch) * p.p_pad + p0 + pg;
    reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(v)[0];
    reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(v)[1];
  }
}

// 8-channel variant (the RGB image padded to 8, or a <= 8-channel gradient): one thread per pixel, no staging.
__global__ void __launch_bounds__(256)
im2col_t_c8_kernel(const void* __restrict__ src, uint16_t* __restrict__ out, const __grid_constant__ ColTParams p) {
  const int tap = blockIdx.z;
  const long long pp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pp >= p.p_pad) return;
  uint16_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (pp < p.p_total) {
    const int gw_i = static_cast<int>(pp % p.gw);
    const int gh_i = static_cast<int>((pp / p.gw) % p.gh);
    const int n = static_cast<int>(pp / (static_cast<long long>(p.gw) * p.gh));
    const int sh = gh_i * p.stride + p.dh[tap] + p.oh0;
    const int sw = gw_i * p.stride + p.dw[tap] + p.ow0;
    if (sh >= 0 && sh < p.hp && sw >= 0 && sw < p.wp) {
      const size_t base = ((static_cast<size_t>(n) * p.hp + sh) * p.wp + sw) * p.cpad;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < p.c_src)
          v[j] = p.src_f32 == 1 ? f32_to_bf16_bits(reinterpret_cast<const float*>(src)[base + j])
                 : p.src_f32 == 2 ? f32_to_bf16_bits(__half2float(reinterpret_cast<const __half*>(src)[base + j]))
                 : p.src_f32 == 3 ? f32_to_f16_bits(reinterpret_cast<const float*>(src)[base + j])
                                  : reinterpret_cast<const uint16_t*>(src)[base + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) out[(static_cast<size_t>(tap) * 8 + j) * p.p_pad + pp] = v[j];
}

// ------------------------------------------------------------------------------------------------
// GEMM result C[m][(tap, c2)] -> weight gradient in torch layout (scaled, optionally accumulated)
//   conv2d           : dW[co = m][ci = c2][ky][kx]
//   conv_transpose2d : dW[ci = m][co = c2][ky][kx]          (both: dW[m][c2][ky][kx])
// ------------------------------------------------------------------------------------------------
struct PermuteParams {
  int32_t m, c2, c2_rows, ldc, kh, kw, ntaps, accumulate;
  float scale;
  int8_t ky[64], kx[64];
};

// One block per GEMM row m: the row C[m][ntaps * c2_rows] is read with consecutive threads on consecutive addresses,
// transposed through shared memory and written as the contiguous torch slab dW[m][:][:][:] (c2 * kh * kw floats).
// (A thread-per-output-element gather reads with a stride of c2_rows floats: one 32 B sector per element.)
// VEC (odd kh*kw, so T == kh*kw and the shared row IS the torch slab in its final order; c2, c2_rows, ldc multiples of 4,
// 16-byte aligned pointers): float4 reads of the GEMM row and float4 writes of the slab -- the scalar version ran the
// 960 x 960 x 3 x 3 gradients at ~2 TB/s (34 dependent 4-byte accesses per thread and phase, latency-bound).
template <bool VEC>
__global__ void __launch_bounds__(256)
permute_wgrad_kernel(const float* __restrict__ c, float* __restrict__ dw, const __grid_constant__ PermuteParams p) {
  extern __shared__ __align__(16) float s_row[];   // [c2][T], T = kh*kw | 1
  const int m = blockIdx.x;
  const int khkw = p.kh * p.kw;
  const int T = khkw | 1;
  const float* crow = c + static_cast<size_t>(m) * p.ldc;
  for (int tap = 0; tap < p.ntaps; ++tap) {
    if (p.ky[tap] < 0) continue;                   // unused column slot (window packing: 8 slots per filter row)
    const int t = p.ky[tap] * p.kw + p.kx[tap];
    if constexpr (VEC) {
      const float4* src4 = reinterpret_cast<const float4*>(crow + static_cast<size_t>(tap) * p.c2_rows);
      for (int q = threadIdx.x; q < (p.c2 >> 2); q += blockDim.x) {
        const float4 v = __ldg(src4 + q);
        float* d = s_row + (4 * q) * T + t;
        d[0] = v.x; d[T] = v.y; d[2 * T] = v.z; d[3 * T] = v.w;
      }
    } else {
      for (int c2 = threadIdx.x; c2 < p.c2; c2 += blockDim.x) s_row[c2 * T + t] = crow[tap * p.c2_rows + c2];
    }
  }
  __syncthreads();
  float* drow = dw + static_cast<size_t>(m) * p.c2 * khkw;
  const int total = p.c2 * khkw;
  if constexpr (VEC) {                             // T == khkw: s_row[e] is element e of the slab
    float4* d4 = reinterpret_cast<float4*>(drow);
    const float4* s4 = reinterpret_cast<const float4*>(s_row);
    for (int q = threadIdx.x; q < (total >> 2); q += blockDim.x) {
      float4 v = s4[q];
      v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;
      if (p.accumulate) { const float4 o = d4[q]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      d4[q] = v;
    }
  } else {
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int c2 = e / khkw, t = e - c2 * khkw;
      const float v = s_row[c2 * T + t] * p.scale;
      drow[e] = p.accumulate ? drow[e] + v : v;
    }
  }
}

// column sums of fp32 rows [P][ld] -> out[c] (+=), e.g. the bias gradient
__global__ void __launch_bounds__(256)
col_sums_kernel(const float* __restrict__ rows, int ld, int c, long long npix, long long rows_per_block, float scale,
                float* __restrict__ out) {
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < npix ? r0 + rows_per_block : npix;
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += rows[r * ld + col];
    atomicAdd(out + col, acc * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// Adjoint of the materialised padding: gradient over the padded domain (fp32 rows, pitch hq x wq) ->
// gradient of the un-padded tensor.  reflect: every border position adds into the pixel it mirrors;
// zero padding: plain crop.
// ------------------------------------------------------------------------------------------------
struct FoldParams {
  int32_t n, h, w, c, ld_in, ld_out, hq, wq, pt, pl, pb, pr, reflect;
};

__global__ void __launch_bounds__(256)
pad_fold_kernel(const float* __restrict__ dxp, float* __restrict__ dx, const __grid_constant__ FoldParams p) {
  // 32-bit index arithmetic (the launcher checks total < 2^31): this kernel is a pure stream and the 64-bit
  // div / mod chain was most of its instruction count
  const unsigned total = static_cast<unsigned>(p.n) * p.h * p.w * (p.c / 4);
  const unsigned cq = static_cast<unsigned>(p.c / 4);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / cq;
    const int c4 = static_cast<int>(i - pix * cq) * 4;
    const unsigned row = pix / static_cast<unsigned>(p.w);
    const int ww = static_cast<int>(pix - row * static_cast<unsigned>(p.w));
    const int nn = static_cast<int>(row / static_cast<unsigned>(p.h));
    const int hh = static_cast<int>(row - static_cast<unsigned>(nn) * static_cast<unsigned>(p.h));
    int rows[3], cols[3];
    const int nr = mirror_targets(hh, p.h, p.pt, p.pb, p.reflect != 0, rows);
    const int nc = mirror_targets(ww, p.w, p.pl, p.pr, p.reflect != 0, cols);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ri = 0; ri < nr; ++ri)
      for (int ci = 0; ci < nc; ++ci) {
        const float4 v = *reinterpret_cast<const float4*>(
            dxp + ((static_cast<size_t>(nn) * p.hq + rows[ri]) * p.wq + cols[ci]) * p.ld_in + c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    *reinterpret_cast<float4*>(dx + static_cast<size_t>(pix) * p.ld_out + c4) = acc;
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

extern "C" int hfc_rows_to_act(const float* rows, int32_t ld, int64_t npix, int32_t c, int32_t cpad, int32_t to_bf16,
                               void* out, void* stream) {
  if (!rows || !out || npix <= 0 || c <= 0 || cpad % 8 != 0 || cpad < c || ld < c)
    return set_error(HFC_ERR_INVALID, "rows_to_act: bad arguments");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  rows_to_act_kernel<<<grid_for(npix * (cpad / 8), sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rows, ld, c, cpad, npix, to_bf16, reinterpret_cast<uint16_t*>(out));
  HFC_CHECK_LAUNCH("rows_to_act launch");
  return HFC_OK;
}

extern "C" int hfc_rows_to_act_geom(const float* rows, int32_t ld, const hfc_act_geom* g, int32_t to_bf16, void* out,
                                    void* stream) {
  if (!rows || !g || !out || g->n <= 0 || g->h <= 0 || g->w <= 0 || g->c <= 0 || g->cpad % 8 != 0 || g->cpad < g->c || ld < g->c)
    return set_error(HFC_ERR_INVALID, "rows_to_act_geom: bad arguments");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  RowsGeomParams p;
  p.n = g->n; p.h = g->h; p.w = g->w; p.c = g->c; p.cpad = g->cpad; p.pt = g->pt; p.pl = g->pl; p.pb = g->pb; p.pr = g->pr;
  p.ld = ld; p.to_bf16 = to_bf16;
  rows_to_act_geom_kernel<<<grid_for(static_cast<long long>(g->n) * g->h * g->w * (g->cpad / 8), sms), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(rows, reinterpret_cast<uint16_t*>(out), p);
  HFC_CHECK_LAUNCH("rows_to_act_geom launch");
  return HFC_OK;
}

extern "C" int hfc_im2col_t(const void* src, int32_t src_f32, int32_t n, int32_t hp, int32_t wp, int32_t cpad,
                            int32_t c_src, int32_t gh, int32_t gw, int32_t stride, int32_t oh0, int32_t ow0,
                            int32_t ntaps, const int8_t* dh_host, const int8_t* dw_host, int32_t c_rows,
                            int64_t p_pad, void* out, void* stream) {
  if (!src || !out || ntaps <= 0 || ntaps > 64 || (c_rows % 64 != 0 && c_rows != 8) || p_pad % 64 != 0 || !dh_host ||
      !dw_host || (c_rows == 8 && c_src > 8))
    return set_error(HFC_ERR_INVALID, "im2col_t: bad arguments");
  const long long p_total = static_cast<long long>(n) * gh * gw;
  if (p_pad < p_total || c_src > cpad) return set_error(HFC_ERR_INVALID, "im2col_t: p_pad < pixels or c_src > cpad");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  ColTParams p;
  p.n = n; p.hp = hp; p.wp = wp; p.cpad = cpad; p.gh = gh; p.gw = gw; p.stride = stride; p.oh0 = oh0; p.ow0 = ow0;
  p.ntaps = ntaps; p.c_rows = c_rows; p.c_src = c_src; p.src_f32 = src_f32; p.p_total = p_total; p.p_pad = p_pad;
  memset(p.dh, 0, sizeof(p.dh)); memset(p.dw, 0, sizeof(p.dw));
  memcpy(p.dh, dh_host, ntaps); memcpy(p.dw, dw_host, ntaps);
  if (c_rows == 8) {
    dim3 grid8(static_cast<unsigned>((p_pad + 255) / 256), 1, ntaps);
    im2col_t_c8_kernel<<<grid8, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<uint16_t*>(out), p);
  } else {
    dim3 grid(static_cast<unsigned>(p_pad / 64), c_rows / 64, ntaps);
    im2col_t_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<uint16_t*>(out), p);
  }
  HFC_CHECK_LAUNCH("im2col_t launch");
  return HFC_OK;
}

extern "C" int hfc_permute_wgrad(const float* c, int32_t ldc, int32_t m, int32_t c2, int32_t c2_rows, int32_t kh,
                                 int32_t kw, int32_t ntaps, const int8_t* ky_host, const int8_t* kx_host, float scale,
                                 int32_t accumulate, float* dw, void* stream) {
  if (!c || !dw || ntaps <= 0 || ntaps > 64 || !ky_host || !kx_host) return set_error(HFC_ERR_INVALID, "permute_wgrad: bad arguments");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  PermuteParams p;
  p.m = m; p.c2 = c2; p.c2_rows = c2_rows; p.ldc = ldc; p.kh = kh; p.kw = kw; p.ntaps = ntaps; p.accumulate = accumulate;
  p.scale = scale;
  memset(p.ky, 0, sizeof(p.ky)); memset(p.kx, 0, sizeof(p.kx));
  memcpy(p.ky, ky_host, ntaps); memcpy(p.kx, kx_host, ntaps);
  const size_t smem = static_cast<size_t>(c2) * ((kh * kw) | 1) * sizeof(float);
  if (smem > 96 * 1024) return set_error(HFC_ERR_UNSUPPORTED, "permute_wgrad: c2 * kh * kw too large for the row buffer");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(permute_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(permute_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  static const bool vec_on = [] { const char* e = getenv("HFC_PERMUTE_VEC"); return !(e && e[0] == '0'); }();
  const bool vec = vec_on && ((kh * kw) & 1) && c2 % 4 == 0 && c2_rows % 4 == 0 && ldc % 4 == 0 &&
                   reinterpret_cast<uintptr_t>(c) % 16 == 0 && reinterpret_cast<uintptr_t>(dw) % 16 == 0;
  if (vec) permute_wgrad_kernel<true><<<m, 256, smem, static_cast<cudaStream_t>(stream)>>>(c, dw, p);
  else permute_wgrad_kernel<false><<<m, 256, smem, static_cast<cudaStream_t>(stream)>>>(c, dw, p);
  HFC_CHECK_LAUNCH("permute_wgrad launch");
  return HFC_OK;
}

extern "C" int hfc_col_sums(const float* rows, int32_t ld, int64_t npix, int32_t c, float scale, float* out,
                            void* stream) {
  if (!rows || !out || npix <= 0 || c <= 0) return set_error(HFC_ERR_INVALID, "col_sums: bad arguments");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  long long rpb = std::max<long long>(64, (npix + sms * 4LL - 1) / (sms * 4LL));
  const long long blocks = (npix + rpb - 1) / rpb;
  col_sums_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, ld, c, npix, rpb, scale, out);
  HFC_CHECK_LAUNCH("col_sums launch");
  return HFC_OK;
}

extern "C" int hfc_pad_fold(const float* dxp, int32_t ld_in, int32_t hq, int32_t wq, const hfc_act_geom* g,
                            int32_t reflect, float* dx, int32_t ld_out, void* stream) {
  if (!dxp || !dx || !g) return set_error(HFC_ERR_INVALID, "pad_fold: null pointer");
  if (g->c % 4 != 0 || ld_in % 4 != 0 || ld_out % 4 != 0 || hq < g->h + g->pt + g->pb || wq < g->w + g->pl + g->pr)
    return set_error(HFC_ERR_INVALID, "pad_fold: channel counts must be multiples of 4 and the padded domain must fit");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  if (static_cast<long long>(g->n) * g->h * g->w * (g->c / 4) >= (1LL << 31))
    return set_error(HFC_ERR_UNSUPPORTED, "pad_fold: more than 2^31 float4 elements");
  FoldParams p;
  p.n = g->n; p.h = g->h; p.w = g->w; p.c = g->c; p.ld_in = ld_in; p.ld_out = ld_out; p.hq = hq; p.wq = wq;
  p.pt = g->pt; p.pl = g->pl; p.pb = g->pb; p.pr = g->pr; p.reflect = reflect;
  pad_fold_kernel<<<grid_for(static_cast<long long>(g->n) * g->h * g->w * (g->c / 4), sms), 256, 0,
                    static_cast<cudaStream_t>(stream)>>>(dxp, dx, p);
  HFC_CHECK_LAUNCH("pad_fold launch");
  return HFC_OK;
}
