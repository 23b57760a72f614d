// Implicit-GEMM weight gradient of a convolution / transposed convolution for sm_100a.
//
//   C[m][tap][c2] = sum over pixels p   P[p][m] * S[p * stride + tap][c2]
//
// P ("plain") and S ("shifted") are NHWC 16-bit activation buffers: for conv2d P = dL/dy and S = the saved layer
// input x (sampled at o*stride + tap - pad), for conv_transpose2d P = x and S = dL/dy.  The GEMM's K dimension is the
// PIXEL index, so both operands are consumed MN-major straight from their NHWC layout: a TMA box {64 channels,
// bw x bh x bn pixels} lands in shared memory as 64 K-rows of 128 B (one 64-channel chunk), which is exactly the
// canonical 128B-swizzled MN-major UMMA operand.  No transposed copy and no im2col matrix is ever materialised
// (the explicit path this replaces wrote and re-read ntaps x the activation tensor).
//
// * work item = (M tile of 128 P-channels, filter tap, N tile of up to 256 S-channels, K split); K loop over
//   64-pixel blocks, 4 x tcgen05.mma (M=128, N=64..256, K=16) per block, fp32 accumulation in TMEM
//   (two accumulator stages: the epilogue of item i overlaps the main loop of item i+1)
// * persistent CTAs, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue
//   (thread == P-channel row; fp32 stores, or atomicAdd when the pixels are split across CTAs)
// * zero padding and tile overhang are TMA out-of-bounds fill (zeros contribute nothing); reflection padding was
//   materialised in S by its producer.
//
// Reference: the autograd of F.conv2d / F.conv_transpose2d w.r.t. the weight in every layer of src/network/*.py as
// run by train.py:49-59.
#include "hfc_internal.h"
#include "hfc_ptx.cuh"

#include <cstdlib>

#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace hfc {

static constexpr int kWgChunkBytes = 64 * 128;   // 64 pixels x 64 channels x 2 B
static constexpr int kWgThreads = 64 + 128;
static constexpr int kWgMaxStages = 8;
static constexpr int kWgAccStride = 256;
static constexpr int kWgTmemCols = 512;
static constexpr int kWgMaxTaps = 64;

struct WgradParams {
  int32_t bw, bh, bn;                  // pixel box, bw * bh * bn == 64
  int32_t tiles_w, tiles_h, tiles_n;   // pixel tiles of the P grid
  int32_t num_kb;                      // tiles_w * tiles_h * tiles_n
  int32_t m_tiles, n_tiles, nch;       // nch = 64-channel chunks per N tile (1..4)
  int32_t p_chunks;                    // P chunks fetched per stage (1 when P has <= 64 channels: the other half stays zero)
  int32_t ntaps, k_splits, kb_per_split;
  int32_t tapg, ngroups;               // taps per work item (1 = one tap per item) and ceil(ntaps / tapg)
  int32_t stages;
  int32_t stride;                      // S sampling stride
  int32_t p_h0, p_w0, s_h0, s_w0;      // interior origin of P / S inside their (bordered) buffers
  int32_t c1, c2_rows, ldc;
  uint32_t fmt;
  int32_t atomic;
  float* out;
  int8_t tap_dh[kWgMaxTaps];
  int8_t tap_dw[kWgMaxTaps];
};

// kPair: CTA pairs (tcgen05 cta_group::2, launched as clusters of 2).  One UMMA then spans M = 256 P-channels (128 in
// each CTA's TMEM) and the S tile is split in halves between the two CTAs' shared memories, so each SM fetches
// 16 KB (P) + 16 KB (S) per 64-pixel block instead of 16 + 32 KB: the single-CTA kernel is bound by the L2 -> SM
// operand traffic (~10 TB/s chip-wide on the 960x960 layers), not by the tensor pipe.
// kNsub (2 only with kPair): N tiles per work item -- one in each 256-column TMEM half, both accumulating against ONE
// fetch of the P tile (16 + 2 x 16 KB per pixel block for two tiles instead of 2 x (16 + 16)); no accumulator double
// buffering then.  Same trade as conv_igemm_kernel<true, 2>.
template <bool kPair, int kNsub>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_s,
                   const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int s_chunks = kPair ? p.nch / 2 : p.nch;            // S chunks held by this CTA
  const int stage_bytes = (2 + kNsub * s_chunks) * kWgChunkBytes;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;     // 0 = pair leader
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgMaxStages;
  uint64_t* tfull_bar = bars + 2 * kWgMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_p);
    tma_prefetch_desc(&tmap_s);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kPair ? 8 : 4);      // pair: the leader waits for the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (kPair) {
      tmem_alloc_pair(tmem_slot, kWgTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kWgTmemCols);
      tmem_relinquish();
    }
  }
  if (p.p_chunks == 1) {
    // P has <= 64 channels: rows 64..127 of the M = 128 tile are never fetched -- zero them once (generic-proxy
    // stores, made visible to the tensor core's async proxy by the fence) instead of TMA zero-filling them per stage
    for (int s = 0; s < p.stages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * stage_bytes + kWgChunkBytes);
      for (int i = threadIdx.x; i < kWgChunkBytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();     // the peer's barriers exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item order: K split fastest, then N tile, tap, M tile -> concurrently running CTAs share P tiles in L2
  // (pair mode: m_tiles counts 256-channel tiles and one "CTA" of the loops below is a pair)
  const int n_groups = p.n_tiles / kNsub;      // the host makes n_tiles a multiple of kNsub
  const int total_items = p.m_tiles * p.ngroups * n_groups * p.k_splits;
  const int cta0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int ncta = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // The whole warp walks the loop with warp-uniform operands and ONE elected lane issues (the first version ran it on lane
    // 0 alone: three integer divisions per K block for the pixel-block coordinates and a divergent context in which every
    // TMA operand went through an ELECT / R2UR.BROADCAST sequence -- ~1 000 cycles per K block against ~500-1 000 cycles
    // of MMA work; same finding as in conv_igemm.cu, profiles/r02_ncu_bigmap_thin.json).  Coordinates advance
    // incrementally.
    {
      int s = 0;
      uint32_t ph = 0;
      for (int it = cta0; it < total_items; it += ncta) {
        int r = it;
        const int ks = r % p.k_splits; r /= p.k_splits;
        const int nt = (r % n_groups) * kNsub; r /= n_groups;
        const int tap = (r % p.ngroups) * p.tapg;      // first tap of the item's group
        const int mt = r / p.ngroups;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        int twi = kb0 % p.tiles_w, thi = (kb0 / p.tiles_w) % p.tiles_h, tni = kb0 / (p.tiles_w * p.tiles_h);
        const int dw = p.tap_dw[tap], dh = p.tap_dh[tap];
        for (int kb = kb0; kb < kb1; ++kb) {
          const int gw = twi * p.bw, gh = thi * p.bh, gn = tni * p.bn;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + 2 * kWgChunkBytes;
          const int sw0 = gw * p.stride + p.s_w0 + dw;
          const int sh0 = gh * p.stride + p.s_h0 + dh;
          if (elect_one()) {
            if constexpr (kPair) {
              // both CTAs fill their own stage; every byte is accounted on the LEADER's barrier
              if (crank == 0) mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(2 * stage_bytes));
#pragma unroll
              for (int j = 0; j < 2; ++j)
                tma_load_4d_pair(sa + j * kWgChunkBytes, &tmap_p, &full_bar[s], (mt * 4 + crank * 2 + j) * 64, gw + p.p_w0,
                                 gh + p.p_h0, gn);
#pragma unroll
              for (int js = 0; js < kNsub; ++js)
                for (int j = 0; j < s_chunks; ++j)
                  tma_load_4d_pair(sb + (js * s_chunks + j) * kWgChunkBytes, &tmap_s, &full_bar[s],
                                   ((nt + js) * p.nch + crank * s_chunks + j) * 64, sw0, sh0, gn);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>((p.p_chunks + p.nch) * kWgChunkBytes));
              for (int j = 0; j < p.p_chunks; ++j)
                tma_load_4d(sa + j * kWgChunkBytes, &tmap_p, &full_bar[s], (mt * 2 + j) * 64, gw + p.p_w0, gh + p.p_h0, gn);
              if (p.tapg > 1) {
                // tap group: the N tile is `nch` TAPS of the one 64-channel chunk of S, all multiplied against ONE fetch of
                // the P tile (a tap past the last one re-loads the last: its columns are never stored)
                for (int j = 0; j < p.nch; ++j) {
                  const int tj = min(tap + j, p.ntaps - 1);
                  tma_load_4d(sb + j * kWgChunkBytes, &tmap_s, &full_bar[s], 0, gw * p.stride + p.s_w0 + p.tap_dw[tj],
                              gh * p.stride + p.s_h0 + p.tap_dh[tj], gn);
                }
              } else {
                for (int j = 0; j < p.nch; ++j)
                  tma_load_4d(sb + j * kWgChunkBytes, &tmap_s, &full_bar[s], (nt * p.nch + j) * 64, sw0, sh0, gn);
              }
            }
          }
          __syncwarp();
          if (++twi == p.tiles_w) { twi = 0; if (++thi == p.tiles_h) { thi = 0; ++tni; } }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && (!kPair || crank == 0)) {
    const uint32_t idesc = make_idesc_f16_mn(p.fmt, kPair ? 256u : 128u, static_cast<uint32_t>(p.nch * 64));
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int it = cta0; it < total_items; it += ncta) {
      const int ks = it % p.k_splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kWgAccStride;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);      // warp-uniform operands, one elected issuer
        const uint32_t b_addr = a_addr + 2 * kWgChunkBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // K = 16 pixels = 16 rows of 128 B = two 1024 B swizzle atoms per step
            const uint64_t a_desc = make_sw128_mnmajor_desc(a_addr + k * 2048, kWgChunkBytes, 1024);
            const uint64_t b_desc = make_sw128_mnmajor_desc(b_addr + k * 2048, kWgChunkBytes, 1024);
            if constexpr (kPair) {
              umma_f16_pair(d_tmem, a_desc, b_desc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
              if constexpr (kNsub == 2) {
                const uint64_t b_desc2 = make_sw128_mnmajor_desc(b_addr + s_chunks * kWgChunkBytes + k * 2048, kWgChunkBytes, 1024);
                umma_f16_pair(d_tmem + kWgAccStride, a_desc, b_desc2, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
              }
            } else {
              umma_f16(d_tmem, a_desc, b_desc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            }
          }
          if constexpr (kPair) {
            umma_commit_pair_mc(&empty_bar[s], 3);                 // frees the stage in both CTAs
            if (kb == kb1 - 1) umma_commit_pair_mc(&tfull_bar[as], 3);
          } else {
            umma_commit(&empty_bar[s]);
            if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (++as == (kNsub == 2 ? 1 : 2)) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 2) {
    // ===================== epilogue (warps 2..5): thread == row m of the tile =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ncols = p.nch * 64;
    int as = 0;
    uint32_t aph = 0;
    for (int it = cta0; it < total_items; it += ncta) {
      int r = it / p.k_splits;
      const int nt0 = (r % n_groups) * kNsub; r /= n_groups;
      const int tap = (r % p.ngroups) * p.tapg;
      const int mt = r / p.ngroups;
      const int row = (kPair ? mt * 2 + static_cast<int>(crank) : mt) * 128 + m;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
#pragma unroll
      for (int jsub = 0; jsub < kNsub; ++jsub) {
      const int nt = nt0 + jsub;
      const uint32_t t_row = tmem_base + (kNsub == 2 ? jsub : as) * kWgAccStride + (static_cast<uint32_t>(q * 32) << 16);
      const int col0 = nt * ncols;
      float* dst = p.out + static_cast<size_t>(row) * p.ldc + static_cast<size_t>(tap) * p.c2_rows + col0;
      // columns that exist: the rest of the tap's c2_rows, or (tap group: c2_rows == 64, consecutive taps are consecutive
      // 64-column blocks of the row) the taps that are left
      const int col_limit = p.tapg > 1 ? (p.ntaps - tap) * 64 : p.c2_rows - col0;
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + c0, v);
        tmem_ld_wait();
        if (row < p.c1 && c0 < col_limit) {
          if (p.atomic) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + c0 + j, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              reinterpret_cast<float4*>(dst + c0)[j4] =
                  make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]), __uint_as_float(v[4 * j4 + 2]),
                              __uint_as_float(v[4 * j4 + 3]));
          }
        }
      }
      }  // jsub
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_leader(&tempty_bar[as]);
        else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == (kNsub == 2 ? 1 : 2)) { as = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();     // nobody exits while the peer may still signal its barriers / read its smem
  if (warp == 1) {
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kWgTmemCols);
    else tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiledW get_encode_fn_w() {
  static PFN_encodeTiledW fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledW>(ptr);
  }
  return fn;
}

static int next_pow2_w(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

static int encode_act_map(PFN_encodeTiledW encode, CUtensorMap* tm, const void* base, const hfc_act_geom& g, int bw, int bh,
                          int bn, int stride, bool window = false) {
  const int Hp = g.h + g.pt + g.pb, Wp = g.w + g.pl + g.pr;
  // window packing (8-channel pitch): "channel" dim = 64 elements = 8 consecutive pixels x 8 channels, pixel dim =
  // window start; the overlapping 16 B stride makes every K row (pixel p) the 8-pixel window that starts at p
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(window ? 64 : g.cpad), static_cast<cuuint64_t>(window ? Wp - 7 : Wp),
                        static_cast<cuuint64_t>(Hp), static_cast<cuuint64_t>(g.n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(g.cpad) * 2, static_cast<cuuint64_t>(Wp) * g.cpad * 2,
                           static_cast<cuuint64_t>(Hp) * Wp * g.cpad * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw * stride), static_cast<cuuint32_t>(bh * stride), static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

}  // namespace hfc

using namespace hfc;

extern "C" int hfc_wgrad(const hfc_wgrad_desc* d, const void* plain, const void* shifted, float* c, int32_t ldc,
                         void* stream) {
  if (!d || !plain || !shifted || !c) return set_error(HFC_ERR_INVALID, "wgrad: null pointer");
  const hfc_act_geom& pg = d->plain;
  const hfc_act_geom& sg = d->shifted;
  if (pg.n <= 0 || pg.h <= 0 || pg.w <= 0 || pg.c <= 0 || sg.c <= 0 || pg.n != sg.n)
    return set_error(HFC_ERR_INVALID, "wgrad: bad geometry");
  const bool window = d->window != 0;
  if (pg.cpad % 64 != 0 || pg.c > pg.cpad || sg.c > sg.cpad || (!window && sg.cpad % 64 != 0))
    return set_error(HFC_ERR_INVALID, "wgrad: both operands need channel pitches that are multiples of 64");
  if (window && (sg.cpad != 8 || d->stride != 1 || sg.w + sg.pl + sg.pr < 8))
    return set_error(HFC_ERR_INVALID, "wgrad: window packing needs an 8-channel-pitch shifted operand, stride 1, width >= 8");
  if (d->ntaps <= 0 || d->ntaps > kWgMaxTaps) return set_error(HFC_ERR_INVALID, "wgrad: 1..64 taps");
  if (d->stride != 1 && d->stride != 2) return set_error(HFC_ERR_INVALID, "wgrad: stride must be 1 or 2");
  if (d->bf16 != 0 && d->bf16 != 1) return set_error(HFC_ERR_INVALID, "wgrad: bf16 must be 0 or 1");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  PFN_encodeTiledW encode = get_encode_fn_w();
  if (!encode) return set_error(HFC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not found");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  WgradParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.bw = std::min(64, next_pow2_w(pg.w));
  kp.bh = std::min(64 / kp.bw, next_pow2_w(pg.h));
  kp.bn = 64 / (kp.bw * kp.bh);
  kp.tiles_w = (pg.w + kp.bw - 1) / kp.bw;
  kp.tiles_h = (pg.h + kp.bh - 1) / kp.bh;
  kp.tiles_n = (pg.n + kp.bn - 1) / kp.bn;
  // overhanging pixel tiles rely on out-of-bounds zero fill of P; a bordered P would deliver border pixels instead
  if ((pg.pt | pg.pl | pg.pb | pg.pr) && (pg.w % kp.bw != 0 || pg.h % kp.bh != 0))
    return set_error(HFC_ERR_UNSUPPORTED, "wgrad: a bordered plain operand needs a grid that tiles exactly (%d x %d by %d x %d)",
                     pg.w, pg.h, kp.bw, kp.bh);
  kp.num_kb = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  const int m_chunks = (pg.c + 63) / 64, n_chunks = window ? 1 : (sg.c + 63) / 64;
  // CTA pairs whenever both operands are wide enough for an M = 256 x N >= 128 tile
  static const bool env_no_pair = getenv("HFC_NO_WGRAD_PAIR") != nullptr;
  const bool pair = !window && !env_no_pair && d->pair != 2 && m_chunks >= 3 && n_chunks >= 2 && sms >= 2;
  kp.m_tiles = pair ? (m_chunks + 3) / 4 : (m_chunks + 1) / 2;
  int best = 1;
  long long best_cost = -1;
  for (int nch = 1; nch <= 4; ++nch) {     // cost ~ tiles x (P chunks + S chunks) of operand traffic / MMA time
    const long long cost = static_cast<long long>((n_chunks + nch - 1) / nch) * (2 + nch);
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = nch; }
  }
  if (pair) best = n_chunks >= 3 ? 4 : 2;   // the S tile is split in halves between the two CTAs
  kp.nch = best;
  kp.ntaps = d->ntaps;
  // Narrow shifted operand (<= 64 channels: the big-map layers E1 / E2 / G.up4 / G3): one tap per item re-fetches the P
  // tile for every filter tap -- 9 x (16 + 8) KB per pixel block on G.up4, L2 -> SM bound at ~390 us.  Group up to four
  // taps into one N tile (four S chunks against ONE P fetch): 9 taps cost 3 x 16 + 9 x 8 KB instead.
  static const bool env_tapg = [] { const char* e = getenv("HFC_WGRAD_TAPG"); return !(e && e[0] == '0'); }();
  kp.tapg = 1;
  if (!pair && env_tapg && n_chunks == 1 && d->ntaps > 1) {
    kp.tapg = std::min(4, static_cast<int>(d->ntaps));
    kp.nch = kp.tapg;
  }
  kp.ngroups = (kp.ntaps + kp.tapg - 1) / kp.tapg;
  kp.n_tiles = kp.tapg > 1 ? 1 : (n_chunks + kp.nch - 1) / kp.nch;
  kp.c2_rows = n_chunks * 64;
  if (ldc < d->ntaps * kp.c2_rows || ldc % 4 != 0)
    return set_error(HFC_ERR_INVALID, "wgrad: ldc (%d) must be a multiple of 4 and >= ntaps * round_up(c2, 64) = %d", ldc,
                     d->ntaps * kp.c2_rows);
  kp.p_chunks = m_chunks == 1 ? 1 : 2;
  // pair mode with two N tiles per work item (cf. conv_igemm.cu): correct (tests/test_gpu_grad.py passes with it) but
  // MEASURED SLOWER here -- 960x960 layers 178 us vs 129 us -- although it moves 25 % fewer bytes: with MN-major
  // operands the 8 MMAs per pixel block take ~2600 clk, i.e. the kernel stops being L2-bound and becomes bound by the
  // tensor core's transposed operand fetch.  Opt-in (HFC_WGRAD_NSUB=1) until that is understood.
  static const bool env_nsub = getenv("HFC_WGRAD_NSUB") != nullptr;
  int nsub = 1;
  if (pair && env_nsub && kp.n_tiles % 2 == 0 && kp.nch == 4) {
    const long long it1 = static_cast<long long>(kp.m_tiles) * kp.ntaps * kp.n_tiles, it2 = it1 / 2;
    const int prs = std::max(1, sms / 2);
    if (((it2 + prs - 1) / prs) * 180 < ((it1 + prs - 1) / prs) * 100) nsub = 2;
  }
  const int items = kp.m_tiles * kp.ngroups * (kp.n_tiles / nsub);
  const int workers = pair ? sms / 2 : sms;      // CTAs, or CTA pairs
  int ks = d->k_splits;
  if (ks <= 0) {
    // fill (at most) two full waves of workers -- never a partial third one -- but keep >= 8 pixel blocks per item so
    // that the pipeline fill and the epilogue amortise
    ks = std::max(1, std::min((2 * workers) / items, kp.num_kb / 8));
  }
  ks = std::max(1, std::min(ks, kp.num_kb));
  kp.kb_per_split = (kp.num_kb + ks - 1) / ks;
  kp.k_splits = (kp.num_kb + kp.kb_per_split - 1) / kp.kb_per_split;
  kp.atomic = kp.k_splits > 1 ? 1 : 0;
  const int stage_bytes = (2 + nsub * (pair ? kp.nch / 2 : kp.nch)) * kWgChunkBytes;
  kp.stages = std::max(2, std::min((226 * 1024 - 1024 - 512) / stage_bytes, kWgMaxStages));
  kp.stride = d->stride;
  kp.p_h0 = pg.pt; kp.p_w0 = pg.pl; kp.s_h0 = sg.pt; kp.s_w0 = sg.pl;
  kp.c1 = pg.c; kp.ldc = ldc;
  kp.fmt = static_cast<uint32_t>(d->bf16);
  kp.out = c;
  memcpy(kp.tap_dh, d->tap_dh, d->ntaps);
  memcpy(kp.tap_dw, d->tap_dw, d->ntaps);

  CUtensorMap tmP, tmS;
  int er = encode_act_map(encode, &tmP, plain, pg, kp.bw, kp.bh, kp.bn, 1);
  if (er) return set_error(HFC_ERR_LAUNCH, "cuTensorMapEncodeTiled(P) failed: %d", er);
  er = encode_act_map(encode, &tmS, shifted, sg, kp.bw, kp.bh, kp.bn, d->stride, window);
  if (er) return set_error(HFC_ERR_LAUNCH, "cuTensorMapEncodeTiled(S) failed: %d", er);

  if (kp.atomic) {
    if (cudaMemsetAsync(c, 0, static_cast<size_t>(pg.c) * ldc * sizeof(float), st) != cudaSuccess)
      return set_error(HFC_ERR_LAUNCH, "wgrad: memset failed");
  }
  const size_t smem = static_cast<size_t>(kp.stages) * stage_bytes + 1024 + 512;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_igemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_igemm_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_igemm_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = std::min(items * kp.k_splits, workers) * (pair ? 2 : 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = !pair ? cudaLaunchKernelEx(&cfg, wgrad_igemm_kernel<false, 1>, tmP, tmS, kp)
                  : nsub == 2 ? cudaLaunchKernelEx(&cfg, wgrad_igemm_kernel<true, 2>, tmP, tmS, kp)
                              : cudaLaunchKernelEx(&cfg, wgrad_igemm_kernel<true, 1>, tmP, tmS, kp);
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "wgrad_igemm launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}

// fp16 -> bf16 copy of an activation buffer (flat): the saved forward activations are fp16, the gradients bf16, and
// tcgen05 kind::f16 wants both operands in one format.
namespace hfc {
__global__ void __launch_bounds__(256)
f16_to_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = src[i];
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = *reinterpret_cast<const __half2*>(&in[j]);
      const float2 f = __half22float2(h);
      const __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
      o[j] = *reinterpret_cast<const uint32_t*>(&b);
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace hfc

extern "C" int hfc_act_to_bf16(const void* src_f16, void* dst_bf16, int64_t count, void* stream) {
  if (!src_f16 || !dst_bf16 || count <= 0 || count % 8 != 0)
    return set_error(HFC_ERR_INVALID, "act_to_bf16: null pointer or count not a multiple of 8");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != HFC_OK) return rc;
  const long long n8 = count / 8;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n8 + 255) / 256, sms * 8LL)));
  f16_to_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(src_f16),
                                                                             reinterpret_cast<uint4*>(dst_bf16), n8);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "act_to_bf16 launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
