// Implicit-GEMM convolution / transposed convolution for sm_100a.
//
//   D[M = batch*grid_h*grid_w pixels, N = cout] = sum over taps, cin  A[pixel + tap, cin] * W[cout, tap, cin]
//
// * A tiles (128 pixels x 64 channels, fp16) are fetched by 4-D TMA boxes straight from the NHWC
//   activation buffer: one box per (filter tap, 64-channel chunk).  Zero padding is TMA's
//   out-of-bounds fill; reflection padding was materialised by the producer of the buffer;
//   stride-2 convolutions use the tensor map's element strides; stride-2 transposed convolutions
//   are decomposed into 4 sub-pixel phases (one launch each).
// * W tiles (block_n x 64) come from a pre-packed K-major matrix via a 2-D TMA box.
// * tcgen05.mma (cta_group::1, kind::f16, M=128, N=block_n, K=16) accumulates into TMEM; two
//   accumulator stages of 256 columns let the epilogue of tile i overlap the main loop of i+1.
// * Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer,
//   warps 2..9 = epilogue (two per TMEM lane quadrant; thread == output pixel, so the per-pixel
//   ChannelNorm over channels is an in-thread reduction with no shuffles).
// * Epilogue: + bias, optional fused ChannelNorm2D (unbiased variance, eps), ReLU / LeakyReLU,
//   then one of: NHWC fp16 with reflected border (input of the next conv), NHWC fp32 rows (input
//   of the stand-alone ChannelNorm for 480/960 channels), NCHW fp32 (module boundary).
//
// Reference call sites this replaces are listed in include/hfc.h (hfc_conv_desc).
#include "hfc_internal.h"
#include "hfc_ptx.cuh"

#include <cstdlib>
#include <type_traits>

#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace hfc {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;                       // 64 x 16-bit = one 128 B swizzle row
static constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
static constexpr int kEpiThreads = 256;                  // 8 epilogue warps (2 per TMEM lane quadrant)
static constexpr int kThreads = 64 + kEpiThreads;        // + TMA producer warp + MMA issuer warp
// "thin" instantiation (narrow N, big maps: the layers whose time is the epilogue): 16 epilogue warps = 4 groups of 4, one
// group per accumulator stage, 4 stages of 128 TMEM columns
static constexpr int kThinEpiThreads = 512;
static constexpr int kThinThreads = 64 + kThinEpiThreads;
static constexpr int kThinAccStages = 4;
static constexpr int kThinAccStride = 128;
static constexpr int kMaxTaps = 64;
static constexpr int kAccStride = 256;                   // TMEM columns per accumulator stage
static constexpr int kTmemCols = 512;
static constexpr int kMaxStages = 8;
static constexpr int kParamStride = 256;                 // floats per epilogue parameter row (>= block_n)
static constexpr int kMaxKb = 512;                       // K blocks per tile the producer's coordinate table holds
static constexpr int kTailBytes = 256 + 2 * 3 * kParamStride * 4 + 2 * 512 * 4 + kMaxKb * 4;  // barriers + parameter rows + stats exchange + K-block table
static constexpr int kTapnLd = 33;                       // padded row pitch (floats) of the tap-in-N staging tile
static constexpr int kTapnBytes = 2 * kBlockM * kTapnLd * 4;  // double-buffered [128][33] fp32 (thin: one per group, x2)

struct ConvKernelParams {
  int32_t tw, th, tn;                 // tile extents, tw*th*tn <= 128 (rows past the product are idle)
  int32_t tx_short;                   // bytes of the 16 KB A tile that the (smaller) TMA box does not deliver
  int32_t tiles_w, tiles_h, tiles_n;  // M tiles per dimension
  int32_t n_tiles;                    // N tiles
  int32_t block_n;
  int32_t num_kb, c_chunks, ntaps;
  int32_t stages;
  int32_t cm, cn;                     // cluster extent along M tiles / N tiles (1 or 2 each)
  int32_t pair;                       // 1: CTA pairs (cta_group::2 UMMA, M = 256 across two SMs); needs cm == 2
  int32_t b_rows;                     // rows of the B tile held by one CTA (block_n, or block_n/2 in pair mode)
  int32_t nsub;                       // pair mode: N tiles per work item (2: both 256-column TMEM halves accumulate
                                      // against ONE fetch of the A tile; no accumulator double buffering then)
  int32_t a_split_n;                  // A slice split: 1 = along the batch dim of the box, 0 = along rows
  int32_t tapn, w_step;               // 'tap-in-N' mode (tiny cout): N = kw*cout, shifted sum in the epilogue;
                                      // tiles advance by w_step = 128-kw+1 output pixels
  int32_t thin;                       // host only: launch the thin instantiation (16 epilogue warps, 4 accumulator stages)
  int32_t b_res;                      // thin: ALL weight K blocks of the (single) N tile stay resident in smem (loaded once
                                      // per CTA); the ring then carries A tiles only -- one TMA per K block instead of two
  int32_t dbg;                        // debugging bits (env HFC_DBG): 1 skip stores, 2 skip norm stats, 4 skip epilogue body
  int32_t wide_boff;                  // debugging: set the descriptor base-offset field for shifted starts
  int32_t winflat;                    // window packing served from a PLAIN pixel segment (un-swizzled descriptor with
                                      // overlapping rows) instead of an 8x inflated window tile; tile = 128 px of a row
  int32_t wide, kw;                   // 'wide' mode: row-resident A halo (128+kw-1 pixels), resident weights
  int32_t a_region;                   // bytes per A stage (wide mode: halo row rounded up to 1024)
  int32_t grid_h, grid_w, batch;
  int32_t sh, sw;                     // input coordinate scale
  int32_t ih0, iw0;                   // input coordinate offset (materialised border)
  int32_t out_mode, osh, osw, ooh, oow;
  int32_t out_h, out_w;
  int32_t out_cpad;
  int32_t out_pt, out_pl, out_pb, out_pr;
  int32_t out_reflect;
  int32_t cout;
  int32_t act, norm;
  float eps;
  uint32_t fmt_a, fmt_b;              // operand formats: 0 fp16, 1 bf16
  int32_t gemm;                       // plain GEMM mode: A rows = 128 consecutive rows of a [M][K] matrix
  int32_t k_splits, kb_per_split;     // split-K (gemm mode): tile index carries the K range; fp32 atomics out
  int32_t out_atomic;                 // NHWC_F32 output accumulated with atomicAdd
  // norm == 2 ("wide" fused ChannelNorm, pair + nsub 2 + cn 2): the channel row of a pixel is spread over the four TMEM
  // halves of two CTAs; statistics are exchanged through distributed shared memory, then y = act(norm(x)) + res1 + res2
  // goes to fp32 rows (out_f32, optional) and to the bordered fp16 buffer (out, optional)
  const float* res1;
  const float* res2;
  float* out_f32;
  int32_t ld_res, ld_f32;
  const float* bias;
  const float* gamma;
  const float* beta;
  void* out;
  int8_t tap_dh[kMaxTaps];
  int8_t tap_dw[kMaxTaps];
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == HFC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HFC_ACT_LEAKY02) return v > 0.f ? v : 0.2f * v;
  return v;
}

// Division of a tile index by a runtime constant (tile counts, < 2^11) as a 64-bit multiply + shift: exact for n * d < 2^40.
// The per-tile decode (work item -> K split, N group, tile column / row / image) is 7 integer divisions; as plain `/` and
// `%` by runtime values that is ~300 dependent instructions per tile in the one-warp TMA producer -- ~0.8 us, more than
// the whole K loop of a short-K tile (G.up4: t_tile = 1.5 us + 0.25 us x K blocks before this).
struct FastDiv { unsigned long long m; int d; };
__device__ __forceinline__ FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.m = d > 1 ? ((1ull << 40) + static_cast<unsigned long long>(d) - 1ull) / static_cast<unsigned long long>(d) : 0ull;
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return f.d == 1 ? n : static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(n)) * f.m) >> 40);
}

// Rows / columns of the bordered NHWC output that receive pixel (oh, ow): itself plus its mirror images in the reflected
// border (ReflectionPad2d of the NEXT conv, materialised by the producer).  Scalars with -1 = none, never an indexed local
// array: with ~220 KB of dynamic shared memory there is next to no L1, so every local-memory access of the old
// `int rows[3], cols[3]` was an L2 round trip -- several per tile, on the critical path of every fused epilogue.
struct BorderDst { int r0, r1, r2, c0, c1, c2; };
__device__ __forceinline__ BorderDst border_dst(const ConvKernelParams& p, int oh, int ow) {
  BorderDst b;
  b.r0 = oh + p.out_pt; b.c0 = ow + p.out_pl;
  b.r1 = b.r2 = b.c1 = b.c2 = -1;
  if (p.out_reflect) {
    if (oh >= 1 && oh <= p.out_pt) b.r1 = p.out_pt - oh;
    if (oh <= p.out_h - 2 && oh >= p.out_h - 1 - p.out_pb) b.r2 = p.out_pt + 2 * (p.out_h - 1) - oh;
    if (ow >= 1 && ow <= p.out_pl) b.c1 = p.out_pl - ow;
    if (ow <= p.out_w - 2 && ow >= p.out_w - 1 - p.out_pr) b.c2 = p.out_pl + 2 * (p.out_w - 1) - ow;
  }
  return b;
}
// body(row, col) for every target of b (fully unrolled: the selects are on compile-time indices)
template <typename F>
__device__ __forceinline__ void for_each_dst(const BorderDst& b, F&& body) {
#pragma unroll
  for (int ri = 0; ri < 3; ++ri) {
    const int rr = ri == 0 ? b.r0 : (ri == 1 ? b.r1 : b.r2);
    if (rr < 0) continue;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const int cc = ci == 0 ? b.c0 : (ci == 1 ? b.c1 : b.c2);
      if (cc < 0) continue;
      body(rr, cc);
    }
  }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// kPair = true is the CTA-pair (cta_group::2) instantiation; it must be launched with clusters of 2 or 4
// (ptxas marks a kernel that contains cta_group::2 instructions as cluster-only), so the single-CTA path
// is a separate instantiation.
// kNsub: N tiles per work item (2 only with kPair: one N tile in each 256-column TMEM half, see tile_and_stages); a
// template parameter so that the single-tile instantiations keep their fully unrolled issue / epilogue loops
// kWideNorm (only with kPair, kNsub == 2 and a 2 x 2 cluster): the epilogue of the "wide" fused ChannelNorm (norm == 2); a
// separate instantiation so that the hot <true, 2, false> kernel keeps its register allocation (162, no spills).
// kThin (single-CTA MMA, one N tile of <= 64 columns, NHWC fp16 output): the two groups of four epilogue warps serve
// DIFFERENT accumulator stages concurrently, each thread holds its pixel's whole channel row in registers (one TMEM pass, no
// cross-warp exchange, no named barrier) and frees the TMEM stage right after the load -- for the big-map 60-channel layers
// whose tiles are epilogue-latency-bound (E1 355 us, G.up4 384 us against ~50 us of HBM time).  Own instantiation.
template <bool kPair, int kNsub, bool kWideNorm = false, bool kThin = false>
__global__ void __launch_bounds__(kThin ? kThinThreads : kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128B-swizzled tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = p.b_rows * kBlockK * 2;
  const int stage_bytes = p.wide ? p.a_region : (p.b_res ? kABytes : kABytes + kNsub * b_bytes);
  // wide mode: every weight sub-tile (num_kb * kw of them) stays resident behind the A ring; b_res: every K block
  uint8_t* w_res = smem + p.stages * stage_bytes;
  const int w_res_bytes = p.wide ? p.num_kb * p.kw * b_bytes : (p.b_res ? p.num_kb * b_bytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_res + w_res_bytes);
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + kMaxStages;       // [stages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;   // [4] (2 used outside the thin instantiation)
  uint64_t* tempty_bar = tfull_bar + 4;          // [4]
  uint64_t* wfull_bar = tempty_bar + 4;          // [1] resident weights landed (wide mode)
  uint64_t* xchg_bar = bars + 25;                // [1] statistics of the peer CTA landed (norm == 2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  float* s_par = reinterpret_cast<float*>(bars + 32);  // [2][3][kParamStride] bias / gamma / beta
  float* s_red = s_par + 2 * 3 * kParamStride;         // [2 acc stages][2 (sum, ssq)][2 warps][128 rows]
  uint32_t* s_kbt = reinterpret_cast<uint32_t*>(s_red + 2 * 512);   // [num_kb] producer table: chunk | dw << 16 | dh << 24
  float* s_tapn = s_red + 2 * 512 + kMaxKb;           // [2 acc stages][128][kTapnLd] (tap-in-N mode only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Cluster geometry: CTA (m_idx, n_idx) of a cm x cn cluster works on M tile mg*cm+m_idx and N tile
  // ng*cn+n_idx.  The A tile is shared by the cn CTAs of a cluster row, the B tile by the cm CTAs of a
  // column: each CTA fetches 1/cn of A and 1/cm of B and TMA-multicasts its slice to the sharers, which
  // divides the L2->SM operand traffic (the measured limiter, ~6.3 KB/clk chip-wide) by up to 2.
  const int csize = p.cm * p.cn;
  const uint32_t crank = csize > 1 ? cluster_ctarank() : 0u;
  const int m_idx = static_cast<int>(crank) % p.cm;
  const int n_idx = static_cast<int>(crank) / p.cm;
  uint16_t mask_a = 0, mask_b = 0;  // CTAs sharing my A tile (same m_idx) / my B tile (same n_idx)
  for (int j = 0; j < p.cn; ++j) mask_a |= static_cast<uint16_t>(1u << (m_idx + p.cm * j));
  for (int i = 0; i < p.cm; ++i) mask_b |= static_cast<uint16_t>(1u << (i + p.cm * n_idx));
  const uint16_t mask_e = mask_a | mask_b;  // producers that write into my stages == consumers I must release

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      // arrivals that free a stage: one commit per CTA (pair: per pair leader) that reads what lands here
      mbar_init(&empty_bar[s], static_cast<uint32_t>(kPair ? p.cn : p.cm + p.cn - 1));
    }
    for (int s = 0; s < (kThin ? kThinAccStages : 2); ++s) {
      mbar_init(&tfull_bar[s], 1);
      // pair mode: the leader's MMA warp waits for the epilogue warps of BOTH CTAs; thin: the 4 warps of the stage's group
      mbar_init(&tempty_bar[s], kThin ? 4u : (kPair ? 2 : 1) * (kEpiThreads / 32));
    }
    mbar_init(wfull_bar, 1);
    if constexpr (kWideNorm) mbar_init(xchg_bar, kBlockM);       // one remote arrival per pixel row of the peer CTA
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (kPair) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts / arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int m_groups = (tiles_m + p.cm - 1) / p.cm;
  const int n_groups = (p.n_tiles + p.cn * kNsub - 1) / (p.cn * kNsub);
  const int total_ctiles = m_groups * n_groups * p.k_splits;   // work items per cluster (x K splits in gemm mode)
  const FastDiv fd_ks = make_fastdiv(p.k_splits), fd_ng = make_fastdiv(n_groups), fd_tw = make_fastdiv(p.tiles_w),
                fd_th = make_fastdiv(p.tiles_h);
  const int cid = blockIdx.x / csize;
  const int ncl = gridDim.x / csize;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int a_slice_bytes = kABytes / p.cn;
      const int b_slice_rows = p.block_n / p.cm;
      if (p.wide) {
        // weights: loaded once, resident for the whole kernel (n_tiles == 1 in this mode)
        mbar_arrive_expect_tx(wfull_bar, static_cast<uint32_t>(w_res_bytes));
        for (int j = 0; j < p.num_kb * p.kw; ++j)
          tma_load_2d(w_res + j * b_bytes, &tmap_b, wfull_bar, j * kBlockK, 0);
        const uint32_t a_bytes = static_cast<uint32_t>((kBlockM + p.kw - 1) * kBlockK * 2);
        for (int ct = cid; ct < total_ctiles; ct += ncl) {
          int mt = ct;
          const int twi = mt % p.tiles_w;
          mt /= p.tiles_w;
          const int thi = mt % p.tiles_h;
          const int tni = mt / p.tiles_h;
          for (int kb = 0; kb < p.num_kb; ++kb) {   // one K block = one filter row: halo row of 128+kw-1 pixels
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], a_bytes);
            tma_load_4d(smem + s * stage_bytes, &tmap_a, &full_bar[s], 0, twi * kBlockM + p.iw0 + p.tap_dw[kb],
                        thi + p.ih0 + p.tap_dh[kb], tni);
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
    if (!p.wide) {
      // The whole warp walks the loop (warp-uniform control flow and operands; one elected lane issues).  The first version
      // ran it on lane 0 alone with the tap / chunk bookkeeping and two constant-bank table look-ups per K block: ~130
      // dependent single-thread instructions = ~0.45 us per K BLOCK (ncu, profiles/r02_ncu_bigmap_thin.json: the
      // epilogue warps of E1 sat in the accumulator-full wait 62 % of the time) -- 3.3 us per tile for the 7-block tiles
      // of E1 / G3 and as long as the MMAs of one K block of the 960-channel residual convolution.  Now the per-K-block
      // coordinates come from a table built once per CTA.
      if (!p.gemm)
        for (int kb = lane; kb < p.num_kb; kb += 32) {
          const int tap = kb / p.c_chunks, chunk = kb - tap * p.c_chunks;
          s_kbt[kb] = static_cast<uint32_t>(chunk) | (static_cast<uint32_t>(static_cast<uint8_t>(p.tap_dw[tap])) << 16) |
                      (static_cast<uint32_t>(static_cast<uint8_t>(p.tap_dh[tap])) << 24);
        }
      __syncwarp();
      if (p.b_res && elect_one()) {            // (n_tiles == 1, no cluster: checked on the host)
        mbar_arrive_expect_tx(wfull_bar, static_cast<uint32_t>(w_res_bytes));
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(w_res + kb * b_bytes, &tmap_b, wfull_bar, kb * kBlockK, 0);
      }
      __syncwarp();
      int s = 0;
      uint32_t ph = 0;
      const int a_slice_bytes = kABytes / p.cn;
      const int b_slice_rows = p.block_n / p.cm;
      const uint32_t tx_bytes = static_cast<uint32_t>(kPair ? 2 * (stage_bytes - p.tx_short) : stage_bytes - p.tx_short);
      const uint32_t kbt_addr = smem_u32(s_kbt);
      const bool use_table = !p.gemm;
      for (int ct = cid; ct < total_ctiles; ct += ncl) {
        const int ctile = fdiv(ct, fd_ks);
        const int ks = ct - ctile * p.k_splits;  // K split (gemm mode), else 0
        const int cg = fdiv(ctile, fd_ng);
        const int nt = ((ctile - cg * n_groups) * p.cn + n_idx) * kNsub;
        const int mt0 = cg * p.cm + m_idx;
        const int mt1 = fdiv(mt0, fd_tw);
        const int twi = mt0 - mt1 * p.tiles_w;
        const int tni = fdiv(mt1, fd_th);        // >= tiles_n for padding tiles: every row is out of bounds (zero fill)
        const int thi = mt1 - tni * p.tiles_h;
        const int w_base = (p.tapn ? twi * p.w_step : twi * p.tw * p.sw) + p.iw0;
        // my slice of the shared A tile: rows [n_idx*128/cn, (n_idx+1)*128/cn)
        const int h_base = (thi * p.th + (p.a_split_n ? 0 : n_idx * (p.th / p.cn))) * p.sh + p.ih0;
        const int n_base = tni * p.tn + (p.a_split_n ? n_idx * (p.tn / p.cn) : 0);
        const int kb0 = ks * p.kb_per_split;     // split-K only exists with a single tap (gemm mode)
        const int b_row = kPair ? nt * p.block_n + m_idx * p.b_rows : nt * p.block_n;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          int c0 = (kb0 + kb) * kBlockK, cw = w_base, chh = h_base;
          if (use_table) {
            uint32_t e;
            asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(e) : "r"(kbt_addr + 4u * static_cast<uint32_t>(kb)));
            c0 = static_cast<int>(e & 0xffffu) * kBlockK;
            cw += static_cast<int8_t>(e >> 16);
            chh += static_cast<int8_t>(e >> 24);
          }
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + kABytes;
          if (elect_one()) {
            if constexpr (kPair) {
              // both CTAs of the pair fill their own stage; all bytes are accounted on the LEADER's barrier
              if (m_idx == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
              if (p.cn > 1)
                tma_load_4d_pair_mc(sa + n_idx * a_slice_bytes, &tmap_a, &full_bar[s], c0, cw, chh, n_base, mask_a);
              else
                tma_load_4d_pair(sa, &tmap_a, &full_bar[s], c0, cw, chh, n_base);
#pragma unroll
              for (int j = 0; j < kNsub; ++j)
                tma_load_2d_pair(sb + j * b_bytes, &tmap_b, &full_bar[s], (kb0 + kb) * kBlockK, b_row + j * p.block_n);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
              if (csize > 1) {
                tma_load_4d_mc(sa + n_idx * a_slice_bytes, &tmap_a, &full_bar[s], c0, cw, chh, n_base, mask_a);
                tma_load_2d_mc(sb + m_idx * b_slice_rows * (kBlockK * 2), &tmap_b, &full_bar[s], (kb0 + kb) * kBlockK,
                               nt * p.block_n + m_idx * b_slice_rows, mask_b);
              } else {
                tma_load_4d(sa, &tmap_a, &full_bar[s], c0, cw, chh, n_base);
                if (!p.b_res) tma_load_2d(sb, &tmap_b, &full_bar[s], (kb0 + kb) * kBlockK, b_row);
              }
            }
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && kPair) {
    // ===================== MMA issuer, CTA pair: only the leader (even rank) issues =====================
    if constexpr (kPair) if (m_idx == 0) {
      const uint32_t idesc = make_idesc_f16(p.fmt_a, p.fmt_b, 2 * kBlockM, static_cast<uint32_t>(p.block_n));
      const uint16_t mask_pair = static_cast<uint16_t>(3u << (2 * n_idx));
      const uint16_t mask_all = static_cast<uint16_t>((1u << csize) - 1u);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      const int acc_stages = kNsub == 2 ? 1 : 2;   // nsub == 2: both TMEM halves belong to one work item
      for (int ct = cid; ct < total_ctiles; ct += ncl) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          // operands are warp-uniform and computed by the whole warp; ONE elected lane issues.  (Inside `if (lane == 0)` the
          // compiler cannot prove uniformity and wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST retry loop: ~100
          // cycles per MMA, i.e. as long as the MMAs of a K block themselves.)
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kNsub; ++j) {
              const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + kABytes + j * b_bytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16_pair(d_tmem + j * kAccStride, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair_mc(&empty_bar[s], mask_all);   // every CTA that writes into this pair's stages
            if (kb == p.num_kb - 1) umma_commit_pair_mc(&tfull_bar[as], mask_pair);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (++as == acc_stages) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_f16(p.fmt_a, p.fmt_b, kBlockM, static_cast<uint32_t>(p.block_n));
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    if (p.wide || p.b_res) mbar_wait(wfull_bar, 0);
    for (int ct = cid; ct < total_ctiles; ct += ncl) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (kThin ? kThinAccStride : kAccStride);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0 && p.wide) {
          // A operand for filter column t = the halo row shifted by t pixels: the descriptor start moves by
          // t*128 B inside the 1024 B swizzle atom, which the base-offset field (bits 49..51) accounts for.
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t w_addr = smem_u32(w_res) + static_cast<uint32_t>(kb * p.kw * b_bytes);
          for (int t = 0; t < p.kw; ++t) {
            // The 128B swizzle is a function of the absolute smem address bits (which is also why the +32 B
            // K advance works), so a start address that is only 128 B aligned needs no base offset.
            const uint64_t a_desc = make_sw128_kmajor_desc(a_addr + t * 128) |
                                    (p.wide_boff ? (static_cast<uint64_t>(t & 7) << 49) : 0ull);
            const uint64_t b_desc = make_sw128_kmajor_desc(w_addr + t * b_bytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | t | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == p.num_kb - 1) umma_commit(&tfull_bar[as]);
        } else if (!p.wide) {
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          // winflat: K = 16 covers two pixels of the window -> the start moves by 2 x 16 B per step, as it does (by
          // 32 B) inside the 128 B swizzle row of the regular layout
          const uint64_t a_desc = p.winflat ? make_nosw_window_desc(a_addr) : make_sw128_kmajor_desc(a_addr);
          const uint64_t b_desc = make_sw128_kmajor_desc(p.b_res ? smem_u32(w_res) + static_cast<uint32_t>(kb * b_bytes)
                                                                   : a_addr + kABytes);
          if (elect_one()) {               // warp-uniform operands, one issuing lane (see the pair issuer)
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (encoded >>4)
            umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (csize > 1) umma_commit_mc(&empty_bar[s], mask_e);
          else umma_commit(&empty_bar[s]);
          if (kb == p.num_kb - 1) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (++as == (kThin ? kThinAccStages : 2)) { as = 0; aph ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quadrant: both own the same 32 pixel rows and take alternate 16-column
    // chunks, so every scheduler has two epilogue warps to interleave (a lone warp issues one dependent
    // instruction every ~5 cycles, which made the 128-wide epilogues the bottleneck of the big-map layers).
    if constexpr (kThin) {
      // Thin epilogue: 16 warps = 4 groups x 4 TMEM lane quadrants.  Group g owns accumulator stage g and the work items
      // it % 4 == g, so FOUR tiles are in their epilogue at any time (the generic epilogue below puts all its warps on
      // one tile: ~1 200 dependent instructions per warp and tile at two warps per scheduler made E1 / G.up4 / G3
      // 3.5 us per tile, profiles/r02_ncu_bigmap_layers.json).  Thread == pixel; the ChannelNorm statistics are an
      // in-thread one-pass shifted sum over the row in TMEM, the second pass re-reads TMEM, normalises and stores: no
      // cross-warp exchange, no CTA-wide barrier.  Per-channel parameters come from shared memory as ld.shared.v4.
      const int q = warp & 3;            // TMEM lane quadrant this warp may access
      const int grp = (warp - 2) >> 2;   // epilogue group == accumulator stage
      const int m = q * 32 + lane;
      const int et = threadIdx.x - 64;
      float* s_bias_raw = s_par + 3 * kParamStride;            // tap-in-N: bias per output channel (cout <= 32 there)
      for (int i = et; i < p.block_n; i += kThinEpiThreads) {  // one N tile: parameters staged once
        const bool real = i < p.cout;
        s_bias_raw[i] = (real && p.bias) ? __ldg(p.bias + i) : 0.f;
        s_par[i] = (real && p.bias && !p.tapn) ? __ldg(p.bias + i) : 0.f;
        s_par[kParamStride + i] = (real && p.norm) ? __ldg(p.gamma + i) : 0.f;
        s_par[2 * kParamStride + i] = (real && p.norm) ? __ldg(p.beta + i) : 0.f;
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(kThinEpiThreads) : "memory");
      const uint32_t sp_bias = smem_u32(s_par), sp_gamma = sp_bias + kParamStride * 4, sp_beta = sp_bias + 2 * kParamStride * 4;
      auto lds4 = [](uint32_t addr) {
        float4 r;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
        return r;
      };
      const int twi_in = m % p.tw;
      const int thi_in = (m / p.tw) % p.th;
      const int tni_in = m / (p.tw * p.th);
      const int Hp = p.out_h + p.out_pt + p.out_pb;
      const int Wp = p.out_w + p.out_pl + p.out_pr;
      const float inv_c = 1.f / static_cast<float>(p.cout);
      int it = 0;
      for (int ct = cid; ct < total_ctiles; ct += ncl, ++it) {
        if ((it & (kThinAccStages - 1)) != grp) continue;
        const uint32_t aph = static_cast<uint32_t>(it / kThinAccStages) & 1u;
        const int mt0 = fdiv(ct, fd_ng) * p.cm + m_idx;
        const int mt1 = fdiv(mt0, fd_tw);
        const int twi = mt0 - mt1 * p.tiles_w;
        const int tni = fdiv(mt1, fd_th);
        const int thi = mt1 - tni * p.tiles_h;
        const int gw = (p.tapn ? twi * p.w_step : twi * p.tw) + twi_in;
        const int gh = thi * p.th + thi_in;
        const int n = tni * p.tn + tni_in;
        const int oh = gh * p.osh + p.ooh;
        const int ow = gw * p.osw + p.oow;
        const bool valid = (tni_in < p.tn) && (n < p.batch) && (gh < p.grid_h) && (gw < p.grid_w) && (oh < p.out_h) &&
                           (ow < p.out_w) && (!p.tapn || m < p.w_step);
        mbar_wait(&tfull_bar[grp], aph);
        tc_fence_after();
        const uint32_t t_row = tmem_base + grp * kThinAccStride + (static_cast<uint32_t>(q * 32) << 16);
        if (p.tapn) {
          // 'tap-in-N' (7x7 60 -> 3 head): column t*cout + co of pixel row m is the partial product of filter column t;
          // out[w][co] = bias[co] + sum_t D[w + t][t*cout + co].  The group stages its tile in ITS quarter of s_tapn and
          // synchronises on its own named barrier; the accumulator stage is free as soon as the row is in registers.
          uint32_t v[2][16];
          tmem_ld16(t_row, v[0]);
          if (p.block_n > 16) tmem_ld16(t_row + 16, v[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[grp]);
          float* st = s_tapn + grp * (kBlockM * kTapnLd);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (16 * c < p.block_n) {
#pragma unroll
              for (int j = 0; j < 16; ++j) st[m * kTapnLd + 16 * c + j] = __uint_as_float(v[c][j]);
            }
          asm volatile("bar.sync %0, 128;\n" ::"r"(3 + grp) : "memory");
          if (valid) {
            float* dst = reinterpret_cast<float*>(p.out);
            const size_t plane = static_cast<size_t>(p.out_h) * p.out_w;
            for (int co = 0; co < p.cout; ++co) {
              float acc = s_bias_raw[co];
              for (int t = 0; t < p.kw; ++t) acc += st[(m + t) * kTapnLd + t * p.cout + co];
              dst[(static_cast<size_t>(n) * p.cout + co) * plane + static_cast<size_t>(oh) * p.out_w + ow] =
                  apply_act(acc, p.act);
            }
          }
          // the next tile of this group overwrites the staging buffer: everyone must be done reading it
          asm volatile("bar.sync %0, 128;\n" ::"r"(3 + grp) : "memory");
          continue;
        }
        // ---- pass 1: ChannelNorm statistics, one pass with shifted sums (shift = the row's first channel), the padding
        // columns (acc = bias = 0, d = -shift) removed analytically.  Packed fp32 (two channels per instruction).
        float mean = 0.f, rstd = 1.f;
        if (p.norm) {
          uint32_t va[16];
          tmem_ld16(t_row, va);
          tmem_ld_wait();
          const float shift = __uint_as_float(va[0]) + lds4(sp_bias).x;
          const f32x2 nshift = pk2(-shift, -shift);
          f32x2 sd = pk2(0.f, 0.f), sq = pk2(0.f, 0.f);
          for (int c0 = 0; c0 < p.block_n; c0 += 16) {        // (the other three warps of this scheduler hide the TMEM latency)
            if (c0) {
              tmem_ld16(t_row + c0, va);
              tmem_ld_wait();
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b4 = lds4(sp_bias + 4 * (c0 + 4 * j4));
              const f32x2 d01 = add2x(add2x(pk2(__uint_as_float(va[4 * j4 + 0]), __uint_as_float(va[4 * j4 + 1])), pk2(b4.x, b4.y)), nshift);
              const f32x2 d23 = add2x(add2x(pk2(__uint_as_float(va[4 * j4 + 2]), __uint_as_float(va[4 * j4 + 3])), pk2(b4.z, b4.w)), nshift);
              sd = add2x(sd, d01); sq = fma2x(d01, d01, sq);
              sd = add2x(sd, d23); sq = fma2x(d23, d23, sq);
            }
          }
          float sd0, sd1, sq0, sq1;
          unpk2(sd, sd0, sd1);
          unpk2(sq, sq0, sq1);
          const float npad = static_cast<float>(p.block_n - p.cout);
          const float sds = (sd0 + sd1) + npad * shift;
          const float sqs = (sq0 + sq1) - npad * shift * shift;
          const float mean_d = sds * inv_c;
          mean = shift + mean_d;
          rstd = rsqrtf(fmaxf(sqs - sds * mean_d, 0.f) / static_cast<float>(p.cout - 1) + p.eps);
        }
        // ---- pass 2: re-read the row, normalise, activate, store (NHWC fp16, with the reflected border of the next conv)
        BorderDst bd = border_dst(p, oh, ow);
        const int nr = valid ? 1 : 0;
        __half* dst0 = reinterpret_cast<__half*>(p.out) + ((static_cast<size_t>(n) * Hp + bd.r0) * Wp + bd.c0) * p.out_cpad;
        const bool interior = bd.r1 < 0 && bd.r2 < 0 && bd.c1 < 0 && bd.c2 < 0;
        // one instantiation per (norm, ReLU) combination: no runtime switch inside the element loops
        auto pass2 = [&](auto norm_tag, auto relu_tag) {
          constexpr bool kNorm = decltype(norm_tag)::value, kRelu = decltype(relu_tag)::value;
          const f32x2 nmean = pk2(-mean, -mean), rs2 = pk2(rstd, rstd);
          uint32_t v[16];
          for (int c0 = 0; c0 < p.block_n; c0 += 16) {
            tmem_ld16(t_row + c0, v);
            tmem_ld_wait();
            if (c0 + 16 >= p.block_n) {      // the last chunk is in registers: hand the accumulator stage back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[grp]);
            }
            uint32_t h[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b4 = lds4(sp_bias + 4 * (c0 + 4 * j4));
              f32x2 x01 = add2x(pk2(__uint_as_float(v[4 * j4 + 0]), __uint_as_float(v[4 * j4 + 1])), pk2(b4.x, b4.y));
              f32x2 x23 = add2x(pk2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), pk2(b4.z, b4.w));
              if constexpr (kNorm) {
                const float4 g4 = lds4(sp_gamma + 4 * (c0 + 4 * j4));
                const float4 e4 = lds4(sp_beta + 4 * (c0 + 4 * j4));
                x01 = fma2x(mul2x(pk2(g4.x, g4.y), rs2), add2x(x01, nmean), pk2(e4.x, e4.y));
                x23 = fma2x(mul2x(pk2(g4.z, g4.w), rs2), add2x(x23, nmean), pk2(e4.z, e4.w));
              }
              float y0, y1, y2, y3;
              unpk2(x01, y0, y1);
              unpk2(x23, y2, y3);
              if constexpr (kRelu) {
                y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f);
              } else {
                y0 = apply_act(y0, p.act); y1 = apply_act(y1, p.act); y2 = apply_act(y2, p.act); y3 = apply_act(y3, p.act);
              }
              h[2 * j4] = pack_half2(y0, y1);
              h[2 * j4 + 1] = pack_half2(y2, y3);
            }
            if (nr == 0 || c0 >= p.out_cpad) continue;
            const uint4 lo = make_uint4(h[0], h[1], h[2], h[3]), hi = make_uint4(h[4], h[5], h[6], h[7]);
            const bool two = c0 + 8 < p.out_cpad;
            if (interior) {
              reinterpret_cast<uint4*>(dst0 + c0)[0] = lo;
              if (two) reinterpret_cast<uint4*>(dst0 + c0)[1] = hi;
            } else {
              for_each_dst(bd, [&](int rr, int cc) {
                __half* dst = reinterpret_cast<__half*>(p.out) + ((static_cast<size_t>(n) * Hp + rr) * Wp + cc) * p.out_cpad + c0;
                reinterpret_cast<uint4*>(dst)[0] = lo;
                if (two) reinterpret_cast<uint4*>(dst)[1] = hi;
              });
            }
          }
        };
        if (p.norm && p.act == HFC_ACT_RELU) pass2(std::true_type{}, std::true_type{});
        else if (p.norm) pass2(std::true_type{}, std::false_type{});
        else pass2(std::false_type{}, std::false_type{});
        if (nr)
          for (int c = p.block_n; c < p.out_cpad; c += 8)        // channel padding the N tile does not cover
            for_each_dst(bd, [&](int rr, int cc) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + ((static_cast<size_t>(n) * Hp + rr) * Wp + cc) * p.out_cpad + c) =
                  make_uint4(0, 0, 0, 0);
            });
      }
    } else {
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int hsel = (warp - 2) >> 2;  // which of the two warps of the quadrant
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 64;   // 0..255 among the epilogue threads
    const int twi_in = m % p.tw;
    const int thi_in = (m / p.tw) % p.th;
    const int tni_in = m / (p.tw * p.th);
    const float inv_c = 1.f / static_cast<float>(p.cout);
    const int act_now = p.tapn ? HFC_ACT_NONE : p.act;  // tap-in-N applies the activation after the shifted sum
    int as = 0;
    uint32_t aph = 0;
    int cur_nt = -1, pbuf = 1;
    [[maybe_unused]] uint32_t xph = 0; // phase of xchg_bar == index of the exchange buffer (norm == 2)
    for (int ct = cid; ct < total_ctiles; ct += ncl) {
      const int ctile = fdiv(ct, fd_ks);
      const int cg = fdiv(ctile, fd_ng);
      const int nt0 = ((ctile - cg * n_groups) * p.cn + n_idx) * kNsub;
      const int mt0 = cg * p.cm + m_idx;
      const int mt1 = fdiv(mt0, fd_tw);
      const int twi = mt0 - mt1 * p.tiles_w;
      const int tni = fdiv(mt1, fd_th);
      const int thi = mt1 - tni * p.tiles_h;
      const int gw = (p.tapn ? twi * p.w_step : twi * p.tw) + twi_in;
      const int gh = thi * p.th + thi_in;
      const int n = tni * p.tn + tni_in;
      const int oh = gh * p.osh + p.ooh;
      const int ow = gw * p.osw + p.oow;
      const bool valid_px = (tni_in < p.tn) && (n < p.batch) && (gh < p.grid_h) && (gw < p.grid_w) && (oh < p.out_h) &&
                            (ow < p.out_w) && (!p.tapn || m < p.w_step);
      bool wide_norm_done = false;
      if constexpr (kWideNorm && kNsub == 2) {
        if (p.norm == 2) {
          // ============ ChannelNorm over a channel row held by TWO CTAs (4 N tiles: 2 TMEM halves here, 2 in the peer
          // pair of the 4-CTA cluster) -- replaces conv -> fp32 rows -> hfc_channelnorm for the 960-channel layers ======
          wide_norm_done = true;
          auto stage_params = [&](int nt) {
            if (nt != cur_nt) {
              cur_nt = nt;
              pbuf ^= 1;
              float* sp = s_par + pbuf * (3 * kParamStride);
              for (int i = et; i < p.block_n; i += kEpiThreads) {
                const int c = nt * p.block_n + i;
                const bool real = c < p.cout;
                sp[i] = (real && p.bias) ? __ldg(p.bias + c) : 0.f;
                sp[kParamStride + i] = real ? __ldg(p.gamma + c) : 0.f;
                sp[2 * kParamStride + i] = real ? __ldg(p.beta + c) : 0.f;
              }
              asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
            }
          };
          auto walk = [&](uint32_t t_row, auto&& body) {      // this warp's alternate 16-column chunks of one TMEM half
            uint32_t va[16], vb[16];
            int c0 = 16 * hsel;
            if (c0 < p.block_n) tmem_ld16(t_row + c0, va);
            while (c0 < p.block_n) {
              tmem_ld_wait();
              if (c0 + 32 < p.block_n) tmem_ld16(t_row + c0 + 32, vb);
              body(va, c0);
              c0 += 32;
              if (c0 >= p.block_n) break;
              tmem_ld_wait();
              if (c0 + 32 < p.block_n) tmem_ld16(t_row + c0 + 32, va);
              body(vb, c0);
              c0 += 32;
            }
          };
          auto walk1 = [&](uint32_t t_row, auto&& body) {     // same chunks, no TMEM prefetch (fewer live registers)
            uint32_t va[16];
            for (int c0 = 16 * hsel; c0 < p.block_n; c0 += 32) {
              tmem_ld16(t_row + c0, va);
              tmem_ld_wait();
              body(va, c0);
            }
          };
          // ---- phase 1: (count, mean, M2) of the channels this CTA holds, per pixel row
          float cnt_a = 0.f, mean_a = 0.f, m2_a = 0.f;
          for (int jsub = 0; jsub < 2; ++jsub) {
            const int nt = nt0 + jsub;
            stage_params(nt);
            const float* s_bias = s_par + pbuf * (3 * kParamStride);
            if (jsub == 0) {
              mbar_wait(&tfull_bar[as], aph);
              tc_fence_after();
            }
            const uint32_t t_row = tmem_base + jsub * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t first;
            tmem_ld1(t_row, first);
            tmem_ld_wait();
            const float shift = __uint_as_float(first) + s_bias[0];
            float sd = 0.f, sq = 0.f;
            walk(t_row, [&](const uint32_t (&v)[16], int c0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float dlt = (__uint_as_float(v[j]) + s_bias[c0 + j]) - shift;
                sd += dlt;
                sq = fmaf(dlt, dlt, sq);
              }
            });
            float* red = s_red;                          // first half of s_red (nsub == 2 has one accumulator stage)
            red[hsel * 128 + m] = sd;
            red[256 + hsel * 128 + m] = sq;
            asm volatile("bar.sync 2, %0;\n" ::"n"(kEpiThreads) : "memory");
            const int real = max(0, min(p.block_n, p.cout - nt * p.block_n));
            const float npad = static_cast<float>(p.block_n - real);   // padding columns hold exact zeros: d = -shift
            const float sd_t = red[m] + red[128 + m] + npad * shift;
            const float sq_t = red[256 + m] + red[384 + m] - npad * shift * shift;
            asm volatile("bar.sync 2, %0;\n" ::"n"(kEpiThreads) : "memory");   // red is rewritten by the next half
            const float cj = static_cast<float>(real);
            if (real > 0) {
              const float mean_j = shift + sd_t / cj;
              const float m2_j = fmaxf(sq_t - sd_t * sd_t / cj, 0.f);
              if (cnt_a == 0.f) {
                cnt_a = cj; mean_a = mean_j; m2_a = m2_j;
              } else {                                   // Chan et al.: merge two (count, mean, M2) summaries
                const float tot = cnt_a + cj, dl = mean_j - mean_a;
                m2_a = m2_a + m2_j + dl * dl * (cnt_a * cj / tot);
                mean_a = mean_a + dl * (cj / tot);
                cnt_a = tot;
              }
            }
          }
          // ---- exchange with the CTA of the other pair that holds the same pixel rows (same m_idx, other n_idx)
          float* xbuf = s_red + 512 + static_cast<int>(xph) * 256;     // [2 (mean, M2)][128], double-buffered by phase
          const uint32_t peer = crank ^ static_cast<uint32_t>(p.cm);
          if (hsel == 0) {
            st_cluster_f32(mapa_u32(smem_u32(xbuf + m), peer), mean_a);
            st_cluster_f32(mapa_u32(smem_u32(xbuf + 128 + m), peer), m2_a);
            mbar_arrive_remote_release(mapa_u32(smem_u32(xchg_bar), peer));
          }
          mbar_wait_acquire_cluster(xchg_bar, xph);
          const float mean_b = xbuf[m], m2_b = xbuf[128 + m];
          xph ^= 1u;
          const float ctot = static_cast<float>(p.cout), cnt_b = ctot - cnt_a;
          const float dl = mean_b - mean_a;
          const float mean = mean_a + dl * (cnt_b / ctot);
          const float m2 = m2_a + m2_b + dl * dl * (cnt_a * cnt_b / ctot);
          const float rstd = rsqrtf(fmaxf(m2, 0.f) / static_cast<float>(p.cout - 1) + p.eps);
          // ---- phase 2: normalise, activate, add the residual streams, write fp32 rows and / or the fp16 buffer
          const int Hp = p.out_h + p.out_pt + p.out_pb;
          const int Wp = p.out_w + p.out_pl + p.out_pr;
          const size_t pix = (static_cast<size_t>(n) * p.out_h + oh) * p.out_w + ow;
          for (int jsub = 0; jsub < 2; ++jsub) {
            const int nt = nt0 + jsub;
            stage_params(nt);
            const float* s_bias = s_par + pbuf * (3 * kParamStride);
            const float* s_gamma = s_bias + kParamStride;
            const float* s_beta = s_bias + 2 * kParamStride;
            const uint32_t t_row = tmem_base + jsub * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
            const int c_base = nt * p.block_n;
            walk1(t_row, [&](const uint32_t (&v)[16], int c0) {
              const int cc = c_base + c0;
              if (!valid_px || cc >= p.cout) return;       // cout is a multiple of 16 in this mode (checked on the host)
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float x = __uint_as_float(v[j]) + s_bias[c0 + j];
                f[j] = apply_act(fmaf(s_gamma[c0 + j] * rstd, x - mean, s_beta[c0 + j]), p.act);
              }
              if (p.res1) {
                const float4* r = reinterpret_cast<const float4*>(p.res1 + pix * p.ld_res + cc);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 t = r[j4];
                  f[4 * j4] += t.x; f[4 * j4 + 1] += t.y; f[4 * j4 + 2] += t.z; f[4 * j4 + 3] += t.w;
                }
              }
              if (p.res2) {
                const float4* r = reinterpret_cast<const float4*>(p.res2 + pix * p.ld_res + cc);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 t = r[j4];
                  f[4 * j4] += t.x; f[4 * j4 + 1] += t.y; f[4 * j4 + 2] += t.z; f[4 * j4 + 3] += t.w;
                }
              }
              if (p.out_f32) {
                float4* dst = reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_f32 + cc);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
              }
              if (p.out) {
                uint4 lo, hi;
                lo.x = pack_half2(f[0], f[1]);   lo.y = pack_half2(f[2], f[3]);
                lo.z = pack_half2(f[4], f[5]);   lo.w = pack_half2(f[6], f[7]);
                hi.x = pack_half2(f[8], f[9]);   hi.y = pack_half2(f[10], f[11]);
                hi.z = pack_half2(f[12], f[13]); hi.w = pack_half2(f[14], f[15]);
                const BorderDst bd = border_dst(p, oh, ow);     // recomputed here: a few integer ops, no long live range
                for_each_dst(bd, [&](int rr, int c2) {
                  __half* dst = reinterpret_cast<__half*>(p.out) + ((static_cast<size_t>(n) * Hp + rr) * Wp + c2) * p.out_cpad + cc;
                  reinterpret_cast<uint4*>(dst)[0] = lo;
                  reinterpret_cast<uint4*>(dst)[1] = hi;
                });
              }
            });
          }
        }
      }
      // nsub == 2 (CTA pairs): the work item owns both TMEM halves, one N tile in each
      if (!wide_norm_done)
      for (int jsub = 0; jsub < kNsub; ++jsub) {
      const int nt = nt0 + jsub;
      const bool valid = valid_px && (nt < p.n_tiles);
      const int c_base = nt * p.block_n;
      const bool last_nt = (nt == p.n_tiles - 1);

      // Per-channel epilogue parameters live in shared memory: with ~220 KB of dynamic smem the
      // L1 is (almost) gone, so per-element __ldg's would each be an L2 round trip.  Reloaded only
      // when the N tile changes; double-buffered so no thread can still be reading the old copy.
      if (nt != cur_nt) {
        cur_nt = nt;
        pbuf ^= 1;
        float* sp = s_par + pbuf * (3 * kParamStride);
        for (int i = et; i < p.block_n; i += kEpiThreads) {
          const int c = c_base + i;
          const bool real = c < p.cout;
          sp[i] = (real && p.bias && !p.tapn) ? __ldg(p.bias + c) : 0.f;
          sp[kParamStride + i] = (real && p.norm) ? __ldg(p.gamma + c) : 0.f;
          sp[2 * kParamStride + i] = (real && p.norm) ? __ldg(p.beta + c) : 0.f;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
      }
      const float* s_bias = s_par + pbuf * (3 * kParamStride);
      const float* s_gamma = s_bias + kParamStride;
      const float* s_beta = s_bias + 2 * kParamStride;

      if (jsub == 0) {     // after the parameter staging, which thereby overlaps the tail of the main loop
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after();
      }
      const uint32_t t_row = tmem_base + (kNsub == 2 ? jsub : as) * kAccStride + (static_cast<uint32_t>(q * 32) << 16);

      // Walks this warp's 16-column chunks (alternate chunks belong to the partner warp of the quadrant);
      // the TMEM load of the next chunk is in flight while the current one is processed.
      auto for_chunks = [&](auto&& body) {
        uint32_t va[16], vb[16];
        int c0 = 16 * hsel;
        if (c0 < p.block_n) tmem_ld16(t_row + c0, va);
        while (c0 < p.block_n) {
          tmem_ld_wait();
          if (c0 + 32 < p.block_n) tmem_ld16(t_row + c0 + 32, vb);
          body(va, c0);
          c0 += 32;
          if (c0 >= p.block_n) break;
          tmem_ld_wait();
          if (c0 + 32 < p.block_n) tmem_ld16(t_row + c0 + 32, va);
          body(vb, c0);
          c0 += 32;
        }
      };

      float mean = 0.f, rstd = 1.f;
      if (p.norm && !(p.dbg & 6)) {
        // ChannelNorm statistics in ONE pass over the channel row held in TMEM, using shifted sums
        // (shift = the row's first channel) so that sum(d^2) - sum(d)^2/C does not cancel:
        //   d = (acc + bias) - shift ;  mean = shift + sum(d)/C ;  var = (sum(d^2) - sum(d)^2/C)/(C-1).
        // Padding columns hold exact zeros (d = -shift) and are removed analytically.  The two warps of a
        // quadrant combine their partial sums through shared memory (double-buffered by accumulator stage).
        uint32_t first;
        tmem_ld1(t_row, first);
        tmem_ld_wait();
        const float shift = __uint_as_float(first) + s_bias[0];
        float sd4[4] = {0.f, 0.f, 0.f, 0.f}, sq4[4] = {0.f, 0.f, 0.f, 0.f};
        for_chunks([&](const uint32_t (&v)[16], int c0) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j4);
            const float d0 = (__uint_as_float(v[4 * j4 + 0]) + b4.x) - shift;
            const float d1 = (__uint_as_float(v[4 * j4 + 1]) + b4.y) - shift;
            const float d2 = (__uint_as_float(v[4 * j4 + 2]) + b4.z) - shift;
            const float d3 = (__uint_as_float(v[4 * j4 + 3]) + b4.w) - shift;
            sd4[0] += d0; sd4[1] += d1; sd4[2] += d2; sd4[3] += d3;
            sq4[0] = fmaf(d0, d0, sq4[0]); sq4[1] = fmaf(d1, d1, sq4[1]);
            sq4[2] = fmaf(d2, d2, sq4[2]); sq4[3] = fmaf(d3, d3, sq4[3]);
          }
        });
        float* red = s_red + as * 512;
        red[hsel * 128 + m] = (sd4[0] + sd4[1]) + (sd4[2] + sd4[3]);
        red[256 + hsel * 128 + m] = (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]);
        asm volatile("bar.sync 2, %0;\n" ::"n"(kEpiThreads) : "memory");
        const float npad = static_cast<float>(p.block_n - p.cout);
        const float sd = red[m] + red[128 + m] + npad * shift;
        const float sq = red[256 + m] + red[384 + m] - npad * shift * shift;
        const float mean_d = sd * inv_c;
        mean = shift + mean_d;
        rstd = rsqrtf(fmaxf(sq - sd * mean_d, 0.f) / static_cast<float>(p.cout - 1) + p.eps);
      }

      // target rows / cols of the (bordered) NHWC fp16 buffer
      const int Hp = p.out_h + p.out_pt + p.out_pb;
      const int Wp = p.out_w + p.out_pl + p.out_pr;

      if (!(p.dbg & 4))
      for_chunks([&](const uint32_t (&v)[16], int c0) {
        float f[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          // padding columns: acc = 0, bias = gamma = beta = 0  ->  exactly 0 after norm / activation
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j4);
          float x0 = __uint_as_float(v[4 * j4 + 0]) + b4.x, x1 = __uint_as_float(v[4 * j4 + 1]) + b4.y;
          float x2 = __uint_as_float(v[4 * j4 + 2]) + b4.z, x3 = __uint_as_float(v[4 * j4 + 3]) + b4.w;
          if (p.norm) {
            const float4 g4 = *reinterpret_cast<const float4*>(s_gamma + c0 + 4 * j4);
            const float4 e4 = *reinterpret_cast<const float4*>(s_beta + c0 + 4 * j4);
            x0 = fmaf(g4.x * rstd, x0 - mean, e4.x); x1 = fmaf(g4.y * rstd, x1 - mean, e4.y);
            x2 = fmaf(g4.z * rstd, x2 - mean, e4.z); x3 = fmaf(g4.w * rstd, x3 - mean, e4.w);
          }
          f[4 * j4 + 0] = apply_act(x0, act_now); f[4 * j4 + 1] = apply_act(x1, act_now);
          f[4 * j4 + 2] = apply_act(x2, act_now); f[4 * j4 + 3] = apply_act(x3, act_now);
        }
        const int cc = c_base + c0;
        if (p.tapn) {
          // stage the partial products D[pixel][tap*cout + co]; the shifted sum over taps follows below
          float* st = s_tapn + as * (kBlockM * kTapnLd) + m * kTapnLd + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) st[j] = f[j];
        } else if (!valid || ((p.dbg & 1) && f[0] != 12345.678f)) {
          // nothing to store for rows outside the image / batch (their A rows were zero-filled)
        } else if (p.out_mode == HFC_OUT_NHWC_F16) {
          if (cc < p.out_cpad) {
            uint4 lo, hi;
            lo.x = pack_half2(f[0], f[1]);   lo.y = pack_half2(f[2], f[3]);
            lo.z = pack_half2(f[4], f[5]);   lo.w = pack_half2(f[6], f[7]);
            hi.x = pack_half2(f[8], f[9]);   hi.y = pack_half2(f[10], f[11]);
            hi.z = pack_half2(f[12], f[13]); hi.w = pack_half2(f[14], f[15]);
            const BorderDst bd = border_dst(p, oh, ow);       // recomputed at the store site: no long live range
            for_each_dst(bd, [&](int rr, int c2) {
              __half* dst = reinterpret_cast<__half*>(p.out) + ((static_cast<size_t>(n) * Hp + rr) * Wp + c2) * p.out_cpad + cc;
              reinterpret_cast<uint4*>(dst)[0] = lo;
              if (cc + 8 < p.out_cpad) reinterpret_cast<uint4*>(dst)[1] = hi;
            });
          }
        } else if (p.out_mode == HFC_OUT_NHWC_F32) {
          if (cc < p.out_cpad) {
            float* dst = reinterpret_cast<float*>(p.out) +
                         ((static_cast<size_t>(n) * p.out_h + oh) * p.out_w + ow) * p.out_cpad + cc;
            if (p.out_atomic) {   // split-K partial sums
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cc + j < p.out_cpad) atomicAdd(dst + j, f[j]);
            } else {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                if (cc + 4 * j4 < p.out_cpad)
                  reinterpret_cast<float4*>(dst)[j4] =
                      make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
              }
            }
          }
        } else {  // NCHW fp32
          float* dst = reinterpret_cast<float*>(p.out);
          const size_t plane = static_cast<size_t>(p.out_h) * p.out_w;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = cc + j;
            if (c < p.cout)
              dst[(static_cast<size_t>(n) * p.cout + c) * plane + static_cast<size_t>(oh) * p.out_w +
                  ow] = f[j];
          }
        }
      });
      if (p.tapn) {
        // out[w, co] = bias[co] + sum_t D[w + t][t*cout + co]   (NCHW fp32 output)
        asm volatile("bar.sync 2, %0;\n" ::"n"(kEpiThreads) : "memory");
        if (valid && hsel == 0) {
          const float* st = s_tapn + as * (kBlockM * kTapnLd);
          float* dst = reinterpret_cast<float*>(p.out);
          const size_t plane = static_cast<size_t>(p.out_h) * p.out_w;
          for (int co = 0; co < p.cout; ++co) {
            float acc = p.bias ? __ldg(p.bias + co) : 0.f;
            for (int t = 0; t < p.kw; ++t) acc += st[(m + t) * kTapnLd + t * p.cout + co];
            dst[(static_cast<size_t>(n) * p.cout + co) * plane + static_cast<size_t>(oh) * p.out_w + ow] =
                apply_act(acc, p.act);
          }
        }
      }
      // zero the channel padding no N tile covers (e.g. cout 220 -> block_n 224 -> cpad 256)
      if (valid && last_nt && hsel == 0 && !p.tapn && !p.out_atomic && p.out_mode != HFC_OUT_NCHW_F32) {
        const int c_end = p.n_tiles * p.block_n;
        if (p.out_mode == HFC_OUT_NHWC_F16) {
          const BorderDst bd = border_dst(p, oh, ow);
          for (int c = c_end; c < p.out_cpad; c += 8)
            for_each_dst(bd, [&](int rr, int c2) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + ((static_cast<size_t>(n) * Hp + rr) * Wp + c2) * p.out_cpad + c) =
                  make_uint4(0, 0, 0, 0);
            });
        } else {
          float* dst = reinterpret_cast<float*>(p.out) +
                       ((static_cast<size_t>(n) * p.out_h + oh) * p.out_w + ow) * p.out_cpad;
          for (int c = c_end; c < p.out_cpad; c += 4)
            *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
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
    }  // !kThin
  }

  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // nobody exits while a peer may still arrive on its barriers
  if (warp == 1) {
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight packing: torch layout fp32 -> K-major [cout_rows][ktot] 16-bit, K ordered (tap, cin_pad)
// ------------------------------------------------------------------------------------------------
struct PackParams {
  int32_t cout, cin, kh, kw;
  int32_t cin_pad, ntaps, ktot, rows;
  int32_t cout_real;                  // tap-in-N: real cout (pp.cout then counts kw*cout rows)
  int32_t transposed, window, tapn, dgrad;
  uint32_t fmt;
  int8_t ky[kMaxTaps];
  int8_t kx[kMaxTaps];
};

__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                    uint16_t* __restrict__ out, const __grid_constant__ PackParams pp) {
  const float mul = scale ? *scale : 1.f;
  const size_t total = static_cast<size_t>(pp.rows) * pp.ktot;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i / pp.ktot);
    const int k = static_cast<int>(i % pp.ktot);
    float val = 0.f;
    if (co < pp.cout) {
      int tap, ci, ky, kx;
      bool ok;
      if (pp.tapn) {
        // row = filter column t * cout + co ; K = (filter row r, input channel)
        tap = k / pp.cin_pad;
        ci = k % pp.cin_pad;
        ky = pp.ky[tap];
        kx = co / pp.cout_real;
        ok = (ci < pp.cin) && (kx < pp.kw);
      } else if (pp.window) {
        tap = k / 64;
        const int r = k % 64;
        kx = r / pp.cin_pad;          // cin_pad == 8: 8 pixels x 8 channels per filter row
        ci = r % pp.cin_pad;
        ky = pp.ky[tap];
        ok = (kx < pp.kw) && (ci < pp.cin);
      } else {
        tap = k / pp.cin_pad;
        ci = k % pp.cin_pad;
        ky = pp.ky[tap];
        kx = pp.kx[tap];
        ok = ci < pp.cin;
      }
      if (ok) {
        const int co_w = pp.tapn ? co % pp.cout_real : co;
        const int cout_w = pp.tapn ? pp.cout_real : pp.cout;
        // dgrad of a stride-1 conv: this conv's (out=co_w, in=ci) weight is the forward weight W[ci][co_w] with the
        // filter flipped in both directions
        const size_t idx = pp.dgrad
                               ? ((static_cast<size_t>(ci) * cout_w + co_w) * pp.kh + (pp.kh - 1 - ky)) * pp.kw +
                                     (pp.kw - 1 - kx)
                           : pp.transposed
                               ? ((static_cast<size_t>(ci) * cout_w + co_w) * pp.kh + ky) * pp.kw + kx
                               : ((static_cast<size_t>(co_w) * pp.cin + ci) * pp.kh + ky) * pp.kw + kx;
        val = w[idx] * mul;
      }
    }
    uint16_t bits;
    if (pp.fmt == 0) {
      __half h = __float2half_rn(val);
      bits = *reinterpret_cast<uint16_t*>(&h);
    } else {
      __nv_bfloat16 h = __float2bfloat16_rn(val);
      bits = *reinterpret_cast<uint16_t*>(&h);
    }
    out[i] = bits;
  }
}

// Fast path for the plain layouts (no window / tap-in-N): one block per packed row.  The row's source values are read
// with consecutive threads on consecutive addresses (the whole [cin][kh*kw] slab of a conv2d filter, or kh*kw-float
// runs of a conv_transpose2d / dgrad filter), converted, transposed through shared memory and written as one
// contiguous K-major row.  The generic kernel above issues one 32 B sector request per ELEMENT (the gather stride is
// kh*kw floats), which made weight packing 1.8 ms of the training step.
struct PackRowParams {
  int32_t rows_real, cin, kh, kw, cin_pad, ntaps, ktot;
  int32_t strided, flip;     // strided: source is W[c][row][ky][kx] (conv_transpose2d / dgrad), else W[row][c][ky][kx]
  uint32_t fmt;
  int8_t ky[kMaxTaps];
  int8_t kx[kMaxTaps];
};

__device__ __forceinline__ uint16_t pack_cvt(float v, uint32_t fmt) {
  if (fmt == 0) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// KHKW: kh*kw as a compile-time constant (the index split is then a multiply-shift), 0 = runtime value.
// VEC (odd KHKW only, so that T == KHKW and the shared row is simply the converted source slab in source order): 16-byte
// global accesses on both sides -- float4 reads of a contiguous conv2d filter slab (4 independent scalar loads per thread
// for the kh*kw-float runs of a conv_transpose2d / dgrad filter) and one 16-byte store per 8 output channels.  The scalar
// version ran the 960 x 960 x 3 x 3 filters at ~2 TB/s: 34 dependent 4-byte loads per thread with every block of the grid
// resident at once, i.e. latency-bound, not bandwidth-bound.
template <int KHKW, bool VEC>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ w, const float* __restrict__ scale, uint16_t* __restrict__ out,
                 const __grid_constant__ PackRowParams pp) {
  extern __shared__ __align__(16) uint16_t s_pack[];  // [cin][T], T odd: conflict-free transposed reads
  const float mul = scale ? *scale : 1.f;
  const int r = blockIdx.x;
  const int khkw = KHKW ? KHKW : pp.kh * pp.kw;
  const int T = khkw | 1;
  const bool real = r < pp.rows_real;
  if (real) {
    const int total = pp.cin * khkw;
    const float* src_row = w + static_cast<size_t>(r) * (pp.strided ? khkw : total);
    const size_t c_stride = pp.strided ? static_cast<size_t>(pp.rows_real) * khkw : static_cast<size_t>(khkw);
    if constexpr (VEC) {
      static_assert(KHKW % 2 == 1, "the vector path needs T == KHKW");
      if (!pp.strided) {                   // contiguous slab: s_pack[e] = cvt(src[e])   (host checked total % 4, alignment)
        const float4* src4 = reinterpret_cast<const float4*>(src_row);
        for (int q = threadIdx.x; q < (total >> 2); q += blockDim.x) {
          const float4 v = __ldg(src4 + q);
          uint2 pk;
          pk.x = static_cast<uint32_t>(pack_cvt(v.x * mul, pp.fmt)) | (static_cast<uint32_t>(pack_cvt(v.y * mul, pp.fmt)) << 16);
          pk.y = static_cast<uint32_t>(pack_cvt(v.z * mul, pp.fmt)) | (static_cast<uint32_t>(pack_cvt(v.w * mul, pp.fmt)) << 16);
          *reinterpret_cast<uint2*>(s_pack + 4 * q) = pk;
        }
      } else {
        for (int e0 = threadIdx.x; e0 < total; e0 += 4 * blockDim.x) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e0 + j * blockDim.x;
            v[j] = 0.f;
            if (e < total) {
              const int c = e / khkw, t = e - c * khkw;
              v[j] = __ldg(src_row + c * c_stride + t);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e0 + j * blockDim.x;
            if (e < total) s_pack[e] = pack_cvt(v[j] * mul, pp.fmt);     // c * T + t == e
          }
        }
      }
    } else {
      for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int c = e / khkw, t = e - c * khkw;
        s_pack[c * T + t] = pack_cvt(src_row[c * c_stride + t] * mul, pp.fmt);
      }
    }
  }
  __syncthreads();
  if constexpr (VEC) {
    const int octs = pp.cin_pad >> 3;                                    // 8 channels = one 16-byte store
    uint4* orow4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * pp.ktot);
    for (int item = threadIdx.x; item < pp.ntaps * octs; item += blockDim.x) {
      const int tap = item / octs, q = item - tap * octs;
      int ky = pp.ky[tap], kx = pp.kx[tap];
      if (pp.flip) { ky = pp.kh - 1 - ky; kx = pp.kw - 1 - kx; }
      const int t = ky * pp.kw + kx;
      uint32_t wd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 8 * q + 2 * j;
        const uint32_t lo = (real && c < pp.cin) ? s_pack[c * T + t] : 0u;
        const uint32_t hi = (real && c + 1 < pp.cin) ? s_pack[(c + 1) * T + t] : 0u;
        wd[j] = lo | (hi << 16);
      }
      orow4[item] = make_uint4(wd[0], wd[1], wd[2], wd[3]);              // item == tap * octs + q
    }
  } else {
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r) * pp.ktot);   // cin_pad is even: 2 channels / thread
    const int half = pp.cin_pad >> 1;
    for (int tap = 0; tap < pp.ntaps; ++tap) {
      int ky = pp.ky[tap], kx = pp.kx[tap];
      if (pp.flip) { ky = pp.kh - 1 - ky; kx = pp.kw - 1 - kx; }
      const int t = ky * pp.kw + kx;
      for (int c2 = threadIdx.x; c2 < half; c2 += blockDim.x) {
        const int c = 2 * c2;
        const uint32_t lo = (real && c < pp.cin) ? s_pack[c * T + t] : 0u;
        const uint32_t hi = (real && c + 1 < pp.cin) ? s_pack[(c + 1) * T + t] : 0u;
        orow[tap * half + c2] = lo | (hi << 16);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host-side planning
// ------------------------------------------------------------------------------------------------
struct Phase {
  int ntaps;
  int8_t dh[kMaxTaps], dw[kMaxTaps], ky[kMaxTaps], kx[kMaxTaps];
  int grid_h, grid_w;
  int osh, osw, ooh, oow;
  int sh, sw;
  int ktot;
  size_t w_offset;  // element offset into the packed weight buffer
};

struct Plan {
  int nphases;
  Phase ph[4];
  int out_h, out_w;
  int block_n, n_tiles, rows;   // rows = n_tiles * block_n (packed weight rows)
  int c_chunks;
  int tapn;                     // tap-in-N mode
  int kblk_per_tap;
  size_t packed_elems;
  double flops;
};

static int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

static int make_plan(const hfc_conv_desc* d, Plan* pl, bool widenorm = false) {
  if (!d) return set_error(HFC_ERR_INVALID, "conv: null descriptor");
  const hfc_act_geom& in = d->in;
  if (in.n <= 0 || in.h <= 0 || in.w <= 0 || in.c <= 0 || d->cout <= 0)
    return set_error(HFC_ERR_INVALID, "conv: non-positive dimension");
  if (d->kh <= 0 || d->kw <= 0 || d->kh * d->kw > kMaxTaps)
    return set_error(HFC_ERR_INVALID, "conv: unsupported filter size %dx%d", d->kh, d->kw);
  if (d->stride != 1 && d->stride != 2)
    return set_error(HFC_ERR_INVALID, "conv: stride must be 1 or 2");
  if ((d->a_bf16 != 0) != (d->b_bf16 != 0))
    return set_error(HFC_ERR_UNSUPPORTED, "conv: tcgen05 kind::f16 faults on mixed fp16 x bf16 operands (a_bf16 != b_bf16)");
  if (d->precision != HFC_PREC_F16)
    return set_error(HFC_ERR_UNSUPPORTED, "conv: precision mode %d not built", d->precision);
  if (d->window) {
    if (d->transposed || in.cpad != 8 || d->kw > 8 || in.c > 8 || d->stride != 1)
      return set_error(HFC_ERR_INVALID, "conv: window packing needs cpad==8, kw<=8, stride 1");
  } else if (in.cpad % 64 != 0) {
    return set_error(HFC_ERR_INVALID, "conv: input cpad (%d) must be a multiple of 64", in.cpad);
  }
  if (in.c > in.cpad) return set_error(HFC_ERR_INVALID, "conv: in.c > in.cpad");

  memset(pl, 0, sizeof(*pl));
  const int s = d->stride;
  if (!d->transposed) {
    pl->out_h = (in.h + d->pad_t + d->pad_b - d->kh) / s + 1;
    pl->out_w = (in.w + d->pad_l + d->pad_r - d->kw) / s + 1;
    if (d->pad_mode == HFC_PAD_REFLECT &&
        (in.pt < d->pad_t || in.pl < d->pad_l || in.pb < d->pad_b || in.pr < d->pad_r))
      return set_error(HFC_ERR_INVALID, "conv: reflect padding needs a materialised input border");
    if (d->pad_mode == HFC_PAD_ZERO && (in.pt | in.pl | in.pb | in.pr))
      return set_error(HFC_ERR_INVALID, "conv: zero padding expects a border-less input buffer");
    pl->nphases = 1;
    Phase& p = pl->ph[0];
    p.grid_h = pl->out_h; p.grid_w = pl->out_w;
    p.osh = p.osw = 1; p.ooh = p.oow = 0;
    p.sh = p.sw = s;
    static const bool env_no_tapn = getenv("HFC_NO_TAPN") != nullptr;
    pl->tapn = (!d->window && s == 1 && d->out_mode == HFC_OUT_NCHW_F32 && !d->norm && d->kw > 1 &&
                d->kw * d->cout <= 32 && pl->out_w >= 64 && d->block_n == 0 &&
                (d->wide == 3 || (d->wide == 0 && !env_no_tapn))) ? 1 : 0;
    if (d->window || pl->tapn) {
      for (int ky = 0; ky < d->kh; ++ky) {
        p.dh[p.ntaps] = static_cast<int8_t>(ky - d->pad_t);
        p.dw[p.ntaps] = static_cast<int8_t>(-d->pad_l);
        p.ky[p.ntaps] = static_cast<int8_t>(ky); p.kx[p.ntaps] = 0;
        ++p.ntaps;
      }
    } else {
      for (int ky = 0; ky < d->kh; ++ky)
        for (int kx = 0; kx < d->kw; ++kx) {
          p.dh[p.ntaps] = static_cast<int8_t>(ky - d->pad_t);
          p.dw[p.ntaps] = static_cast<int8_t>(kx - d->pad_l);
          p.ky[p.ntaps] = static_cast<int8_t>(ky); p.kx[p.ntaps] = static_cast<int8_t>(kx);
          ++p.ntaps;
        }
    }
  } else {
    if (in.pt | in.pl | in.pb | in.pr)
      return set_error(HFC_ERR_INVALID, "conv_transpose: expects a border-less input buffer");
    const int op = s - 1;
    pl->out_h = (in.h - 1) * s - 2 * d->pad_t + d->kh + op;
    pl->out_w = (in.w - 1) * s - 2 * d->pad_l + d->kw + op;
    pl->nphases = s * s;
    for (int a = 0; a < s; ++a)
      for (int b = 0; b < s; ++b) {
        Phase& p = pl->ph[a * s + b];
        p.grid_h = (pl->out_h - a + s - 1) / s;
        p.grid_w = (pl->out_w - b + s - 1) / s;
        p.osh = p.osw = s; p.ooh = a; p.oow = b;
        p.sh = p.sw = 1;
        // out[o] += in[i] * w[k] with o = i*s - pad + k  =>  i = (o + pad - k)/s, o = g*s + a
        for (int ky = 0; ky < d->kh; ++ky) {
          if (((a + d->pad_t - ky) % s) != 0) continue;
          for (int kx = 0; kx < d->kw; ++kx) {
            if (((b + d->pad_l - kx) % s) != 0) continue;
            p.dh[p.ntaps] = static_cast<int8_t>((a + d->pad_t - ky) / s);
            p.dw[p.ntaps] = static_cast<int8_t>((b + d->pad_l - kx) / s);
            p.ky[p.ntaps] = static_cast<int8_t>(ky); p.kx[p.ntaps] = static_cast<int8_t>(kx);
            ++p.ntaps;
          }
        }
      }
  }
  if (pl->out_h <= 0 || pl->out_w <= 0)
    return set_error(HFC_ERR_INVALID, "conv: empty output (%d x %d)", pl->out_h, pl->out_w);

  // N tiling
  int bn = d->block_n;
  const int n_cols = pl->tapn ? d->kw * d->cout : d->cout;   // GEMM columns
  if (pl->tapn) bn = (n_cols + 15) / 16 * 16;
  if (bn == 0) {
    const int nt = (d->cout + 255) / 256;
    bn = ((d->cout + nt - 1) / nt + 15) / 16 * 16;
  }
  if (bn % 16 != 0 || bn < 16 || bn > 256)
    return set_error(HFC_ERR_INVALID, "conv: block_n %d must be a multiple of 16 in [16,256]", bn);
  pl->block_n = bn;
  pl->n_tiles = (n_cols + bn - 1) / bn;
  pl->rows = pl->n_tiles * bn;
  if (widenorm) {
    if (pl->n_tiles != 4 || d->cout % 16 != 0 || !d->norm || d->transposed || d->window || pl->tapn ||
        d->out_mode != HFC_OUT_NHWC_F16)
      return set_error(HFC_ERR_INVALID, "conv: the wide fused ChannelNorm needs a conv2d with exactly 4 N tiles "
                                        "(768 < cout <= 1024, cout %% 16 == 0), norm = 1 and an NHWC fp16 output geometry");
  } else if (d->norm && pl->n_tiles != 1)
    return set_error(HFC_ERR_INVALID,
                     "conv: fused ChannelNorm needs cout <= 256 (use NHWC_F32 + hfc_channelnorm)");
  if (d->norm && d->cout < 2) return set_error(HFC_ERR_INVALID, "conv: ChannelNorm needs cout >= 2");

  pl->c_chunks = d->window ? 1 : in.cpad / kBlockK;
  size_t off = 0;
  for (int i = 0; i < pl->nphases; ++i) {
    Phase& p = pl->ph[i];
    p.ktot = p.ntaps * pl->c_chunks * kBlockK;
    p.w_offset = off;
    off += static_cast<size_t>(pl->rows) * p.ktot;
  }
  pl->packed_elems = off;
  pl->flops = 2.0 * in.n * (d->transposed ? in.h * in.w : pl->out_h * pl->out_w) * (double)d->cout *
              in.c * d->kh * d->kw;

  // output checks
  const hfc_act_geom& o = d->out;
  if (o.n != in.n || o.h != pl->out_h || o.w != pl->out_w)
    return set_error(HFC_ERR_INVALID, "conv: out geometry (%d,%d,%d) != expected (%d,%d,%d)", o.n,
                     o.h, o.w, in.n, pl->out_h, pl->out_w);
  if (d->out_mode == HFC_OUT_NHWC_F16) {
    if (o.cpad % 8 != 0 || o.cpad < d->cout)
      return set_error(HFC_ERR_INVALID, "conv: out.cpad must be a multiple of 8 and >= cout");
    if (d->out_reflect && (o.pt >= o.h || o.pb >= o.h || o.pl >= o.w || o.pr >= o.w))
      return set_error(HFC_ERR_INVALID, "conv: reflected border wider than the image");
  } else if (d->out_mode == HFC_OUT_NHWC_F32) {
    if (o.cpad % 4 != 0 || o.cpad < d->cout)
      return set_error(HFC_ERR_INVALID, "conv: fp32 row pitch must be a multiple of 4 and >= cout");
  } else if (d->out_mode != HFC_OUT_NCHW_F32) {
    return set_error(HFC_ERR_INVALID, "conv: bad out_mode");
  }
  return HFC_OK;
}

static int tile_and_stages(const hfc_conv_desc* d, const Plan& pl, const Phase& ph, int batch,
                           ConvKernelParams* kp, bool widenorm = false) {
  kp->tw = std::min(16, next_pow2(ph.grid_w));
  kp->th = std::min(kBlockM / kp->tw, next_pow2(ph.grid_h));
  kp->tn = kBlockM / (kp->tw * kp->th);
  // Grids that the power-of-two tile covers badly (e.g. the 18x18 padded domain of a 3x3 data gradient: 6 tiles of
  // 16x8 per image, most of them nearly empty) get the free-form tile with the fewest tiles: tw*th*tn <= 128, the
  // TMA box simply delivers fewer rows and the idle accumulator rows are masked in the epilogue.
  kp->tx_short = 0;
  bool free_tile = false;
  {
    auto count = [&](int tw, int th, int tn) {
      return static_cast<long long>((ph.grid_w + tw - 1) / tw) * ((ph.grid_h + th - 1) / th) * ((batch + tn - 1) / tn);
    };
    const long long legacy = count(kp->tw, kp->th, kp->tn);
    long long best = legacy;
    int bw_ = kp->tw, bh_ = kp->th, bn_ = kp->tn;
    static const bool env_no_free = getenv("HFC_NO_FREE_TILE") != nullptr;
    for (int tw = std::min(ph.grid_w, kBlockM); tw >= 8 && !env_no_free; --tw) {
      if (tw * ph.sw > 256) continue;
      for (int th = std::min(ph.grid_h, kBlockM / tw); th >= 1; --th) {
        if (th * ph.sh > 256) continue;
        const int tn = std::max(1, std::min(batch, kBlockM / (tw * th)));
        const long long c = count(tw, th, tn);
        // accept only a clear win (>= 20 % fewer tiles): the power-of-two tiles have the better store pattern
        if (c * 5 <= legacy * 4 && c < best) { best = c; bw_ = tw; bh_ = th; bn_ = tn; }
      }
    }
    if (best < legacy) {
      kp->tw = bw_; kp->th = bh_; kp->tn = bn_;
      kp->tx_short = (kBlockM - bw_ * bh_ * bn_) * kBlockK * 2;
      free_tile = true;
    }
  }
  // 'wide' mode (few output channels, big maps, e.g. the 7x7 60->3 head): a tile is 128 consecutive
  // pixels of one row; per filter ROW one halo row of 128+kw-1 pixels is fetched and the kw filter
  // columns are served from it by shifting the UMMA descriptor; all weights stay resident in smem.
  // L2->SM traffic drops from kh*kw*16 KB to kh*(128+kw-1)*128 B per tile.
  const int w_bytes_total = ph.ntaps * pl.c_chunks * pl.block_n * kBlockK * 2;
  static const bool env_no_wide = getenv("HFC_NO_WIDE") != nullptr;       // debugging switches
  static const bool env_no_cluster = getenv("HFC_NO_CLUSTER") != nullptr;
  kp->wide = (d->wide != 2) && !(env_no_wide && d->wide == 0) && !d->transposed && !d->window && d->stride == 1 && d->kw > 1 && pl.c_chunks == 1 &&
             pl.n_tiles == 1 && ph.grid_w >= 64 && w_bytes_total <= 112 * 1024 && (d->wide == 1 || pl.block_n <= 32);
  kp->kw = d->kw;
  static const char* env_dbg = getenv("HFC_DBG");
  kp->dbg = env_dbg ? atoi(env_dbg) : 0;
  static const bool env_wide_boff = getenv("HFC_WIDE_BASEOFF") != nullptr;
  kp->wide_boff = env_wide_boff ? 1 : 0;
  kp->a_region = 0;
  kp->tapn = pl.tapn;
  kp->w_step = kBlockM - d->kw + 1;
  if (pl.tapn) kp->wide = 0;
  // window packing on wide images: 128 consecutive pixels per tile and a plain (128 + 7)-pixel segment per filter
  // row (2 160 B) instead of 128 inflated windows (16 KB)
  static const bool env_no_winflat = getenv("HFC_NO_WINFLAT") != nullptr;
  kp->winflat = (d->window && !env_no_winflat && ph.grid_w >= 96) ? 1 : 0;
  if (kp->wide || pl.tapn || kp->winflat) {
    kp->tw = kBlockM; kp->th = 1; kp->tn = 1;
    kp->tx_short = kp->winflat ? kABytes - (kBlockM + 7) * 16 : 0;
    free_tile = false;
  }
  // Thin instantiation (conv_igemm_kernel<false, 1, false, true>): one N tile of <= 64 columns (tap-in-N: <= 32) on a
  // map big enough that the launch is a long stream of tiny tiles -- the epilogue-bound layers E1, G.up4, G3.  Single
  // CTAs (no pair / cluster: the K loops are short), 4 accumulator stages.  Default level 2: one-N-tile layers of up to
  // 128 columns as well (E2 164 -> 130 us, G.up3 190 -> 165 us); HFC_THIN_EPILOGUE=1 restricts it to <= 64 columns, 0
  // disables it.
  static const int env_thin = getenv("HFC_THIN_EPILOGUE") ? atoi(getenv("HFC_THIN_EPILOGUE")) : 2;
  {
    // the choice depends on the map size of ONE image only, never on the batch: a sample's result must not depend on
    // its batch neighbours bit for bit (tests/test_gpu_parity.py), and the two epilogues sum the statistics in a
    // different order
    const int n_max = env_thin >= 2 ? 128 : 64;
    kp->thin = (env_thin >= 1 && !widenorm && !kp->wide && pl.n_tiles == 1 && ph.grid_h * ph.grid_w >= 4096 &&
                (pl.tapn ? (d->out_mode == HFC_OUT_NCHW_F32 && pl.block_n <= 32)
                         : (d->out_mode == HFC_OUT_NHWC_F16 && pl.block_n <= n_max && d->norm != 2))) ? 1 : 0;
  }
  kp->tiles_w = pl.tapn ? (ph.grid_w + (kBlockM - d->kw + 1) - 1) / (kBlockM - d->kw + 1)
                        : (ph.grid_w + kp->tw - 1) / kp->tw;
  kp->tiles_h = (ph.grid_h + kp->th - 1) / kp->th;
  kp->tiles_n = (batch + kp->tn - 1) / kp->tn;
  int stage_bytes = kABytes + pl.block_n * kBlockK * 2;
  int budget = 226 * 1024 - 1024 - kTailBytes - (pl.tapn ? (kp->thin ? 2 : 1) * kTapnBytes : 0);
  {
    // thin layers: the whole weight matrix of the N tile (num_kb x block_n x 128 B) stays in smem when it fits in 72 KB
    static const bool env_no_bres = getenv("HFC_NO_BRES") != nullptr;
    const int w_all = ph.ntaps * pl.c_chunks * pl.block_n * kBlockK * 2;
    kp->b_res = (kp->thin && !env_no_bres && w_all <= 72 * 1024) ? 1 : 0;
    if (kp->b_res) {
      stage_bytes = kABytes;
      budget -= w_all;
    }
  }
  if (kp->wide) {
    kp->a_region = ((kBlockM + d->kw - 1) * kBlockK * 2 + 1023) / 1024 * 1024;
    stage_bytes = kp->a_region;
    budget -= w_bytes_total;
  }
  int stages = budget / stage_bytes;
  stages = std::max(2, std::min(stages, kMaxStages));
  kp->stages = stages;
  // Cluster shape (TMA multicast of the shared operand tiles).  Auto: only for launches with enough tiles
  // to fill the machine, along N when the N tiles pair up, along M when the M tiles pair up.
  const int tiles_m = kp->tiles_w * kp->tiles_h * kp->tiles_n;
  int cm = d->cluster_m, cn = d->cluster_n;
  if (env_no_cluster && cm == 0) cm = 1;
  if (env_no_cluster && cn == 0) cn = 1;
  if (cm == 0 || cn == 0) {
    const bool big = static_cast<long long>(tiles_m) * pl.n_tiles >= 128;
    cn = (cn == 0) ? ((big && pl.n_tiles % 2 == 0) ? 2 : 1) : cn;
    cm = (cm == 0) ? ((big && tiles_m % 2 == 0) ? 2 : 1) : cm;
  }
  if (cm < 1 || cm > 2 || cn < 1 || cn > 2) return -1;
  if (kp->wide || pl.tapn || kp->winflat || kp->thin) cm = cn = 1;
  if (free_tile) cn = 1;                 // the multicast A slices assume a full 128-row tile
  if ((pl.block_n / cm) % 8 != 0 || pl.block_n % cm != 0) cm = 1;
  kp->a_split_n = 0;
  if (cn > 1) {
    if (kp->tn % cn == 0) kp->a_split_n = 1;
    else if (kp->th % cn != 0) cn = 1;
  }
  kp->cm = cm;
  kp->cn = cn;
  // CTA pairs (cta_group::2): whenever two M tiles share a cluster and the K loop is long enough for the
  // operand bandwidth to matter.  Each CTA then holds only half of the B tile -> deeper smem ring.
  static const bool env_no_pair = getenv("HFC_NO_PAIR") != nullptr;
  kp->pair = (cm == 2 && !kp->wide && !env_no_pair && d->pair != 2 && (pl.block_n / 2) % 8 == 0 &&
              (d->pair == 1 || ph.ntaps * pl.c_chunks >= 8)) ? 1 : 0;
  kp->b_rows = kp->pair ? pl.block_n / 2 : pl.block_n;
  kp->nsub = 1;
  if (kp->pair) {
    // Two N tiles per work item (one in each 256-column TMEM half) share ONE fetch of the A tile: the pair kernel is
    // bound by the L2 -> SM operand traffic (~85 % of the 6.3 KB/clk LTS cap in ncu), which this cuts from
    // 16 + 15 KB to (16 + 30) / 2 = 23 KB per tile and K block.  Costs the accumulator double buffering, so it is taken
    // only when the wave count says it wins: waves x (1 | 1.8; measured: one wave of 64 double items 100.6 us vs two
    // waves of single items 111.6 us on the 960x960 layer).
    static const bool env_no_nsub = getenv("HFC_NO_NSUB") != nullptr;
    const int pairs = 74;
    const long long items1 = static_cast<long long>((tiles_m + 1) / 2) * pl.n_tiles;
    const long long items2 = static_cast<long long>((tiles_m + 1) / 2) * (pl.n_tiles / 2);
    if (!env_no_nsub && (cn == 1 || d->cluster_n == 0) && !d->norm && pl.n_tiles % 2 == 0 && ph.ntaps * pl.c_chunks >= 16 &&
        ((items2 + pairs - 1) / pairs) * 180 < ((items1 + pairs - 1) / pairs) * 100) {
      kp->nsub = 2;
      kp->cn = cn = 1;          // an auto-chosen N multicast gives way: it does not reduce the LTS traffic (csz <= 4)
      kp->a_split_n = 0;
    }
    stage_bytes = kABytes + kp->nsub * kp->b_rows * kBlockK * 2;
    kp->stages = std::max(2, std::min(budget / stage_bytes, kMaxStages));
  }
  if (widenorm) {
    // the channel row of a pixel = 4 N tiles = 2 TMEM halves in each of the two pairs of a 2 x 2 cluster
    if (free_tile || kp->wide || kp->winflat || tiles_m % 2 != 0 || (pl.block_n / 2) % 8 != 0)
      return -2;
    kp->cm = 2; kp->cn = 2; kp->pair = 1; kp->nsub = 2;
    kp->a_split_n = 0;
    if (kp->tn % 2 == 0) kp->a_split_n = 1;
    else if (kp->th % 2 != 0) return -2;
    kp->b_rows = pl.block_n / 2;
    stage_bytes = kABytes + 2 * kp->b_rows * kBlockK * 2;
    kp->stages = std::max(2, std::min(budget / stage_bytes, kMaxStages));
  }
  return stage_bytes;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace hfc

using namespace hfc;

extern "C" int hfc_conv_query(const hfc_conv_desc* d, hfc_conv_info* info) {
  Plan pl;
  int rc = make_plan(d, &pl);
  if (rc != HFC_OK) return rc;
  if (!info) return set_error(HFC_ERR_INVALID, "conv_query: null info");
  ConvKernelParams kp;
  if (tile_and_stages(d, pl, pl.ph[0], d->in.n, &kp) < 0)
    return set_error(HFC_ERR_INVALID, "conv: cluster_m / cluster_n must be 0 (auto), 1 or 2");
  info->packed_weight_bytes = pl.packed_elems * 2;
  info->out_h = pl.out_h;
  info->out_w = pl.out_w;
  info->phases = pl.nphases;
  info->block_n = pl.block_n;
  info->n_tiles = pl.n_tiles;
  info->m_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  info->stages = kp.stages;
  info->wide = kp.wide;
  info->pair = kp.pair;
  info->tapn = kp.tapn;
  info->nsub = kp.nsub;
  info->cluster_m = kp.cm;
  info->cluster_n = kp.cn;
  info->k_total = 0;
  for (int i = 0; i < pl.nphases; ++i) info->k_total += pl.ph[i].ktot;
  info->flops = pl.flops;
  return HFC_OK;
}

static int pack_weights_impl(const hfc_conv_desc* d, const float* w, const float* scale, void* packed,
                             void* stream);

extern "C" int hfc_conv_pack_weights(const hfc_conv_desc* d, const float* w, void* packed,
                                     void* stream) {
  return pack_weights_impl(d, w, nullptr, packed, stream);
}

extern "C" int hfc_conv_pack_weights_scaled(const hfc_conv_desc* d, const float* w, const float* scale,
                                            void* packed, void* stream) {
  if (!scale) return set_error(HFC_ERR_INVALID, "conv_pack_weights_scaled: null scale");
  return pack_weights_impl(d, w, scale, packed, stream);
}

static int pack_weights_impl(const hfc_conv_desc* d, const float* w, const float* scale, void* packed,
                             void* stream) {
  Plan pl;
  int rc = make_plan(d, &pl);
  if (rc != HFC_OK) return rc;
  if (!w || !packed) return set_error(HFC_ERR_INVALID, "conv_pack_weights: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < pl.nphases; ++i) {
    const Phase& ph = pl.ph[i];
    PackParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.cout = d->cout; pp.cin = d->in.c; pp.kh = d->kh; pp.kw = d->kw;
    pp.cin_pad = d->in.cpad; pp.ntaps = ph.ntaps; pp.ktot = ph.ktot; pp.rows = pl.rows;
    pp.transposed = d->transposed; pp.window = d->window; pp.fmt = d->b_bf16 ? 1u : 0u;
    pp.dgrad = d->dgrad;
    pp.tapn = pl.tapn; pp.cout_real = d->cout;
    if (pl.tapn) pp.cout = d->kw * d->cout;   // GEMM rows = (filter column, output channel)
    memcpy(pp.ky, ph.ky, sizeof(pp.ky));
    memcpy(pp.kx, ph.kx, sizeof(pp.kx));
    const size_t row_smem = static_cast<size_t>(d->in.c) * ((d->kh * d->kw) | 1) * 2;
    if (!d->window && !pl.tapn && row_smem <= 48 * 1024) {
      PackRowParams pr;
      memset(&pr, 0, sizeof(pr));
      pr.rows_real = d->cout; pr.cin = d->in.c; pr.kh = d->kh; pr.kw = d->kw; pr.cin_pad = d->in.cpad;
      pr.ntaps = ph.ntaps; pr.ktot = ph.ktot;
      pr.strided = (d->dgrad || d->transposed) ? 1 : 0;
      pr.flip = d->dgrad ? 1 : 0;
      pr.fmt = d->b_bf16 ? 1u : 0u;
      memcpy(pr.ky, ph.ky, sizeof(pr.ky));
      memcpy(pr.kx, ph.kx, sizeof(pr.kx));
      uint16_t* dst = reinterpret_cast<uint16_t*>(packed) + ph.w_offset;
      static const bool vec_on = [] { const char* e = getenv("HFC_PACK_VEC"); return !(e && e[0] == '0'); }();
      const int khkw = d->kh * d->kw;
      const bool vec = vec_on && (khkw & 1) && d->in.cpad % 8 == 0 && ph.ktot % 8 == 0 && ph.w_offset % 8 == 0 &&
                       reinterpret_cast<uintptr_t>(packed) % 16 == 0 &&
                       (pr.strided || ((static_cast<long long>(d->in.c) * khkw) % 4 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0));
      if (vec) {
        switch (khkw) {
          case 1: pack_rows_kernel<1, true><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 9: pack_rows_kernel<9, true><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 25: pack_rows_kernel<25, true><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 49: pack_rows_kernel<49, true><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          default: pack_rows_kernel<0, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
        }
      } else {
        switch (khkw) {
          case 1: pack_rows_kernel<1, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 9: pack_rows_kernel<9, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 16: pack_rows_kernel<16, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 25: pack_rows_kernel<25, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          case 49: pack_rows_kernel<49, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
          default: pack_rows_kernel<0, false><<<pl.rows, 256, row_smem, st>>>(w, scale, dst, pr); break;
        }
      }
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "pack_rows launch: %s", cudaGetErrorString(e));
      continue;
    }
    const size_t total = static_cast<size_t>(pl.rows) * ph.ktot;
    const int threads = 256;
    const int blocks = static_cast<int>(std::min<size_t>((total + threads - 1) / threads, 148 * 16));
    pack_weights_kernel<<<blocks, threads, 0, st>>>(
        w, scale, reinterpret_cast<uint16_t*>(packed) + ph.w_offset, pp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return set_error(HFC_ERR_LAUNCH, "pack_weights launch: %s", cudaGetErrorString(e));
  }
  return HFC_OK;
}

struct WideNormArgs {
  const float* res1;
  const float* res2;
  int32_t ld_res;
  float* out_f32;
  int32_t ld_f32;
};

// The s*s sub-pixel phases of a transposed convolution write disjoint output pixels and read the same input: they are
// independent launches.  On one stream they run back to back, each with its own ramp-up and tail (on the 4x4 ... 16x16
// maps of the hyper-synthesis networks a phase is a dozen CTAs); forked onto helper streams they overlap and share the
// input tiles in L2.  Fork / join is two event edges per helper stream, valid in eager mode and under stream capture
// (the helper streams join the capture through the event and are joined back before the call returns).
struct PhaseFork {
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
  int device = -1;
  bool ok = false;
};

static PhaseFork* phase_fork(cudaStream_t st) {
  static thread_local PhaseFork pools[16];     // per calling thread: the entry points stay thread-safe for distinct streams
  static const bool enabled = [] {
    const char* e = getenv("HFC_PHASE_STREAMS");
    return !(e && e[0] == '0');
  }();
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  PhaseFork& pf = pools[dev];
  if (!pf.ok) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;                      // never create streams / events inside a capture: this call stays serial
    }
    bool good = cudaEventCreateWithFlags(&pf.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && good; ++i)
      good = cudaStreamCreateWithFlags(&pf.aux[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&pf.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!good) {
      cudaGetLastError();
      return nullptr;
    }
    pf.device = dev;
    pf.ok = true;
  }
  return &pf;
}

static int conv_forward_impl(const hfc_conv_desc* d, const void* in, const void* packed, const float* bias,
                             const float* gamma, const float* beta, void* out, void* stream, const WideNormArgs* wn) {
  Plan pl;
  int rc = make_plan(d, &pl, wn != nullptr);
  if (rc != HFC_OK) return rc;
  if (!in || !packed || (!out && !(wn && wn->out_f32))) return set_error(HFC_ERR_INVALID, "conv_forward: null pointer");
  if (d->norm && (!gamma || !beta))
    return set_error(HFC_ERR_INVALID, "conv_forward: fused norm needs gamma and beta");
  int sm_count = 0;
  rc = device_sm_count(&sm_count);
  if (rc != HFC_OK) return rc;
  PFN_encodeTiled encode = get_encode_fn();
  if (!encode) return set_error(HFC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not found");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const hfc_act_geom& ig = d->in;
  const int Hp = ig.h + ig.pt + ig.pb;
  const int Wp = ig.w + ig.pl + ig.pr;

  PhaseFork* pf = (pl.nphases > 1 && pl.nphases <= 4) ? phase_fork(st) : nullptr;
  if (pf) {
    cudaError_t e = cudaEventRecord(pf->fork, st);
    for (int i = 1; i < pl.nphases && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(pf->aux[i - 1], pf->fork, 0);
    if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "conv: phase fork: %s", cudaGetErrorString(e));
  }
  const cudaStream_t st_main = st;
  auto launch_phase = [&](int i) -> int {
    st = (pf && i > 0) ? pf->aux[i - 1] : st_main;
    const Phase& ph = pl.ph[i];
    ConvKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    const int stage_bytes = tile_and_stages(d, pl, ph, ig.n, &kp, wn != nullptr);
    if (stage_bytes == -2)
      return set_error(HFC_ERR_UNSUPPORTED, "conv: the wide fused ChannelNorm needs power-of-two pixel tiles that pair up "
                                            "along M and split in two (this geometry does not)");
    if (stage_bytes < 0)
      return set_error(HFC_ERR_INVALID, "conv: cluster_m / cluster_n must be 0 (auto), 1 or 2");
    kp.n_tiles = pl.n_tiles;
    kp.block_n = pl.block_n;
    kp.c_chunks = pl.c_chunks;
    kp.ntaps = ph.ntaps;
    kp.num_kb = kp.wide ? d->kh : ph.ntaps * pl.c_chunks;
    if (kp.num_kb > kMaxKb)
      return set_error(HFC_ERR_UNSUPPORTED, "conv: %d K blocks per tile (taps x 64-channel chunks) exceed the %d the producer's table holds",
                       kp.num_kb, kMaxKb);
    kp.grid_h = ph.grid_h; kp.grid_w = ph.grid_w; kp.batch = ig.n;
    kp.sh = ph.sh; kp.sw = ph.sw;
    kp.ih0 = ig.pt; kp.iw0 = ig.pl;
    kp.out_mode = d->out_mode;
    kp.osh = ph.osh; kp.osw = ph.osw; kp.ooh = ph.ooh; kp.oow = ph.oow;
    kp.out_h = pl.out_h; kp.out_w = pl.out_w;
    kp.out_cpad = (d->out_mode == HFC_OUT_NCHW_F32) ? d->cout : d->out.cpad;
    kp.out_pt = d->out.pt; kp.out_pl = d->out.pl; kp.out_pb = d->out.pb; kp.out_pr = d->out.pr;
    if (d->out_mode != HFC_OUT_NHWC_F16) kp.out_pt = kp.out_pl = kp.out_pb = kp.out_pr = 0;
    kp.out_reflect = d->out_reflect;
    kp.cout = d->cout;
    kp.act = d->act; kp.norm = d->norm; kp.eps = d->eps;
    kp.k_splits = 1; kp.kb_per_split = kp.num_kb; kp.gemm = 0; kp.out_atomic = 0;
    kp.fmt_a = d->a_bf16 ? 1u : 0u;
    kp.fmt_b = d->b_bf16 ? 1u : 0u;
    kp.bias = bias; kp.gamma = gamma; kp.beta = beta;
    kp.out = out;
    if (wn) {
      kp.norm = 2;
      kp.res1 = wn->res1; kp.res2 = wn->res2; kp.ld_res = wn->ld_res;
      kp.out_f32 = wn->out_f32; kp.ld_f32 = wn->ld_f32;
    }
    memcpy(kp.tap_dh, ph.dh, sizeof(kp.tap_dh));
    memcpy(kp.tap_dw, ph.dw, sizeof(kp.tap_dw));
    if (kp.wide) {  // one entry per filter row: (ky - pad_t, -pad_l)
      for (int ky = 0; ky < d->kh; ++ky) {
        kp.tap_dh[ky] = static_cast<int8_t>(ky - d->pad_t);
        kp.tap_dw[ky] = static_cast<int8_t>(-d->pad_l);
      }
    }

    // A: 4-D map over the NHWC buffer {channels, W, H, N}
    CUtensorMap tmA, tmB;
    {
      cuuint64_t dims[4], strides[3];
      cuuint32_t box[4], estr[4];
      if (d->window && !kp.winflat) {
        // dim0 = 64 elements = 8 consecutive pixels x 8 channels, dim1 = window start (pixel)
        if (Wp < 8) return set_error(HFC_ERR_INVALID, "conv: window packing needs Wp >= 8");
        dims[0] = 64; dims[1] = static_cast<cuuint64_t>(Wp - 7);
      } else {
        dims[0] = static_cast<cuuint64_t>(ig.cpad); dims[1] = static_cast<cuuint64_t>(Wp);
      }
      dims[2] = static_cast<cuuint64_t>(Hp); dims[3] = static_cast<cuuint64_t>(ig.n);
      strides[0] = static_cast<cuuint64_t>(ig.cpad) * 2;
      strides[1] = static_cast<cuuint64_t>(Wp) * ig.cpad * 2;
      strides[2] = static_cast<cuuint64_t>(Hp) * Wp * ig.cpad * 2;
      // per-CTA slice of the A tile when it is shared across a cluster row (multicast)
      const int th_box = kp.a_split_n ? kp.th : kp.th / kp.cn;
      const int tn_box = kp.a_split_n ? kp.tn / kp.cn : kp.tn;
      box[0] = kBlockK; box[1] = kp.tw * ph.sw; box[2] = th_box * ph.sh; box[3] = tn_box;
      if (kp.wide) { box[1] = kBlockM + d->kw - 1; box[2] = 1; box[3] = 1; }
      if (kp.winflat) { box[0] = 8; box[1] = kBlockM + 7; box[2] = 1; box[3] = 1; }   // plain segment, 16 B per pixel
      estr[0] = 1; estr[1] = ph.sw; estr[2] = ph.sh; estr[3] = 1;
      CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(in), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          kp.winflat ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS)
        return set_error(HFC_ERR_LAUNCH, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    }
    {
      cuuint64_t dims[2] = {static_cast<cuuint64_t>(ph.ktot), static_cast<cuuint64_t>(pl.rows)};
      cuuint64_t strides[1] = {static_cast<cuuint64_t>(ph.ktot) * 2};
      cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(pl.block_n / kp.cm)};
      cuuint32_t estr[2] = {1, 1};
      void* wptr = const_cast<uint16_t*>(reinterpret_cast<const uint16_t*>(packed) + ph.w_offset);
      CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wptr, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS)
        return set_error(HFC_ERR_LAUNCH, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }

    const int tiles_m = kp.tiles_w * kp.tiles_h * kp.tiles_n;
    const int csize = kp.cm * kp.cn;
    const int ctiles = ((tiles_m + kp.cm - 1) / kp.cm) * ((kp.n_tiles + kp.cn * kp.nsub - 1) / (kp.cn * kp.nsub));
    // clusters of 2 can use all 148 SMs; clusters of 4 only 132 (GPCs of 16/18/20 SMs)
    const int max_clusters = csize == 4 ? (sm_count * 132 / 148) / 4 : sm_count / csize;
    const int grid = std::min(ctiles, std::max(1, max_clusters)) * csize;
    const size_t smem = static_cast<size_t>(kp.stages) * stage_bytes + 1024 /*align*/ + kTailBytes +
                        (kp.tapn ? (kp.thin ? 2 : 1) * kTapnBytes : 0) +
                        (kp.b_res ? static_cast<size_t>(kp.num_kb) * pl.block_n * kBlockK * 2 : 0) + (kp.wide ? static_cast<size_t>(kp.num_kb) * kp.kw * pl.block_n * kBlockK * 2 : 0);
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<false, 1>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv_igemm_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv_igemm_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv_igemm_kernel<true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv_igemm_kernel<false, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024);
      if (e != cudaSuccess)
        return set_error(HFC_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kp.thin ? kThinThreads : kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const bool thin = kp.thin && !wn && !kp.pair && kp.cn == 1 && kp.cm == 1 && kp.k_splits == 1;
    if (kp.thin && !thin) return set_error(HFC_ERR_INVALID, "conv: internal: thin plan with a pair / cluster launch");
    cudaError_t e = thin ? cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false, 1, false, true>, tmA, tmB, kp)
                    : wn ? cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, 2, true>, tmA, tmB, kp)
                    : !kp.pair ? cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false, 1>, tmA, tmB, kp)
                    : kp.nsub == 2 ? cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, 2>, tmA, tmB, kp)
                                   : cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, 1>, tmA, tmB, kp);
    if (e != cudaSuccess)
      return set_error(HFC_ERR_LAUNCH, "conv_igemm launch: %s", cudaGetErrorString(e));
    note_launch();
    return HFC_OK;
  };
  int rc_phase = HFC_OK;
  for (int i = 0; i < pl.nphases && rc_phase == HFC_OK; ++i) rc_phase = launch_phase(i);
  if (pf) {                                // join (also after an error: a capture must not be left with unjoined streams)
    cudaError_t e = cudaSuccess;
    for (int i = 1; i < pl.nphases && e == cudaSuccess; ++i) {
      e = cudaEventRecord(pf->join[i - 1], pf->aux[i - 1]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st_main, pf->join[i - 1], 0);
    }
    if (e != cudaSuccess && rc_phase == HFC_OK) return set_error(HFC_ERR_LAUNCH, "conv: phase join: %s", cudaGetErrorString(e));
  }
  return rc_phase;
}

extern "C" int hfc_conv_forward(const hfc_conv_desc* d, const void* in, const void* packed,
                                const float* bias, const float* gamma, const float* beta, void* out,
                                void* stream) {
  return conv_forward_impl(d, in, packed, bias, gamma, beta, out, stream, nullptr);
}

extern "C" int hfc_conv_widenorm_supported(const hfc_conv_desc* d) {
  Plan pl;
  int rc = make_plan(d, &pl, true);
  if (rc != HFC_OK) return rc;
  if (d->out.cpad != d->cout)
    return set_error(HFC_ERR_UNSUPPORTED, "conv: the wide fused ChannelNorm writes no channel padding (out.cpad must equal cout)");
  ConvKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  const int sb = tile_and_stages(d, pl, pl.ph[0], d->in.n, &kp, true);
  if (sb < 0) return set_error(HFC_ERR_UNSUPPORTED, "conv: pixel tiles of this geometry do not pair up for the wide fused ChannelNorm");
  return HFC_OK;
}

extern "C" int hfc_conv_forward_widenorm(const hfc_conv_desc* d, const void* in, const void* packed, const float* bias,
                                         const float* gamma, const float* beta, const float* res1, const float* res2,
                                         int32_t ld_res, float* out_f32, int32_t ld_f32, void* out_act, void* stream) {
  if (!d || !gamma || !beta) return set_error(HFC_ERR_INVALID, "conv_forward_widenorm: descriptor, gamma and beta required");
  if (!out_f32 && !out_act) return set_error(HFC_ERR_INVALID, "conv_forward_widenorm: no output requested");
  if ((res1 || res2) && (ld_res < d->cout || ld_res % 4 != 0))
    return set_error(HFC_ERR_INVALID, "conv_forward_widenorm: residual row pitch must be a multiple of 4 and >= cout");
  if (out_f32 && (ld_f32 < d->cout || ld_f32 % 4 != 0))
    return set_error(HFC_ERR_INVALID, "conv_forward_widenorm: fp32 row pitch must be a multiple of 4 and >= cout");
  if (d->out.cpad != d->cout)
    return set_error(HFC_ERR_UNSUPPORTED, "conv_forward_widenorm: out.cpad must equal cout (no channel padding is written)");
  WideNormArgs wn{res1, res2, ld_res, out_f32, ld_f32};
  return conv_forward_impl(d, in, packed, bias, gamma, beta, out_act, stream, &wn);
}


// ------------------------------------------------------------------------------------------------
// Plain GEMM  C[M][N] (fp32) = A[M][K] * B[N][K]^T  on the same kernel (one filter tap, K-major 16-bit
// operands, optional split-K with fp32 atomics).  Used by the weight-gradient path: A = gradient (or
// activation) matrix transposed to [channels][pixels], B = the transposed im2col matrix.
// ------------------------------------------------------------------------------------------------
extern "C" int hfc_gemm_nt(const void* a, int32_t a_bf16, const void* b, int32_t b_bf16, int32_t m, int32_t n,
                           int32_t k, float* c, int32_t ldc, int32_t k_splits, void* stream) {
  if (!a || !b || !c || m <= 0 || n <= 0 || k <= 0) return set_error(HFC_ERR_INVALID, "gemm_nt: null pointer or empty matrix");
  if (k % kBlockK != 0) return set_error(HFC_ERR_INVALID, "gemm_nt: K (%d) must be a multiple of 64", k);
  if (ldc % 4 != 0 || ldc < n) return set_error(HFC_ERR_INVALID, "gemm_nt: ldc must be a multiple of 4 and >= N");
  if ((a_bf16 != 0) != (b_bf16 != 0))
    return set_error(HFC_ERR_UNSUPPORTED, "gemm_nt: tcgen05 kind::f16 faults on mixed fp16 x bf16 operands; convert one side");
  int sm_count = 0;
  int rc = device_sm_count(&sm_count);
  if (rc != HFC_OK) return rc;
  PFN_encodeTiled encode = get_encode_fn();
  if (!encode) return set_error(HFC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not found");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  ConvKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  const int nt = (n + 255) / 256;
  kp.block_n = ((n + nt - 1) / nt + 15) / 16 * 16;
  kp.n_tiles = (n + kp.block_n - 1) / kp.block_n;
  kp.tw = kBlockM; kp.th = 1; kp.tn = 1;
  kp.tiles_w = (m + kBlockM - 1) / kBlockM; kp.tiles_h = 1; kp.tiles_n = 1;
  kp.c_chunks = k / kBlockK; kp.ntaps = 1;
  const int tiles = kp.tiles_w * kp.n_tiles;
  int ks = k_splits;
  if (ks <= 0) ks = std::max(1, std::min(sm_count / std::max(1, tiles), kp.c_chunks / 8));
  ks = std::max(1, std::min(ks, kp.c_chunks));
  kp.kb_per_split = (kp.c_chunks + ks - 1) / ks;
  kp.k_splits = (kp.c_chunks + kp.kb_per_split - 1) / kp.kb_per_split;
  kp.num_kb = kp.kb_per_split;
  kp.out_atomic = kp.k_splits > 1;
  kp.gemm = 1;
  kp.cm = kp.cn = 1; kp.b_rows = kp.block_n; kp.nsub = 1;
  kp.grid_h = 1; kp.grid_w = m; kp.batch = 1; kp.sh = kp.sw = 1;
  kp.out_mode = HFC_OUT_NHWC_F32; kp.osh = kp.osw = 1;
  kp.out_h = 1; kp.out_w = m; kp.out_cpad = ldc; kp.cout = n;
  kp.fmt_a = a_bf16 ? 1u : 0u; kp.fmt_b = b_bf16 ? 1u : 0u;
  kp.out = c;
  kp.kw = 1; kp.w_step = kBlockM;
  const int stage_bytes = kABytes + kp.block_n * kBlockK * 2;
  kp.stages = std::max(2, std::min((226 * 1024 - 1024 - kTailBytes) / stage_bytes, kMaxStages));

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(m), 1, 1};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(k) * 2, static_cast<cuuint64_t>(m) * k * 2,
                             static_cast<cuuint64_t>(m) * k * 2};
    cuuint32_t box[4] = {kBlockK, kBlockM, 1, 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(HFC_ERR_LAUNCH, "gemm_nt: cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    cuuint64_t bdims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(n)};
    cuuint64_t bstr[1] = {static_cast<cuuint64_t>(k) * 2};
    cuuint32_t bbox[2] = {kBlockK, static_cast<cuuint32_t>(kp.block_n)}, bes[2] = {1, 1};
    r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(b), bdims, bstr, bbox, bes,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(HFC_ERR_LAUNCH, "gemm_nt: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  if (kp.out_atomic) {
    cudaError_t e = cudaMemsetAsync(c, 0, static_cast<size_t>(m) * ldc * sizeof(float), st);
    if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "gemm_nt: memset: %s", cudaGetErrorString(e));
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int work = tiles * kp.k_splits;
  const int grid = std::min(work, sm_count);
  const size_t smem = static_cast<size_t>(kp.stages) * stage_bytes + 1024 + kTailBytes;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false, 1>, tmA, tmB, kp);
  if (e != cudaSuccess) return set_error(HFC_ERR_LAUNCH, "gemm_nt launch: %s", cudaGetErrorString(e));
  note_launch();
  return HFC_OK;
}
