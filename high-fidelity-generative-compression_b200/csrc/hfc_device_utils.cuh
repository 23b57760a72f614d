// Small device helpers shared by the elementwise kernels.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace hfc {

// Positions (in a buffer with a materialised border of p_lo / p_hi) that logical coordinate `o`
// of an axis of length `size` must be written to: its own slot plus, when `reflect`, the slots of
// the border that mirror it (ReflectionPad2d semantics: border[-k] = x[k], border[size-1+k] = x[size-1-k]).
__device__ __forceinline__ int mirror_targets(int o, int size, int p_lo, int p_hi, bool reflect,
                                              int (&out)[3]) {
  int n = 0;
  out[n++] = o + p_lo;
  if (reflect) {
    if (o >= 1 && o <= p_lo) out[n++] = p_lo - o;
    if (o <= size - 2 && o >= size - 1 - p_hi) out[n++] = p_lo + 2 * (size - 1) - o;
  }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  return v;
}

}  // namespace hfc
