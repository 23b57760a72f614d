// Device functions of the conditional (mean-scale) likelihood shared by the training / evaluation kernel
// (elementwise.cu) and the compress-path kernel (symbols.cu): src/hyperprior.py:124-139, src/helpers/maths.py:102-109.
#pragma once
#include <cuda_runtime.h>

namespace hfc {

// erfc with fractional error < 1.2e-7 everywhere (Chebyshev fit of Numerical Recipes' erfcc): one
// exp, one reciprocal and a 9-term Horner chain, about half the instructions of erfcf.  The likelihood
// kernel evaluates four of these per element and is otherwise instruction-bound, not HBM-bound.
__device__ __forceinline__ float fast_erfc(float x) {
  const float z = fabsf(x);
  const float t = __fdividef(1.f, fmaf(0.5f, z, 1.f));
  float pl = 0.17087277f;
  pl = fmaf(pl, t, -0.82215223f);
  pl = fmaf(pl, t, 1.48851587f);
  pl = fmaf(pl, t, -1.13520398f);
  pl = fmaf(pl, t, 0.27886807f);
  pl = fmaf(pl, t, -0.18628806f);
  pl = fmaf(pl, t, 0.09678418f);
  pl = fmaf(pl, t, 0.37409196f);
  pl = fmaf(pl, t, 1.00002368f);
  pl = fmaf(pl, t, -1.26551223f);
  const float r = t * __expf(fmaf(-z, z, pl));
  return x >= 0.f ? r : 2.f - r;
}

__device__ __forceinline__ float std_cdf(float v, int type) {
  if (type == 0) return 0.5f * fast_erfc(v * -0.70710678118654752440f);   // maths.py:102-105
  return __fdividef(1.f, 1.f + __expf(-v));                                 // maths.py:107-109
}

}  // namespace hfc
