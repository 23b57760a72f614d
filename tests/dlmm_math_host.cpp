// TEST INFRASTRUCTURE: compiles the per-element code of the `-LMM` kernels (csrc/dlmm_math.cuh -- the very functions
// dlmm_likelihood_kernel / dlmm_likelihood_bwd_kernel call per thread) with g++ and runs them over a whole tensor on the
// host, so that the arithmetic and the index decomposition the GPU executes are checked on the CPU
// (tests/test_dlmm_cpu.py).  Only the grid-stride loop and the block reduction of csrc/dlmm.cu are not covered here.
#include "../high-fidelity-generative-compression_b200/csrc/dlmm_math.cuh"

extern "C" void dlmm_forward_host(const float* x, const float* noise, const float* params, int n, int c, int k, int hw,
                                  int type, int straight_through, float* decoded, double* sums) {
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  for (int64_t i = 0; i < count; ++i) {
    float an = 0.f, aq = 0.f;
    hfc::dlmm_element_fwd(i, x, noise, params, c, k, hw, type, straight_through, decoded, &an, &aq);
    sums[0] += an;
    sums[1] += aq;
  }
}

extern "C" void dlmm_backward_host(const float* x, const float* noise, const float* params, const float* d_decoded, float g,
                                   int n, int c, int k, int hw, int type, float* dx, float* dparams) {
  const int64_t count = static_cast<int64_t>(n) * c * hw;
  for (int64_t i = 0; i < count; ++i)
    hfc::dlmm_element_bwd(i, x, noise, params, d_decoded, g, c, k, hw, type, dx, dparams);
}
