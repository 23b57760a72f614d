// Internal helpers shared by the libhfc translation units (not part of the C ABI).
#pragma once
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/hfc.h"

namespace hfc {

// Records a thread-local message and returns `code` (so call sites can `return set_error(...)`).
int set_error(int code, const char* fmt, ...);
// Number of SMs of the current device; HFC_ERR_NO_DEVICE if it is not an sm_100 part.
int device_sm_count(int* sm_count);
// Launch accounting (bench.py reports it as "gpu_launches").
void note_launch();
// Second schedule of the Gaussian latent-likelihood kernel (likelihood_v2.cu); same contract as hfc_latent_likelihood.
int launch_latent_likelihood_v2(const float* y, const float* mean, const float* scale_raw, const float* noise,
                                int64_t count, float lb, float* decoded, double* sums, int sms, bool prefetch,
                                cudaStream_t st);
// Schedule 4: inputs staged by cp.async.bulk (16-byte aligned pointers only).
int launch_latent_likelihood_bulk(const float* y, const float* mean, const float* scale_raw, const float* noise,
                                  int64_t count, float lb, float* decoded, double* sums, cudaStream_t st);

}  // namespace hfc
