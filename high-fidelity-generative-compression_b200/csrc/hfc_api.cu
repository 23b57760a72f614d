// Library-level entry points of libhfc: version, error reporting, device checks, launch counter.
#include "hfc_internal.h"

namespace hfc {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int device_sm_count(int* sm_count) {
  static int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return set_error(HFC_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess)
      return set_error(HFC_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
      return set_error(HFC_ERR_NO_DEVICE, "libhfc is built for sm_100a only; device is sm_%d%d",
                       prop.major, prop.minor);
    cached_dev = dev;
    cached_sms = prop.multiProcessorCount;
  }
  *sm_count = cached_sms;
  return HFC_OK;
}

void note_launch() { ++g_launches; }

}  // namespace hfc

extern "C" int hfc_abi_version(void) { return HFC_ABI_VERSION; }

extern "C" const char* hfc_last_error(void) { return hfc::g_err; }

extern "C" int hfc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return hfc::set_error(HFC_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess)
    return hfc::set_error(HFC_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return HFC_OK;
}

extern "C" unsigned long long hfc_launch_count(void) { return hfc::g_launches; }
