// Library-level entry points: version, error string, device check.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

namespace dsf {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels launched through the C ABI (bench.py's gpu_launches)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return DSF_ELAUNCH;
  }
  return DSF_OK;
}

static int g_pdl = 1;
static int g_sm_margin = 0;  // SMs left to concurrent kernels of other streams (NCCL), see dsf_set_sm_margin
bool pdl_enabled() { return g_pdl != 0; }

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
    if (cached <= 0) cached = 148;
  }
  return std::max(2, cached - g_sm_margin);
}

}  // namespace dsf

extern "C" int dsf_version(void) { return DSF_VERSION; }
extern "C" int dsf_set_sm_margin(int32_t sms) {
  if (sms < 0 || sms > 64 || (sms & 1)) {
    dsf::set_error("set_sm_margin: margin must be an even number in [0, 64]");
    return DSF_EINVAL;
  }
  dsf::g_sm_margin = sms;
  return DSF_OK;
}
extern "C" int dsf_set_pdl(int32_t on) {
  dsf::g_pdl = on ? 1 : 0;
  return DSF_OK;
}
extern "C" int64_t dsf_launch_count(void) { return (int64_t)__atomic_load_n(&dsf::g_launches, __ATOMIC_RELAXED); }
extern "C" const char* dsf_last_error(void) { return dsf::g_err; }

extern "C" int dsf_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    dsf::set_error("no CUDA device");
    return DSF_EARCH;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    dsf::set_error("libdsfuse is built for sm_100a only; device has compute capability %d.x", major);
    return DSF_EARCH;
  }
  return DSF_OK;
}
