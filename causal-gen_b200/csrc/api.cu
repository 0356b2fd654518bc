// Version, error text and device gate of the C ABI.
#include <cstdarg>
#include <cstdio>

#include "cg_common.cuh"

static thread_local char g_err[512] = "";

void cg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int cg_version(void) { return 100; }
extern "C" const char* cg_last_error(void) { return g_err; }

extern "C" int cg_device_sms(void) {
  static int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev == cached_dev) return cached_sms;
  int major = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cached_dev = dev;
  cached_sms = (major == 10) ? sms : 0;
  return cached_sms;
}

int cg_require_sm100() {
  if (cg_device_sms() <= 0) {
    cg_set_error("causalgen_b200 needs an sm_100 (B200) device; there is no CPU or other-arch fallback");
    return CG_ERR_ARCH;
  }
  return CG_OK;
}

// debug builds (-DCG_TIMELINE): kernels of CTA (0,0) write %globaltimer marks into this device buffer
unsigned long long* cg_tl_ptr = nullptr;
extern "C" void cg_debug_timeline(unsigned long long* dev_buf) { cg_tl_ptr = dev_buf; }
