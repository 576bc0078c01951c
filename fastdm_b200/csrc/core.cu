// Error plumbing, device checks and version string of libfastdm_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fdm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return FDM_ERR_CUDA;
}

struct DevInfo {
  int major = -1;
  int sms = 0;
};
static DevInfo g_dev[64];

static int dev_info(DevInfo** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= 64) {
    set_error("device index %d out of range", dev);
    return FDM_ERR_ARG;
  }
  DevInfo& d = g_dev[dev];
  if (d.major < 0) {
    int major = 0, sms = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(cc major)");
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(sm count)");
    d.sms = sms;
    d.major = major;
  }
  *out = &d;
  return FDM_OK;
}

int require_sm100() {
  DevInfo* d = nullptr;
  int rc = dev_info(&d);
  if (rc) return rc;
  if (d->major != 10) {
    set_error("fastdm_b200 kernels are built for sm_100a only; current device has compute capability %d.x",
              d->major);
    return FDM_ERR_ARCH;
  }
  return FDM_OK;
}

int num_sms() {
  DevInfo* d = nullptr;
  if (dev_info(&d)) return 148;
  return d->sms > 0 ? d->sms : 148;
}

}  // namespace fdm

extern "C" {

const char* fdm_last_error(void) { return fdm::g_err; }

const char* fdm_version(void) { return "fastdm_b200 0.1.0 sm_100a"; }

int fdm_check_device(int dev) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return fdm::cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    fdm::set_error("device %d has compute capability %d.x, need 10.x (B200)", dev, major);
    return FDM_ERR_ARCH;
  }
  return FDM_OK;
}

}  // extern "C"
