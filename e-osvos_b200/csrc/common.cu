// Error state, device checks and library identification for libeosvos_b200.so.
#include "common.h"
#include <stdlib.h>
#include "../../include/eosvos_b200.h"
#include <string.h>
#include <cuda.h>
#include "act.cuh"

namespace eosvos {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof g_err, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return EOSVOS_ERR_CUDA;
}
static unsigned long long g_launches = 0;
int check_launch(const char* name) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, name);
  return 0;
}
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("EOSVOS_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace eosvos

extern "C" const char* eosvos_last_error(void) { return eosvos::g_err; }

// number of kernel launches issued by this library so far (bench.py's gpu_launches)
extern "C" unsigned long long eosvos_launch_count(void) { return eosvos::g_launches; }

// 0 = bfloat16, 1 = float16: storage type of activations / tensor-core operands in this build
extern "C" int eosvos_act_dtype(void) { return EOSVOS_ACT_CODE; }

extern "C" int eosvos_version(void) { return EOSVOS_B200_VERSION; }

// sm_100 only: hard error elsewhere, there is no fallback path.
extern "C" int eosvos_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return eosvos::set_error(EOSVOS_ERR_ARCH, "no CUDA device visible");
  if (device < 0 || device >= n) return eosvos::set_error(EOSVOS_ERR_ARG, "device index out of range");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) {
    char buf[128];
    snprintf(buf, sizeof buf, "device %d is sm_%d%d; libeosvos_b200 is built for sm_100a only", device, major, minor);
    return eosvos::set_error(EOSVOS_ERR_ARCH, buf);
  }
  return 0;
}
