// Error plumbing shared by all translation units of libeosvos_b200.so.
// Convention: every entry point returns 0 or a negative code and never throws;
// eosvos_last_error() returns the message of the last failure on the calling thread.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define EOSVOS_OK 0
#define EOSVOS_ERR_ARG (-1)
#define EOSVOS_ERR_CUDA (-2)
#define EOSVOS_ERR_ARCH (-3)

namespace eosvos {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* name);
int num_sms();
bool pdl_enabled();   // EOSVOS_PDL != 0 (default on)

// Programmatic dependent launch: the kernel may be scheduled while the previous kernel of the stream drains (its launch
// latency and prologue overlap that kernel's tail); it must execute pdl_wait() before it touches global memory the
// previous kernels produce or still read.  Inside stream capture the edge becomes a programmatic graph dependency.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
// all memory operations of the kernels this launch depends on are complete and visible after this returns
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next kernel of the stream be scheduled as soon as every CTA of this grid has got here (or exited)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
}  // namespace eosvos

#define EOSVOS_REQUIRE(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) return ::eosvos::set_error(EOSVOS_ERR_ARG, msg);   \
  } while (0)
#define EOSVOS_TRY(expr)        \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)
