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
