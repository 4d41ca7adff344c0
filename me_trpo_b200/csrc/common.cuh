// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include "metrpo.h"

namespace metrpo {

char* last_error_buf();  // thread-local, 512 bytes (api.cu)

inline int set_error(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return status;
}

#define METRPO_CUDA_OK(expr)                                                               \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::metrpo::set_error(METRPO_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, \
                                 cudaGetErrorString(_e));                                  \
  } while (0)

}  // namespace metrpo
