// Thread-local error buffer + version string of libmetrpo.so.
#include "common.cuh"

namespace metrpo {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace metrpo

extern "C" const char* metrpo_last_error(void) { return metrpo::last_error_buf(); }
extern "C" const char* metrpo_version(void) {
  return "metrpo-b200 0.1 (sm_100a; tcgen05 + bulk-copy persistent rollout)";
}
