// ABI bookkeeping entry points (include/cilrs_b200.h).
#include "common.cuh"
#include "../../include/cilrs_b200.h"
#include <stdio.h>

#include <stdlib.h>
namespace cilrs {
long long g_cilrs_launches = 0;
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("CILRS_NO_PDL") ? 0 : 1;
  return on != 0;
}
}  // namespace cilrs

extern "C" {

long long cilrs_launch_count(void) { return cilrs::g_cilrs_launches; }

int cilrs_abi_version(void) { return CILRS_ABI_VERSION; }

const char* cilrs_status_string(int status) {
  static thread_local char buf[160];
  switch (status) {
    case cilrs::OK: return "ok";
    case cilrs::ERR_INVALID: return "invalid argument";
    case cilrs::ERR_UNSUPPORTED: return "unsupported shape or configuration";
    case cilrs::ERR_WORKSPACE: return "workspace too small";
    case cilrs::ERR_DRIVER: return "CUDA driver entry point (cuTensorMapEncodeTiled) unavailable";
    default: break;
  }
  if (status >= cilrs::ERR_CUDA_BASE) {
    snprintf(buf, sizeof(buf), "CUDA error %d: %s", status - cilrs::ERR_CUDA_BASE,
             cudaGetErrorString((cudaError_t)(status - cilrs::ERR_CUDA_BASE)));
    return buf;
  }
  return "unknown status";
}

}  // extern "C"
