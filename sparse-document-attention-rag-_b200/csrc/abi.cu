// Library-level entry points of the C ABI (include/sdag_b200.h).
#include "common.cuh"

extern "C" {

int sdag_b200_version(void) { return 100; }

const char* sdag_b200_error_string(int code) {
  switch (code) {
    case SDAG_OK: return "ok";
    case SDAG_ERR_INVALID: return "invalid argument";
    case SDAG_ERR_UNSUPPORTED: return "unsupported shape";
    case SDAG_ERR_WORKSPACE: return "workspace too small";
    case SDAG_ERR_DRIVER: return "tensor-map encoding failed (driver entry point unavailable?)";
    default: return cudaGetErrorString(static_cast<cudaError_t>(code));
  }
}

}  // extern "C"
