// Shared includes for the library translation units.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdag_b200.h"

#define SDAG_CUDA_TRY(expr)                              \
  do {                                                   \
    cudaError_t e__ = (expr);                            \
    if (e__ != cudaSuccess) return static_cast<int>(e__); \
  } while (0)
