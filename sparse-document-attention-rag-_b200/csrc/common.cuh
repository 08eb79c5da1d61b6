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

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still draining; it must call pdl_wait() before touching anything the
// predecessor wrote (or still reads).  pdl_trigger() in the predecessor lets the dependent grid be scheduled as soon as
// every CTA of this grid has issued it (the dependent then fills SM slots as this grid's CTAs retire).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
