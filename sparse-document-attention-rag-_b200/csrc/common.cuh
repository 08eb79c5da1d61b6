// Shared includes for the library translation units.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

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

// Host: launch `kernel` as a programmatic dependent of the preceding kernel in `st` (SDAG_NO_PDL=1: plain launch).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static int use_pdl = -1;
  if (use_pdl < 0) {
    const char* e = getenv("SDAG_NO_PDL");
    use_pdl = (e && atoi(e)) ? 0 : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
