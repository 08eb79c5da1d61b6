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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per-device facts: launch sites keep a small table
// indexed by device ordinal instead of a process-wide "configured" flag, so a second GPU used from the same process gets its
// own opt-in.  device_slot() returns the current device's ordinal (callers make the tensor's device current).
constexpr int SDAG_MAX_DEVICES = 64;
inline int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SDAG_MAX_DEVICES) return -1;
  return dev;
}
inline int device_sm_count() {
  static int n_sm[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return 148;
  if (n_sm[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    n_sm[dev] = n;
  }
  return n_sm[dev];
}

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

// x * Phi(x) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 on erf, i.e. <= 0.75e-7 * |x| on the result —
// three orders of magnitude below a bf16 ulp of the output wherever the output is not itself below 1e-6): one reciprocal,
// one exponential and five FMAs instead of erff's two divergent polynomial branches, which made this pass compute-bound.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfc_z = p * t * __expf(-z * z);           // 1 - erf(|x| / sqrt 2)
  const float phi = x >= 0.f ? 1.f - 0.5f * erfc_z : 0.5f * erfc_z;
  return x * phi;
}

