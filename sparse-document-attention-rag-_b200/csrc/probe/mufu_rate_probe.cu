// Bring-up probe: issue rate of the exp unit on sm_100a for ex2.approx.ftz.f32, .f16x2 and .bf16x2 (does a packed
// exponential give two results per exp-unit slot?), and of the packed FMA / conversion instructions the softmax uses.
// 16 warps per SM (4 per scheduler), 8 independent chains per thread; prints cycles per warp-instruction per scheduler.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(512) probe(uint32_t* out, int iters, long long* cycles) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0x3c003c00u + threadIdx.x + i;  // small positive halves / harmless floats
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = -0.5f - 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) asm volatile("{.reg .b64 t; mov.b64 t, {%0, %1}; fma.rn.f32x2 t, t, t, t; mov.b64 {%0, %1}, t;}" : "+f"(f[i]), "+f"(f[(i + 1) & 7]));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(x[i]) : "f"(f[i]), "f"(f[(i + 1) & 7]));
      if (MODE == 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x[i]) : "f"(f[i]), "f"(f[(i + 1) & 7]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= x[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  uint32_t* out;
  long long* cyc;
  const int blocks = 148, iters = 4096;
  cudaMalloc(&out, blocks * 512 * 4);
  cudaMalloc(&cyc, blocks * 8);
  probe<MODE><<<blocks, 512>>>(out, iters, cyc);
  probe<MODE><<<blocks, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += h[i];
  mean /= blocks;
  // per scheduler: 4 warps x 8 instructions x iters warp-instructions
  printf("%-28s %8.2f cycles per warp-instruction per scheduler (%s)\n", name, mean / (4.0 * 8 * iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("ex2.approx.f16x2");
  run<2>("ex2.approx.ftz.bf16x2");
  run<3>("fma.rn.f32x2");
  run<4>("fma.rn.f32");
  run<5>("cvt.rn.bf16x2.f32");
  run<6>("cvt.rn.f16x2.f32");
  return 0;
}
