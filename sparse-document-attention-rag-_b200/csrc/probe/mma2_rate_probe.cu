// Bring-up probe (not part of the product library): rate of tcgen05.mma.cta_group::2 chains (M = 256 over a CTA pair,
// SS operands, bf16 -> fp32) issued by the leader CTA of each cluster; operands are zeros.  Prints cycles per MMA.
#include <stdio.h>

#include "../sm100_ptx.cuh"

using namespace sm100;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(int N, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc_2cta(&tmem_base_s, 512);
    tmem_relinquish_2cta();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1 && cluster_ctarank() == 0) {
    const uint32_t idesc = umma_idesc_bf16_f32(256, N, 0, 0);
    const uint64_t dA = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t dB = umma_smem_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) umma_ss_2cta(tmem, dA + (((r & 3) * 32) >> 4), dB + (((r & 3) * 32) >> 4), idesc, 1);
      t1 = clock64();
      umma_commit_2cta(&bar, 1);
    }
    __syncwarp();
    mbar_wait_spin(&bar, 0);
    t2 = clock64();
    long long issue = t1 - t0;
    for (int o = 16; o > 0; o >>= 1) {
      issue = max(issue, __shfl_xor_sync(0xffffffffu, issue, o));
      t0 = max(t0, __shfl_xor_sync(0xffffffffu, t0, o));
    }
    if (lane_id() == 0 && blockIdx.x == 0) {
      out[0] = issue;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2cta(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int smem = 65536 + 1024;
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 512;
  for (int N : {64, 128, 256}) {
    long long h[2] = {0, 0};
    for (int it = 0; it < 2; ++it) {
      rate2_kernel<<<148, 128, smem>>>(N, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("error %s\n", cudaGetErrorString(e));
        return 1;
      }
    }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("cta_group::2 SS M=256 N=%3d: issue %.1f  total %.1f cycles / MMA (74 pairs busy)\n", N, double(h[0]) / reps, double(h[1]) / reps);
  }
  return 0;
}
