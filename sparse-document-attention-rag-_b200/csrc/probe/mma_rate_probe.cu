// Bring-up probe (not part of the product library): issue / completion rate of tcgen05.mma chains on sm_100a.
// One thread per CTA issues `reps` MMAs of shape 128 x N x 16 (bf16 -> fp32) and commits; clock64 is read before the
// first issue, after the last issue and after the commit barrier completes.  Operand contents are irrelevant (zeros).
//   usage: mma_rate_probe            (prints a table: N, SS/TS, B major, accumulators, CTAs per SM)
#include <stdio.h>
#include <stdlib.h>

#include "../sm100_ptx.cuh"

using namespace sm100;

struct Cfg {
  int N, ts, b_mn, n_acc, reps;
};

__global__ void __launch_bounds__(128, 2) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // 32 KB
  uint8_t* sB = smem + 32768;   // 64 KB (N up to 256)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32768 + 65536) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, c.N, 0, c.b_mn);
    const uint64_t dA = umma_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t dB = c.b_mn ? umma_smem_desc_sw128(smem_u32(sB), 64 * 128, 1024) : umma_smem_desc_sw128(smem_u32(sB), 16, 1024);
    const uint32_t acc_stride = c.N * c.n_acc <= 192 ? c.N : 0;  // second accumulator if it fits below the A columns
    long long t0 = 0, t1 = 0, t2 = 0;
    __syncwarp();
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < c.reps; ++r) {
        const int ks = r & 3;
        const uint32_t d = tmem + (r % c.n_acc) * acc_stride;
        if (c.ts)
          umma_ts(d, tmem + 192 + ks * 8, dB + (c.b_mn ? (ks * 16 * 128) >> 4 : (ks * 32) >> 4), idesc, 1);
        else
          umma_ss(d, dA + ((ks * 32) >> 4), dB + (c.b_mn ? (ks * 16 * 128) >> 4 : (ks * 32) >> 4), idesc, 1);
      }
      t1 = clock64();
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait_spin(&bar, 0);
    t2 = clock64();
    if (lane_id() == 0 && blockIdx.x == 0) {
      out[0] = __shfl_sync(1u, t1, 0) - __shfl_sync(1u, t0, 0);
    }
    // elected lane may not be lane 0: reduce through max over the warp
    long long issue = t1 - t0, total = t2 - t0;
    for (int o = 16; o > 0; o >>= 1) {
      issue = max(issue, __shfl_xor_sync(0xffffffffu, issue, o));
      t0 = max(t0, __shfl_xor_sync(0xffffffffu, t0, o));
    }
    total = t2 - t0;
    if (lane_id() == 0 && blockIdx.x == 0) {
      out[0] = issue;
      out[1] = total;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  printf("%5s %3s %5s %4s %8s | %10s %10s\n", "N", "TS", "B_MN", "acc", "CTAs/SM", "issue/mma", "total/mma");
  const int reps = 512;
  for (int per_sm = 1; per_sm <= 2; ++per_sm)
    for (int ts = 0; ts <= 1; ++ts)
      for (int b_mn = 0; b_mn <= 1; ++b_mn)
        for (int N : {32, 64, 128, 256})
          for (int n_acc = 1; n_acc <= 2; ++n_acc) {
            if (N * n_acc > 192 && n_acc > 1) continue;
            if (N == 256 && b_mn) continue;
            Cfg c{N, ts, b_mn, n_acc, reps};
            long long h[2] = {0, 0};
            for (int it = 0; it < 2; ++it) {
              rate_kernel<<<n_sm * per_sm, 128, smem>>>(c, out);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("error %s\n", cudaGetErrorString(e));
                return 1;
              }
            }
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("%5d %3d %5d %4d %8d | %10.1f %10.1f\n", N, ts, b_mn, n_acc, per_sm, double(h[0]) / reps, double(h[1]) / reps);
          }
  return 0;
}
