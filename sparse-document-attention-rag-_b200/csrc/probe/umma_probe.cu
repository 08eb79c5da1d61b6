// Bring-up probe for the sm_100a building blocks (not part of the product library).
// One single-CTA GEMM D[128 x N] = A[128 x K] * B^T, run in several operand modes, each checked
// against a host computation. Usage: umma_probe <mode>
//   1  A,B K-major via TMA (SW128)                     -> retrieval GEMM, Q*K^T
//   2  B MN-major via TMA (LBO = n-block stride, SBO = 1024)   -> P*V with V in its natural [kv][d] layout
//   3  B MN-major, LBO/SBO swapped (diagnostic)
//   4  A written by threads with the SW128 formula, B K-major via TMA -> P staged through smem
//   5  A from TMEM (tcgen05.st, bf16x2 packed), B MN-major via TMA  -> P kept in TMEM
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../sm100_ptx.cuh"

using namespace sm100;

constexpr int M = 128;
constexpr int N = 128;
constexpr int K = 128;

struct Params {
  int mode;
  const __nv_bfloat16* A;  // [M][K] row-major (modes 4,5 read it directly)
  float* D;                // [M][N]
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 2 k-blocks x [128 rows x 128 B]
  uint8_t* sB = smem + 32768;      // K-major: 2 k-blocks x [N rows x 128 B]; MN-major: 2 n-blocks x [K rows x 128 B]
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool mn_major_b = (p.mode == 2 || p.mode == 3 || p.mode == 5);

  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;        // columns [0,128): accumulator
  const uint32_t tmem_a = tmem + 128;  // columns [128,192): A as bf16x2 (mode 5)

  if (tid == 0) {
    uint32_t bytes = 32768;  // B always by TMA
    if (p.mode <= 3) bytes += 32768;
    mbar_expect_tx(&bar_full, bytes);
    if (p.mode <= 3) {
      tma_load_2d(sA, &tmA, &bar_full, 0, 0);
      tma_load_2d(sA + 16384, &tmA, &bar_full, 64, 0);
    }
    if (!mn_major_b) {
      tma_load_2d(sB, &tmB, &bar_full, 0, 0);
      tma_load_2d(sB + 16384, &tmB, &bar_full, 64, 0);
    } else {
      tma_load_2d(sB, &tmB, &bar_full, 0, 0);           // n 0..63, all K rows
      tma_load_2d(sB + 16384, &tmB, &bar_full, 64, 0);  // n 64..127
    }
  }
  if (p.mode == 4) {
    // thread r writes row r of A into both k-blocks with the swizzle formula
    const uint4* src = reinterpret_cast<const uint4*>(p.A + tid * K);
    for (int kb = 0; kb < 2; ++kb)
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_offset(tid, c)) = src[kb * 8 + c];
    fence_proxy_async_smem();
  }
  if (p.mode == 5) {
    // thread r owns TMEM lane r: write K bf16 as K/2 packed 32-bit columns
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.A + tid * K);
    uint32_t regs[32];
    for (int half = 0; half < 2; ++half) {
      for (int i = 0; i < 32; ++i) regs[i] = src[half * 32 + i];
      tmem_st_32x32b_x32(tmem_a + (uint32_t(warp * 32) << 16) + half * 32, regs);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16_f32(M, N, 0, mn_major_b ? 1 : 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      const int kb = ks / 4, kin = ks % 4;
      uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA) + kb * 16384 + kin * 32, 16, 1024);
      uint64_t bdesc;
      if (!mn_major_b) {
        bdesc = umma_smem_desc_sw128(smem_u32(sB) + kb * 16384 + kin * 32, 16, 1024);
      } else {
        const uint32_t start = smem_u32(sB) + ks * 16 * 128;  // 16 k-rows of 128 B
        bdesc = (p.mode == 3) ? umma_smem_desc_sw128(start, 1024, 16384) : umma_smem_desc_sw128(start, 16384, 1024);
      }
      if (p.mode == 5)
        umma_ts(tmem_d, tmem_a + ks * 8, bdesc, idesc, ks > 0);
      else
        umma_ss(tmem_d, adesc, bdesc, idesc, ks > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_d + (uint32_t(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.D[tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  const bool mn_major_b = (mode == 2 || mode == 3 || mode == 5);
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);  // fB[n*K + k]
  srand(1234);
  for (int i = 0; i < M * K; ++i) {
    fA[i] = float(rand() % 9 - 4);
    hA[i] = __float2bfloat16(fA[i]);
  }
  for (int i = 0; i < N * K; ++i) fB[i] = float(rand() % 7 - 3);
  // device layout of B: K-major [N][K] or MN-major [K][N]
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 v = __float2bfloat16(fB[n * K + k]);
      if (mn_major_b)
        hB[k * N + n] = v;
      else
        hB[n * K + k] = v;
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, M * K * 2));
  CK(cudaMalloc(&dB, N * K * 2));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {K, M};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, M};
    if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box)) return printf("tmap A failed\n"), 2;
  }
  if (!mn_major_b) {
    uint64_t dims[2] = {K, N};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, N};
    if (make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box)) return printf("tmap B failed\n"), 2;
  } else {
    uint64_t dims[2] = {N, K};
    uint64_t str[1] = {N * 2};
    uint32_t box[2] = {64, K};
    if (make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box)) return printf("tmap B failed\n"), 2;
  }
  Params p{mode, dA, dD};
  const int smem_bytes = 65536 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmB, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += fA[m * K + k] * fB[n * K + k];
      double e = fabs(double(ref) - double(hD[m * N + n]));
      if (!(e <= 1e-3)) {
        if (bad < 5) printf("  mismatch (%d,%d): got %f want %f\n", m, n, hD[m * N + n], ref);
        ++bad;
      }
      if (e > maxerr) maxerr = e;
    }
  printf("PROBE mode %d: %s (bad=%d maxerr=%g)\n", mode, bad ? "FAIL" : "PASS", bad, maxerr);
  return bad ? 1 : 0;
}
