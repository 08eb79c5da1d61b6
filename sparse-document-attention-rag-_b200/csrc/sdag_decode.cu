// SDAG decode attention for sm_100a (path A5): one new token attends to every cached position — the
// reference passes no mask after the prefill (SDAG.py:193-197), so decode is dense over the KV cache and
// HBM-bound.  Split-KV: each CTA streams a contiguous range of cache rows of one kv head for all G = H/Hkv
// query heads sharing it (K/V bytes are read exactly once), writes an (m, l, acc) partial; a second kernel
// merges the partials.  CUDA-core arithmetic on purpose: 2*G FLOP per cache byte is far below the tensor
// ridge.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int TILE_ROWS = 64;
constexpr int MAX_G = 8;
constexpr int MAX_CHUNK = 1024;  // cache rows per CTA
constexpr int THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = pred ? 16 : 0;  // zero-fill when out of range
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int D>
struct DecodeSmem {
  static constexpr int kRowBytes = D * 2 + 16;  // +16 B: conflict-free 16-byte row-strided reads
  static constexpr int kTile = TILE_ROWS * kRowBytes;
  static constexpr int kTiles = 2 * kTile;
  static constexpr int kScores = MAX_CHUNK * MAX_G * 4;
  static constexpr int kQ = MAX_G * D * 4;
  static constexpr int kTotal = kTiles + kScores + kQ;
};

template <int D>
__device__ __forceinline__ void load_tile(uint8_t* dst, const __nv_bfloat16* base, long long s_l, int row0, int row_end) {
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < TILE_ROWS * CPR; i += THREADS) {
    const int r = i / CPR, c = i % CPR;
    const bool ok = row0 + r < row_end;
    const __nv_bfloat16* src = base + static_cast<long long>(ok ? row0 + r : row0) * s_l + c * 8;
    cp_async16(dst + r * DecodeSmem<D>::kRowBytes + c * 16, src, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(THREADS)
decode_split_kernel(const __nv_bfloat16* __restrict__ q, long long sq_b, long long sq_h,
                    const __nv_bfloat16* __restrict__ k, long long sk_b, long long sk_h, long long sk_l,
                    const __nv_bfloat16* __restrict__ v, long long sv_b, long long sv_h, long long sv_l,
                    float* __restrict__ partial, int H, int Hkv, int kv_len_host, int rows_per_split, float scale_log2,
                    const int32_t* __restrict__ kv_len_dev) {
  const int kv_len = kv_len_dev ? *kv_len_dev : kv_len_host;  // device value: CUDA-graph replay per token
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* tiles = smem;
  float* sp = reinterpret_cast<float*>(smem + DecodeSmem<D>::kTiles);          // [MAX_CHUNK][MAX_G]
  float* sq = reinterpret_cast<float*>(smem + DecodeSmem<D>::kTiles + DecodeSmem<D>::kScores);  // [MAX_G][D]
  __shared__ float s_m[MAX_G], s_l[MAX_G];

  const int split = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int nsplit = gridDim.x;
  const int G = H / Hkv;
  const int tid = threadIdx.x;
  const int r0 = split * rows_per_split;
  const int r1 = min(kv_len, r0 + rows_per_split);
  const int n_rows = max(0, r1 - r0);
  const int n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;

  for (int i = tid; i < MAX_G * D; i += THREADS) {
    const int g = i / D, d = i % D;
    sq[i] = g < G ? __bfloat162float(q[b * sq_b + (hkv * G + g) * sq_h + d]) * scale_log2 : 0.f;
  }
  const __nv_bfloat16* kbase = k + b * sk_b + hkv * sk_h;
  const __nv_bfloat16* vbase = v + b * sv_b + hkv * sv_h;

  // ---------------- pass 1: scores (log2 domain) ----------------
  if (n_tiles > 0) load_tile<D>(tiles, kbase, sk_l, r0, r1);
  cp_async_commit();
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) load_tile<D>(tiles + ((t + 1) & 1) * DecodeSmem<D>::kTile, kbase, sk_l, r0 + (t + 1) * TILE_ROWS, r1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint8_t* tile = tiles + (t & 1) * DecodeSmem<D>::kTile;
    const int row = tid & (TILE_ROWS - 1);
    for (int g = tid / TILE_ROWS; g < G; g += THREADS / TILE_ROWS) {
      const uint4* kr = reinterpret_cast<const uint4*>(tile + row * DecodeSmem<D>::kRowBytes);
      const float4* qg = reinterpret_cast<const float4*>(sq + g * D);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        const uint4 w = kr[c];
        const float4 qa = qg[2 * c], qb = qg[2 * c + 1];
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
        const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
        const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.z));
        const float2 f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.w));
        acc = fmaf(f0.x, qa.x, acc);
        acc = fmaf(f0.y, qa.y, acc);
        acc = fmaf(f1.x, qa.z, acc);
        acc = fmaf(f1.y, qa.w, acc);
        acc = fmaf(f2.x, qb.x, acc);
        acc = fmaf(f2.y, qb.y, acc);
        acc = fmaf(f3.x, qb.z, acc);
        acc = fmaf(f3.y, qb.w, acc);
      }
      const int lr = t * TILE_ROWS + row;
      sp[lr * MAX_G + g] = lr < n_rows ? acc : -INFINITY;
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // start streaming V while the softmax statistics are computed
  if (n_tiles > 0) load_tile<D>(tiles, vbase, sv_l, r0, r1);
  cp_async_commit();

  // ---------------- softmax statistics per head: warp g handles head g ----------------
  {
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < G) {
      float mx = -INFINITY;
      for (int r = lane; r < n_rows; r += 32) mx = fmaxf(mx, sp[r * MAX_G + warp]);
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_use = mx == -INFINITY ? 0.f : mx;
      float sum = 0.f;
      for (int r = lane; r < n_rows; r += 32) {
        const float pv = exp2f(sp[r * MAX_G + warp] - m_use);
        sp[r * MAX_G + warp] = pv;
        sum += pv;
      }
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) {
        s_m[warp] = mx;
        s_l[warp] = sum;
      }
    }
  }
  __syncthreads();

  // ---------------- pass 2: acc[g][d] = sum_r p[r][g] * V[r][d] ----------------
  constexpr int TPG = D / 2;            // threads per row group (one bf16 pair each)
  constexpr int GROUPS = THREADS / TPG;  // row groups
  const int dp = tid % TPG, rg = tid / TPG;
  float acc[MAX_G][2];
#pragma unroll
  for (int g = 0; g < MAX_G; ++g) acc[g][0] = acc[g][1] = 0.f;
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) load_tile<D>(tiles + ((t + 1) & 1) * DecodeSmem<D>::kTile, vbase, sv_l, r0 + (t + 1) * TILE_ROWS, r1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint8_t* tile = tiles + (t & 1) * DecodeSmem<D>::kTile;
    const int rows_here = min(TILE_ROWS, n_rows - t * TILE_ROWS);
    for (int r = rg; r < rows_here; r += GROUPS) {
      const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(tile + r * DecodeSmem<D>::kRowBytes + dp * 4));
      const float4* pr = reinterpret_cast<const float4*>(sp + (t * TILE_ROWS + r) * MAX_G);
      const float4 pa = pr[0];
      acc[0][0] = fmaf(pa.x, vv.x, acc[0][0]), acc[0][1] = fmaf(pa.x, vv.y, acc[0][1]);
      acc[1][0] = fmaf(pa.y, vv.x, acc[1][0]), acc[1][1] = fmaf(pa.y, vv.y, acc[1][1]);
      acc[2][0] = fmaf(pa.z, vv.x, acc[2][0]), acc[2][1] = fmaf(pa.z, vv.y, acc[2][1]);
      acc[3][0] = fmaf(pa.w, vv.x, acc[3][0]), acc[3][1] = fmaf(pa.w, vv.y, acc[3][1]);
      if (G > 4) {
        const float4 pb = pr[1];
        acc[4][0] = fmaf(pb.x, vv.x, acc[4][0]), acc[4][1] = fmaf(pb.x, vv.y, acc[4][1]);
        acc[5][0] = fmaf(pb.y, vv.x, acc[5][0]), acc[5][1] = fmaf(pb.y, vv.y, acc[5][1]);
        acc[6][0] = fmaf(pb.z, vv.x, acc[6][0]), acc[6][1] = fmaf(pb.z, vv.y, acc[6][1]);
        acc[7][0] = fmaf(pb.w, vv.x, acc[7][0]), acc[7][1] = fmaf(pb.w, vv.y, acc[7][1]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();

  // reduce the row groups through shared memory (tile buffers are free now): red[GROUPS][MAX_G][D]
  float* red = reinterpret_cast<float*>(tiles);
  static_assert(GROUPS * MAX_G * D * 4 <= DecodeSmem<D>::kTiles, "reduction scratch must fit in the tile buffers");
#pragma unroll
  for (int g = 0; g < MAX_G; ++g) {
    red[(rg * MAX_G + g) * D + dp * 2] = acc[g][0];
    red[(rg * MAX_G + g) * D + dp * 2 + 1] = acc[g][1];
  }
  __syncthreads();
  // partial record per (b, hkv, split, g): [m, l, acc[D]]
  float* out = partial + ((static_cast<long long>(b) * Hkv + hkv) * nsplit + split) * G * (D + 2);
  for (int i = tid; i < G * D; i += THREADS) {
    const int g = i / D, d = i % D;
    float s = 0.f;
#pragma unroll
    for (int x = 0; x < GROUPS; ++x) s += red[(x * MAX_G + g) * D + d];
    out[g * (D + 2) + 2 + d] = s;
  }
  if (tid < G) {
    out[tid * (D + 2)] = s_m[tid];
    out[tid * (D + 2) + 1] = s_l[tid];
  }
}

template <int D>
__global__ void __launch_bounds__(D)
decode_merge_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out, int H, int Hkv, int nsplit) {
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const int G = H / Hkv, hkv = h / G, g = h % G;
  const float* base = partial + (static_cast<long long>(b) * Hkv + hkv) * nsplit * G * (D + 2) + g * (D + 2);
  const long long stride = static_cast<long long>(G) * (D + 2);
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, base[s * stride]);
  float num = 0.f, den = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float m = base[s * stride];
    const float w = m == -INFINITY ? 0.f : exp2f(m - M);
    num = fmaf(w, base[s * stride + 2 + d], num);
    den = fmaf(w, base[s * stride + 1], den);
  }
  out[(static_cast<long long>(b) * H + h) * D + d] = __float2bfloat16(num / den);
}

inline void split_plan(int B, int Hkv, int kv_len, int* nsplit, int* rows_per_split) {
  int want = (2 * 148 + B * Hkv - 1) / (B * Hkv);             // ~2 CTAs per SM in flight
  int by_len = (kv_len + 4 * TILE_ROWS - 1) / (4 * TILE_ROWS);  // at least 256 rows per CTA
  int n = want < by_len ? want : by_len;
  if (n < 1) n = 1;
  int rows = (kv_len + n - 1) / n;
  rows = (rows + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
  if (rows > MAX_CHUNK) rows = MAX_CHUNK;
  n = (kv_len + rows - 1) / rows;
  *nsplit = n;
  *rows_per_split = rows;
}

}  // namespace

extern "C" {

size_t sdag_decode_workspace_bytes(int B, int H, int D, int kv_len) {
  if (B <= 0 || H <= 0 || D <= 0 || kv_len <= 0) return 0;
  // worst case over Hkv: nsplit is bounded by kv_len / TILE_ROWS + 1
  const size_t max_split = static_cast<size_t>(kv_len) / TILE_ROWS + 1;
  return sizeof(float) * static_cast<size_t>(B) * H * max_split * (D + 2);
}

int sdag_decode_fwd(const void* q, int64_t sq_b, int64_t sq_h, const void* k, int64_t sk_b, int64_t sk_h, int64_t sk_l,
                    const void* v, int64_t sv_b, int64_t sv_h, int64_t sv_l, void* out, int B, int H, int Hkv, int kv_len,
                    int D, float scaling, void* workspace, size_t workspace_bytes, const int32_t* kv_len_dev, void* stream) {
  if (!q || !k || !v || !out || !workspace) return SDAG_ERR_INVALID;
  if (B <= 0 || H <= 0 || Hkv <= 0 || kv_len <= 0 || H % Hkv) return SDAG_ERR_INVALID;
  if (D != 64 && D != 128) return SDAG_ERR_UNSUPPORTED;
  if (H / Hkv > MAX_G) return SDAG_ERR_UNSUPPORTED;
  if ((sk_l % 8) || (sv_l % 8) || (sk_h % 8) || (sv_h % 8) || (sk_b % 8) || (sv_b % 8)) return SDAG_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) return SDAG_ERR_INVALID;
  int nsplit, rows;
  split_plan(B, Hkv, kv_len, &nsplit, &rows);
  const size_t need = sizeof(float) * static_cast<size_t>(B) * H * nsplit * (D + 2);
  if (workspace_bytes < need) return SDAG_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sl2 = scaling * 1.4426950408889634f;
  const dim3 grid(nsplit, Hkv, B);
  auto* qq = static_cast<const __nv_bfloat16*>(q);
  auto* kk = static_cast<const __nv_bfloat16*>(k);
  auto* vv = static_cast<const __nv_bfloat16*>(v);
  float* part = static_cast<float*>(workspace);
  if (D == 128) {
    static bool cfg = false;
    if (!cfg) {
      SDAG_CUDA_TRY(cudaFuncSetAttribute(decode_split_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         DecodeSmem<128>::kTotal));
      cfg = true;
    }
    decode_split_kernel<128><<<grid, THREADS, DecodeSmem<128>::kTotal, st>>>(qq, sq_b, sq_h, kk, sk_b, sk_h, sk_l, vv, sv_b,
                                                                             sv_h, sv_l, part, H, Hkv, kv_len, rows, sl2, kv_len_dev);
    SDAG_CUDA_TRY(cudaGetLastError());
    decode_merge_kernel<128><<<dim3(H, B), 128, 0, st>>>(part, static_cast<__nv_bfloat16*>(out), H, Hkv, nsplit);
  } else {
    static bool cfg = false;
    if (!cfg) {
      SDAG_CUDA_TRY(cudaFuncSetAttribute(decode_split_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         DecodeSmem<64>::kTotal));
      cfg = true;
    }
    decode_split_kernel<64><<<grid, THREADS, DecodeSmem<64>::kTotal, st>>>(qq, sq_b, sq_h, kk, sk_b, sk_h, sk_l, vv, sv_b,
                                                                           sv_h, sv_l, part, H, Hkv, kv_len, rows, sl2, kv_len_dev);
    SDAG_CUDA_TRY(cudaGetLastError());
    decode_merge_kernel<64><<<dim3(H, B), 64, 0, st>>>(part, static_cast<__nv_bfloat16*>(out), H, Hkv, nsplit);
  }
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
