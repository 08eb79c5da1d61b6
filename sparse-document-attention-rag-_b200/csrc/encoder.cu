// Glue of the passage / query encoder around the attention kernel (path B1 / B6: the BERT-large forward behind
// SentenceTransformer.encode, reached from embed_queries, dense.py:44-65, encode_texts_with_ranker,
// malicious_selection.py:12-34, and compute_doc_knn_for_docs, SDAG.py:46-48).  The encoder is bidirectional, so its
// attention is the prefill kernel with a block map in which every tile of a sequence is fully visible; the linear
// layers are cuBLAS.  What is left between them is HBM-bound row work, one pass each here:
//   LayerNorm   y = bf16(w * (s - mean) * rstd + b),  s = bf16(x + residual)     (BertSelfOutput / BertOutput)
//   GELU        y = bf16(0.5 x (1 + erf(x / sqrt 2)))                             (BertIntermediate, erf form; erf to 1.5e-7)
//   pooling     e = sum_{t < len} h_t / len;  e / max(|e|, 1e-12)                (Pooling(mean) + Normalize)
// The pooled vector is written as fp32 (and, optionally, its bf16 shadow) straight into the rows of the HBM-resident
// index shard the search kernels read.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ void unpack8(const uint4& w, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 w;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return w;
}

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) tot += red[i];
  return tot;
}

// One warp per row, the (summed) row held in registers between the statistics passes and the output pass: no shared
// memory, no block barrier.  VPL = 16-byte vectors per lane (dim = 256 * VPL).
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_warp_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                      const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out,
                      long long rows, float eps) {
  constexpr int dim = 256 * VPL;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  float f[VPL][8];
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    unpack8(reinterpret_cast<const uint4*>(x + row * dim)[j * 32 + lane], f[j]);
    if (res) {
      float g[8];
      unpack8(reinterpret_cast<const uint4*>(res + row * dim)[j * 32 + lane], g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[j][i] = bf(f[j][i] + g[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s1 += f[j][i];
  }
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  const float mean = s1 / static_cast<float>(dim);
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) s2 = fmaf(f[j][i] - mean, f[j][i] - mean, s2);
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  const float rstd = rsqrtf(s2 / static_cast<float>(dim) + eps);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    float g[8], h[8];
    unpack8(reinterpret_cast<const uint4*>(w)[j * 32 + lane], g);
    unpack8(reinterpret_cast<const uint4*>(b)[j * 32 + lane], h);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[j][i] = fmaf((f[j][i] - mean) * rstd, g[i], h[i]);
    reinterpret_cast<uint4*>(out + row * dim)[j * 32 + lane] = pack8(f[j]);
  }
}

// General widths: one CTA per row; the (summed) row stays in shared memory between the statistics pass and the output pass.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                 const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out,
                 int dim, float eps) {
  const long long row = blockIdx.x;
  const int nvec = dim / 8;
  extern __shared__ __align__(16) uint8_t dyn[];
  uint4* srow = reinterpret_cast<uint4*>(dyn);
  __shared__ float red[THREADS / 32];
  float s1 = 0.f;
  for (int v = threadIdx.x; v < nvec; v += THREADS) {
    uint4 a = reinterpret_cast<const uint4*>(x + row * dim)[v];
    float f[8];
    unpack8(a, f);
    if (res) {
      float g[8];
      unpack8(reinterpret_cast<const uint4*>(res + row * dim)[v], g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = bf(f[i] + g[i]);
      a = pack8(f);
    }
    srow[v] = a;
#pragma unroll
    for (int i = 0; i < 8; ++i) s1 += f[i];
  }
  const float mean = block_sum<THREADS>(s1, red) / static_cast<float>(dim);
  float s2 = 0.f;
  for (int v = threadIdx.x; v < nvec; v += THREADS) {  // two-pass variance: no cancellation
    float f[8];
    unpack8(srow[v], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) s2 = fmaf(f[i] - mean, f[i] - mean, s2);
  }
  const float rstd = rsqrtf(block_sum<THREADS>(s2, red) / static_cast<float>(dim) + eps);
  for (int v = threadIdx.x; v < nvec; v += THREADS) {
    float f[8], g[8], h[8];
    unpack8(srow[v], f);
    unpack8(reinterpret_cast<const uint4*>(w)[v], g);
    unpack8(reinterpret_cast<const uint4*>(b)[v], h);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf((f[i] - mean) * rstd, g[i], h[i]);
    reinterpret_cast<uint4*>(out + row * dim)[v] = pack8(f);
  }
}

// gelu_erf: common.cuh (shared with the GEMM epilogue of llm_gemm)
__global__ void __launch_bounds__(256) gelu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                   long long nvec) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(x)[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = gelu_erf(f[e]);
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}

// One CTA per sequence: thread t owns 8 consecutive channels (dim <= 8 * THREADS); rows are read with 16-byte loads,
// coalesced across the CTA.  Output row = row0 + sequence index.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
mean_pool_normalize_kernel(const __nv_bfloat16* __restrict__ h, long long s_b, long long s_l, const int32_t* __restrict__ lengths,
                           int L, int dim, const int64_t* __restrict__ dst_rows, long long row0, float* __restrict__ out_f32,
                           __nv_bfloat16* __restrict__ out_bf16, int normalize) {
  const int seq = blockIdx.x;
  const int n = min(max(lengths[seq], 0), L);
  const int v = threadIdx.x;
  const bool live = v < dim / 8;
  __shared__ float red[THREADS / 32];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    const __nv_bfloat16* base = h + seq * s_b + v * 8;
    for (int t = 0; t < n; ++t) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base + t * s_l), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
    const float inv = 1.f / fmaxf(static_cast<float>(n), 1e-9f);  // Pooling: sum / clamp(mask.sum(), min=1e-9)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss = fmaf(acc[i], acc[i], ss);
  const float norm = sqrtf(block_sum<THREADS>(ss, red));
  const float scale = normalize ? 1.f / fmaxf(norm, 1e-12f) : 1.f;  // F.normalize(p=2, dim=1, eps=1e-12)
  if (live) {
    const long long r = dst_rows ? dst_rows[seq] : row0 + seq;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= scale;
    float4* o = reinterpret_cast<float4*>(out_f32 + r * dim + v * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (out_bf16) reinterpret_cast<uint4*>(out_bf16 + r * dim)[v] = pack8(acc);
  }
}

// Block maps of a batch of bidirectional sequences: T = ceil(n / 128) query tiles, every one of them sees the same list
// of T fully visible kv tiles (doff = 2^20: no diagonal), pairs = n^2.  Same record layout as sdag_blockmap.cu.
__global__ void full_blockmap_kernel(const int32_t* __restrict__ lengths, int B, int max_L, int cap_q, int32_t* __restrict__ out,
                                     long long stride_ints) {
  const int b = blockIdx.x;
  int32_t* map = out + b * stride_ints;
  const int n = min(max(lengths[b], 0), max_L);
  const int T = (n + SDAG_TILE - 1) / SDAG_TILE;
  int4* qtiles = reinterpret_cast<int4*>(map + SDAG_BLOCKMAP_HEADER);
  int4* kvtiles = qtiles + cap_q;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int r0 = t * SDAG_TILE, cnt = min(SDAG_TILE, n - r0);
    qtiles[t] = make_int4(r0, cnt, 0, T);              // q0, nrows, kv_off, kv_cnt
    kvtiles[t] = make_int4(r0, cnt, 1 << 20, 0);       // k0, ncols, doff, flags
  }
  if (threadIdx.x == 0) {
    const long long pairs = static_cast<long long>(n) * n;
    map[0] = T, map[1] = T, map[2] = 0, map[3] = 0, map[4] = n, map[5] = cap_q, map[6] = cap_q, map[7] = 0;
    map[8] = static_cast<int>(pairs & 0xffffffffll), map[9] = static_cast<int>(pairs >> 32), map[10] = 0;
    for (int i = 11; i < SDAG_BLOCKMAP_HEADER; ++i) map[i] = 0;
  }
}

}  // namespace

extern "C" {

int enc_layernorm(const void* x, const void* residual, const void* weight, const void* bias, void* out, int64_t rows, int dim,
                  float eps, void* stream) {
  if (!x || !weight || !bias || !out || rows < 0 || dim <= 0 || (dim % 8)) return SDAG_ERR_INVALID;
  if (dim * 2 > 48 * 1024) return SDAG_ERR_UNSUPPORTED;
  if (rows == 0) return SDAG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto *xp = static_cast<const __nv_bfloat16*>(x), *rp = static_cast<const __nv_bfloat16*>(residual),
       *wp = static_cast<const __nv_bfloat16*>(weight), *bp = static_cast<const __nv_bfloat16*>(bias);
  auto* op = static_cast<__nv_bfloat16*>(out);
  const unsigned wgrid = static_cast<unsigned>((rows + 7) / 8);
  if (dim == 1024) {  // BERT-large / e5-large-v2
    layernorm_warp_kernel<4><<<wgrid, 256, 0, st>>>(xp, rp, wp, bp, op, rows, eps);
    return static_cast<int>(cudaGetLastError());
  }
  if (dim == 768) {
    layernorm_warp_kernel<3><<<wgrid, 256, 0, st>>>(xp, rp, wp, bp, op, rows, eps);
    return static_cast<int>(cudaGetLastError());
  }
  if (dim == 256) {
    layernorm_warp_kernel<1><<<wgrid, 256, 0, st>>>(xp, rp, wp, bp, op, rows, eps);
    return static_cast<int>(cudaGetLastError());
  }
  constexpr int THREADS = 128;
  layernorm_kernel<THREADS><<<static_cast<unsigned>(rows), THREADS, static_cast<size_t>(dim) * 2, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual), static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(out), dim, eps);
  return static_cast<int>(cudaGetLastError());
}

int enc_gelu(const void* x, void* out, int64_t n_elems, void* stream) {
  if (!x || !out || n_elems < 0 || (n_elems % 8)) return SDAG_ERR_INVALID;
  if (n_elems == 0) return SDAG_OK;
  const long long nvec = n_elems / 8;
  long long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gelu_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), nvec);
  return static_cast<int>(cudaGetLastError());
}

int enc_mean_pool_normalize(const void* hidden, int64_t stride_b, int64_t stride_l, const int32_t* lengths, int B, int L, int dim,
                            const int64_t* dst_rows, int64_t row0, float* out_f32, void* out_bf16, int normalize, void* stream) {
  if (!hidden || !lengths || !out_f32 || B < 0 || L <= 0 || dim <= 0 || (dim % 8) || row0 < 0) return SDAG_ERR_INVALID;
  if ((stride_b % 8) || (stride_l % 8)) return SDAG_ERR_INVALID;
  if (dim > 8 * 256) return SDAG_ERR_UNSUPPORTED;
  if (B == 0) return SDAG_OK;
  constexpr int THREADS = 256;
  mean_pool_normalize_kernel<THREADS><<<static_cast<unsigned>(B), THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(hidden), stride_b, stride_l, lengths, L, dim, dst_rows, row0, out_f32,
      static_cast<__nv_bfloat16*>(out_bf16), normalize);
  return static_cast<int>(cudaGetLastError());
}

size_t sdag_full_blockmap_bytes(int max_L) {
  if (max_L <= 0) return 0;
  const size_t T = (static_cast<size_t>(max_L) + SDAG_TILE - 1) / SDAG_TILE;
  return sizeof(int32_t) * (SDAG_BLOCKMAP_HEADER + 4 * 2 * T);
}

int sdag_build_blockmap_full(const int32_t* lengths_dev, int B, int max_L, int32_t* out, size_t stride_bytes, void* stream) {
  if (!lengths_dev || !out || B <= 0 || max_L <= 0) return SDAG_ERR_INVALID;
  if (stride_bytes < sdag_full_blockmap_bytes(max_L) || (stride_bytes % 16)) return SDAG_ERR_WORKSPACE;
  const int cap_q = (max_L + SDAG_TILE - 1) / SDAG_TILE;
  full_blockmap_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(lengths_dev, B, max_L, cap_q, out,
                                                                         static_cast<long long>(stride_bytes / 4));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
