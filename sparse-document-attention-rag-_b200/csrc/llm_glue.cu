// Element-wise glue of the decoder layer around the attention kernel (path A4: the prefill forward that
// generate_with_custom_mask drives, SDAG.py:175-176 -> HF LlamaDecoderLayer, models/llama/modeling_llama.py).
// HF runs these as ~25 separate torch kernels per layer (RMSNorm in fp32 with 6 passes, RoPE with slice / neg /
// cat, SiLU then mul, residual adds); here each is one HBM-bound pass with 16-byte accesses.  Roundings follow
// the HF bf16 module order so the result matches the reference forward up to GEMM summation order:
//   RMSNorm   y = w * bf16(x * rsqrt(mean(x^2) + eps))            (LlamaRMSNorm.forward)
//   RoPE      q' = bf16(bf16(q*cos) + bf16(rot(q)*sin))           (apply_rotary_pos_emb / rotate_half)
//   SwiGLU    h = bf16(bf16(silu(g)) * u)                         (LlamaMLP.forward)
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

// One CTA per row.  Optionally adds a residual first: s = x + res (bf16 add), writes s to sum_out, normalises s.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
               const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ sum_out, __nv_bfloat16* __restrict__ out,
               int dim, float eps) {
  const long long row = blockIdx.x;
  const int nvec = dim / 8;
  extern __shared__ __align__(16) uint8_t dyn[];
  uint4* srow = reinterpret_cast<uint4*>(dyn);  // the (summed) row, kept for the second pass
  __shared__ float red[THREADS / 32];
  float ss = 0.f;
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
      reinterpret_cast<uint4*>(sum_out + row * dim)[v] = a;
    }
    srow[v] = a;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / static_cast<float>(dim) + eps);
  for (int v = threadIdx.x; v < nvec; v += THREADS) {
    float f[8], g[8];
    unpack8(srow[v], f);
    unpack8(reinterpret_cast<const uint4*>(w)[v], g);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = g[i] * bf(f[i] * rstd);
    reinterpret_cast<uint4*>(out + row * dim)[v] = pack8(f);
  }
}

// qkv [T, (H + 2*Hkv) * D] (row = token t = b * L + l) -> q_out [B, L, H, D] (RoPE applied), k_out / v_out
// [B, Hkv, S_cap, D] written at positions pos0 + l (RoPE on k).  cos / sin: bf16 tables indexed by ABSOLUTE position
// (row pos0 + l).  pos0 may come from device memory (pos0_dev) so a captured CUDA graph can be replayed per token;
// pos0_dev_stride = 1 gives every prompt of a ragged batch its own position (0: one value shared by the batch).
// One thread per 8 consecutive elements of the first half of a head; it also owns the partner elements d + D/2.
__global__ void __launch_bounds__(256)
rope_qkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ cosb,
                const __nv_bfloat16* __restrict__ sinb, __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_out,
                __nv_bfloat16* __restrict__ v_out, int B, int L, int H, int Hkv, int D, long long kv_stride_b,
                long long kv_stride_h, int pos0_host, const int32_t* __restrict__ pos0_dev, int pos0_dev_stride) {
  const int heads = H + 2 * Hkv;
  const int vec_per_head = D / 16;  // threads per head (each handles 8 + 8 elements)
  const long long total = static_cast<long long>(B) * L * heads * vec_per_head;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vh = static_cast<int>(i % vec_per_head);
    const long long r = i / vec_per_head;
    const int head = static_cast<int>(r % heads);
    const long long t = r / heads;
    const int l = static_cast<int>(t % L), b = static_cast<int>(t / L);
    const int pos0 = pos0_dev ? pos0_dev[b * pos0_dev_stride] : pos0_host;
    const __nv_bfloat16* src = qkv + (t * heads + head) * D;
    const uint4 lo = reinterpret_cast<const uint4*>(src)[vh];
    const uint4 hi = reinterpret_cast<const uint4*>(src + D / 2)[vh];
    __nv_bfloat16* dst;
    if (head < H) {
      dst = q_out + ((static_cast<long long>(b) * L + l) * H + head) * D;
    } else if (head < H + Hkv) {
      dst = k_out + b * kv_stride_b + (head - H) * kv_stride_h + static_cast<long long>(pos0 + l) * D;
    } else {
      dst = v_out + b * kv_stride_b + (head - H - Hkv) * kv_stride_h + static_cast<long long>(pos0 + l) * D;
      reinterpret_cast<uint4*>(dst)[vh] = lo;
      reinterpret_cast<uint4*>(dst + D / 2)[vh] = hi;
      continue;
    }
    float a[8], c[8], cl[8], ch[8], sl[8], sh[8];
    unpack8(lo, a);
    unpack8(hi, c);
    unpack8(reinterpret_cast<const uint4*>(cosb + static_cast<long long>(pos0 + l) * D)[vh], cl);
    unpack8(reinterpret_cast<const uint4*>(cosb + static_cast<long long>(pos0 + l) * D + D / 2)[vh], ch);
    unpack8(reinterpret_cast<const uint4*>(sinb + static_cast<long long>(pos0 + l) * D)[vh], sl);
    unpack8(reinterpret_cast<const uint4*>(sinb + static_cast<long long>(pos0 + l) * D + D / 2)[vh], sh);
    float o1[8], o2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      // rotate_half(x) = cat(-x[D/2:], x[:D/2])
      o1[e] = bf(a[e] * cl[e]) + bf(-c[e] * sl[e]);
      o2[e] = bf(c[e] * ch[e]) + bf(a[e] * sh[e]);
    }
    reinterpret_cast<uint4*>(dst)[vh] = pack8(o1);
    reinterpret_cast<uint4*>(dst + D / 2)[vh] = pack8(o2);
  }
}

// gate_up [T, 2*I] (gate columns first) -> out [T, I] = bf16(bf16(silu(g)) * u)
__global__ void __launch_bounds__(256)
swiglu_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ out, long long T, int I) {
  const int nvec = I / 8;
  const long long total = T * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / nvec;
    const int v = static_cast<int>(i % nvec);
    float g[8], u[8];
    unpack8(reinterpret_cast<const uint4*>(gu + t * 2 * I)[v], g);
    unpack8(reinterpret_cast<const uint4*>(gu + t * 2 * I + I)[v], u);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = bf(g[e] / (1.f + __expf(-g[e]))) * u[e];
    reinterpret_cast<uint4*>(out + t * I)[v] = pack8(g);
  }
}

}  // namespace

extern "C" {

int llm_rmsnorm(const void* x, const void* residual, const void* weight, void* sum_out, void* out, int64_t rows, int dim,
                float eps, void* stream) {
  if (!x || !weight || !out || rows < 0 || dim <= 0 || (dim % 8)) return SDAG_ERR_INVALID;
  if (residual && !sum_out) return SDAG_ERR_INVALID;
  if (dim * 2 > 64 * 1024) return SDAG_ERR_UNSUPPORTED;
  if (rows == 0) return SDAG_OK;
  constexpr int THREADS = 256;
  const size_t smem = static_cast<size_t>(dim) * 2;
  static bool cfg = false;
  if (!cfg && smem > 48 * 1024) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(rmsnorm_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    cfg = true;
  }
  rmsnorm_kernel<THREADS><<<static_cast<unsigned>(rows), THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual),
      static_cast<const __nv_bfloat16*>(weight), static_cast<__nv_bfloat16*>(sum_out), static_cast<__nv_bfloat16*>(out), dim,
      eps);
  return static_cast<int>(cudaGetLastError());
}

int llm_rope_qkv(const void* qkv, const void* cos, const void* sin, void* q_out, void* k_out, void* v_out, int B, int L,
                 int H, int Hkv, int D, int64_t kv_stride_b, int64_t kv_stride_h, int pos0, const int32_t* pos0_dev,
                 int pos0_dev_stride, void* stream) {
  if (!qkv || !cos || !sin || !q_out || !k_out || !v_out) return SDAG_ERR_INVALID;
  if (B <= 0 || L <= 0 || H <= 0 || Hkv <= 0 || D <= 0 || (D % 16) || pos0 < 0) return SDAG_ERR_INVALID;
  if (pos0_dev_stride != 0 && pos0_dev_stride != 1) return SDAG_ERR_INVALID;
  const long long total = static_cast<long long>(B) * L * (H + 2 * Hkv) * (D / 16);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  rope_qkv_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(cos), static_cast<const __nv_bfloat16*>(sin),
      static_cast<__nv_bfloat16*>(q_out), static_cast<__nv_bfloat16*>(k_out), static_cast<__nv_bfloat16*>(v_out), B, L, H, Hkv,
      D, kv_stride_b, kv_stride_h, pos0, pos0_dev, pos0_dev_stride);
  return static_cast<int>(cudaGetLastError());
}

int llm_swiglu(const void* gate_up, void* out, int64_t rows, int inter, void* stream) {
  if (!gate_up || !out || rows < 0 || inter <= 0 || (inter % 8)) return SDAG_ERR_INVALID;
  if (rows == 0) return SDAG_OK;
  const long long total = rows * (inter / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  swiglu_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(gate_up), static_cast<__nv_bfloat16*>(out), rows, inter);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
