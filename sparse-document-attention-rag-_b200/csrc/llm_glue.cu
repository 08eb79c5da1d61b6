// Element-wise glue of the decoder layer around the attention kernel (path A4: the prefill forward that
// generate_with_custom_mask drives, SDAG.py:175-176 -> HF LlamaDecoderLayer, models/llama/modeling_llama.py).
// HF runs these as ~25 separate torch kernels per layer (RMSNorm in fp32 with 6 passes, RoPE with slice / neg /
// cat, SiLU then mul, residual adds); here each is one HBM-bound pass with 16-byte accesses.  Roundings follow
// the HF bf16 module order so the result matches the reference forward up to GEMM summation order:
//   RMSNorm   y = w * bf16(x * rsqrt(mean(x^2) + eps))            (LlamaRMSNorm.forward)
//   RoPE      q' = bf16(bf16(q*cos) + bf16(rot(q)*sin))           (apply_rotary_pos_emb / rotate_half)
//   SwiGLU    h = bf16(bf16(silu(g)) * u)                         (LlamaMLP.forward)
#include <math.h>
#include <stdlib.h>

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
  pdl_wait();     // (decode step) launched as a programmatic dependent of the q/k/v GEMV
  pdl_trigger();  // the attention kernel that follows waits for this grid before it reads the cache
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
  static bool cfg[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (!cfg[dev] && smem > 48 * 1024) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(rmsnorm_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    cfg[dev] = true;
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
  return static_cast<int>(launch_pdl(
      rope_qkv_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(cos), static_cast<const __nv_bfloat16*>(sin),
      static_cast<__nv_bfloat16*>(q_out), static_cast<__nv_bfloat16*>(k_out), static_cast<__nv_bfloat16*>(v_out), B, L, H, Hkv,
      D, static_cast<long long>(kv_stride_b), static_cast<long long>(kv_stride_h), pos0, pos0_dev, pos0_dev_stride));
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

// ----------------------------------------------------------------------------------------------------------------
// Decode-step linear layers (path A5: one token per prompt, SDAG.py:191-211): out[b][n] = sum_k W[n][k] * xin[b][k].
// Weight rows are streamed from HBM exactly once with 16-byte loads (HBM-bound: 2*B FLOP per weight byte); the small
// input is staged in shared memory, transformed on the way in (nothing / RMSNorm / SwiGLU of a gate|up row) so the
// norm and activation launches of the layer disappear, and the residual add rides in the epilogue.  Roundings follow
// the HF bf16 module order (same helpers as the stand-alone glue kernels above).
// ----------------------------------------------------------------------------------------------------------------
namespace {

constexpr int GEMV_THREADS = 256;
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;  // volatile: the loads of one iteration stay together, ahead of the arithmetic that consumes them
  // read-only weights, never reused: no L1 allocation (with large shared-memory carve-outs L1 is only a few tens of KB and
  // would otherwise bound the bytes in flight)
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

enum GemvPrologue { PRO_NONE = 0, PRO_RMSNORM = 1, PRO_SWIGLU = 2 };

// R output rows per warp pass (the staged input chunk is reused from registers for all of them): 4 for wide layers,
// 2 for narrow ones so that every SM still gets warps to run.  `pair_swiglu`: weight rows come in (gate j, up j + N/2)
// pairs and the epilogue writes out[b][j] = silu(gate) * up — the MLP activation never exists as a separate pass.
// q/k/v projection of a decode step with the rotary embedding and the cache write in the epilogue (ROPE = true): the
// warp's rows come in (d, d + D/2) pairs of one head, so that both partners of a rotation end up in one lane.
struct RopeArgs {
  const __nv_bfloat16* cos;  // [positions, D]
  const __nv_bfloat16* sin;
  __nv_bfloat16* q_out;      // [B, 1, H, D]
  __nv_bfloat16* k_out;      // cache [B, Hkv, S_cap, D] (element strides below)
  __nv_bfloat16* v_out;
  long long kv_stride_b, kv_stride_h;
  const int32_t* pos_dev;    // position of the new token: pos_dev[b * pos_dev_stride], or pos_host when null
  int pos_dev_stride, pos_host;
  int H, Hkv, D;
};

template <int B, int R, bool ROPE>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ bias,
            const __nv_bfloat16* __restrict__ norm_w, const __nv_bfloat16* __restrict__ resid, __nv_bfloat16* __restrict__ out,
            int K, int N, int prologue, float eps, int pair_swiglu, const RopeArgs rope) {
  extern __shared__ __align__(16) uint8_t dyn[];
  __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(dyn);  // [B][K] transformed input
  __shared__ float red[B][GEMV_THREADS / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = K / 8;
  const int warps_total = gridDim.x * (GEMV_THREADS / 32);
  const int n_first = (blockIdx.x * (GEMV_THREADS / 32) + warp) * R;
  const int half = N / 2;
  // weight row of the warp's r-th row at pass base n0: plain, or interleaved (gate, up) pairs
  auto wrow_of = [&](int n0, int r) -> const uint4* {
    int row;
    if (ROPE) {  // pair index -> (head, i): rows head*D + i and head*D + i + D/2
      const int pidx = (n0 >> 1) + (r >> 1), hd = rope.D >> 1;
      row = (pidx / hd) * rope.D + pidx % hd + ((r & 1) ? hd : 0);
    } else {
      row = pair_swiglu ? ((r & 1) ? half : 0) + (n0 >> 1) + (r >> 1) : n0 + r;
    }
    return reinterpret_cast<const uint4*>(W + static_cast<long long>(row) * K);
  };

  // The first weight loads of this warp go out before the input is staged: the prologue costs a few microseconds per
  // CTA, comparable to the whole streaming time of the narrow layers; the registers ride across the barrier.
  constexpr int U = R == 2 ? 4 : 2;  // chunk groups per iteration: U * R independent 16-byte loads per lane
  const int nvec_main = nvec - nvec % (32 * U);
  uint4 wa[U][R], wb[U][R];
  if (n_first < N && lane < nvec_main) {
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < R; ++r) wa[u][r] = ld_stream(wrow_of(n_first, r) + lane + u * 32);
  }

  // Launched as a programmatic dependent: everything above only touched the weights, which no kernel of the step
  // writes; the activations are read (and `out` written) only once the preceding kernel has completed.  The trigger
  // lets the next layer's kernel take over SM slots as this grid's CTAs retire, its first weight loads already in flight.
  pdl_wait();
  pdl_trigger();

  // ---- prologue: stage the (transformed) input ----
  if (prologue == PRO_SWIGLU) {
    // x = gate|up rows [B][2K]: xin = bf16(bf16(silu(g)) * u)   (LlamaMLP.forward)
    for (int i = tid; i < B * nvec; i += GEMV_THREADS) {
      const int b = i / nvec, v = i % nvec;
      float g[8], u[8];
      unpack8(reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * 2 * K)[v], g);
      unpack8(reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * 2 * K + K)[v], u);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = bf(g[e] / (1.f + __expf(-g[e]))) * u[e];
      reinterpret_cast<uint4*>(sx + b * K)[v] = pack8(g);
    }
  } else if (prologue == PRO_RMSNORM) {
    float ss[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      ss[b] = 0.f;
      for (int v = tid; v < nvec; v += GEMV_THREADS) {
        const uint4 a = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * K)[v];
        reinterpret_cast<uint4*>(sx + b * K)[v] = a;
        float f[8];
        unpack8(a, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss[b] = fmaf(f[e], f[e], ss[b]);
      }
      for (int o = 16; o > 0; o >>= 1) ss[b] += __shfl_xor_sync(0xffffffffu, ss[b], o);
      if (lane == 0) red[b][warp] = ss[b];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < GEMV_THREADS / 32; ++w) tot += red[b][w];
      const float rstd = rsqrtf(tot / static_cast<float>(K) + eps);
      for (int v = tid; v < nvec; v += GEMV_THREADS) {  // each thread re-reads exactly the chunks it wrote
        float f[8], g[8];
        unpack8(reinterpret_cast<const uint4*>(sx + b * K)[v], f);
        unpack8(reinterpret_cast<const uint4*>(norm_w)[v], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = g[e] * bf(f[e] * rstd);  // LlamaRMSNorm.forward
        reinterpret_cast<uint4*>(sx + b * K)[v] = pack8(f);
      }
    }
  } else {
    for (int i = tid; i < B * nvec; i += GEMV_THREADS) reinterpret_cast<uint4*>(sx)[i] = reinterpret_cast<const uint4*>(x)[i];
  }
  __syncthreads();

  // ---- main loop: R weight rows per warp pass; the loads of iteration i+1 are issued before iteration i is consumed ----
  for (int n0 = n_first; n0 < N; n0 += warps_total * R) {
    float acc[R][B];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < B; ++b) acc[r][b] = 0.f;
    const uint4* wr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) wr[r] = wrow_of(n0, r);
    auto chunk_fma = [&](const uint4 (&w)[R], int v) {
#pragma unroll
      for (int b = 0; b < B; ++b) {
        float xv[8];
        unpack8(reinterpret_cast<const uint4*>(sx + b * K)[v], xv);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float wf[8];
          unpack8(w[r], wf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r][b] = fmaf(wf[e], xv[e], acc[r][b]);
        }
      }
    };
    auto load_iter = [&](uint4 (&w)[U][R], int v0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) w[u][r] = ld_stream(wr[r] + v0 + u * 32);
    };
    auto use_iter = [&](const uint4 (&w)[U][R], int v0) {
#pragma unroll
      for (int u = 0; u < U; ++u) chunk_fma(w[u], v0 + u * 32);
    };
    if (n0 != n_first && lane < nvec_main) load_iter(wa, lane);  // the first pass was loaded before the prologue
    for (int v0 = lane; v0 < nvec_main; v0 += 64 * U) {
      const bool more1 = v0 + 32 * U < nvec_main;
      if (more1) load_iter(wb, v0 + 32 * U);
      use_iter(wa, v0);
      if (more1) {
        if (v0 + 64 * U < nvec_main) load_iter(wa, v0 + 64 * U);
        use_iter(wb, v0 + 32 * U);
      }
    }
    for (int v = nvec_main + lane; v < nvec; v += 32) {  // tail when K is not a multiple of 256 * U
      uint4 w[R];
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = ld_stream(wr[r] + v);
      chunk_fma(w, v);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < B; ++b) {
        float sum = acc[r][b];
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        acc[r][b] = sum;
      }
    if (ROPE) {
      if (lane < (R / 2) * B) {
        const int p = lane / B, b = lane % B;
        float x1 = 0.f, x2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < R / 2; ++pp)
#pragma unroll
          for (int bb = 0; bb < B; ++bb)
            if (pp == p && bb == b) x1 = acc[2 * pp][bb], x2 = acc[2 * pp + 1][bb];
        const int D = rope.D, hd = D >> 1;
        const int pidx = (n0 >> 1) + p, head = pidx / hd, i = pidx % hd;
        if (bias) {
          x1 += __bfloat162float(bias[head * D + i]);
          x2 += __bfloat162float(bias[head * D + i + hd]);
        }
        x1 = bf(x1), x2 = bf(x2);  // nn.Linear output in bf16
        const int pos = rope.pos_dev ? rope.pos_dev[b * rope.pos_dev_stride] : rope.pos_host;
        __nv_bfloat16* dst;
        if (head < rope.H)
          dst = rope.q_out + (static_cast<long long>(b) * rope.H + head) * D;
        else if (head < rope.H + rope.Hkv)
          dst = rope.k_out + b * rope.kv_stride_b + (head - rope.H) * rope.kv_stride_h + static_cast<long long>(pos) * D;
        else
          dst = rope.v_out + b * rope.kv_stride_b + (head - rope.H - rope.Hkv) * rope.kv_stride_h + static_cast<long long>(pos) * D;
        float o1 = x1, o2 = x2;
        if (head < rope.H + rope.Hkv) {  // apply_rotary_pos_emb in bf16: q*cos + rotate_half(q)*sin, rotate_half = cat(-x2, x1)
          const __nv_bfloat16* cp = rope.cos + static_cast<long long>(pos) * D;
          const __nv_bfloat16* sp = rope.sin + static_cast<long long>(pos) * D;
          o1 = bf(x1 * __bfloat162float(cp[i])) + bf(-x2 * __bfloat162float(sp[i]));
          o2 = bf(x2 * __bfloat162float(cp[i + hd])) + bf(x1 * __bfloat162float(sp[i + hd]));
        }
        dst[i] = __float2bfloat16(o1);
        dst[i + hd] = __float2bfloat16(o2);
      }
    } else if (!pair_swiglu) {
      if (lane < R * B) {
        const int r = lane / B, b = lane % B;
        float sum = 0.f;
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
          for (int bb = 0; bb < B; ++bb)
            if (rr == r && bb == b) sum = acc[rr][bb];
        const int n = n0 + r;
        if (bias) sum += __bfloat162float(bias[n]);
        sum = bf(sum);                                                                          // nn.Linear output in bf16
        if (resid) sum = bf(sum + __bfloat162float(resid[static_cast<long long>(b) * N + n]));  // residual + hidden
        out[static_cast<long long>(b) * N + n] = __float2bfloat16(sum);
      }
    } else if (lane < (R / 2) * B) {
      const int p = lane / B, b = lane % B;  // pair p of this pass: rows 2p (gate) and 2p + 1 (up)
      float g = 0.f, u = 0.f;
#pragma unroll
      for (int pp = 0; pp < R / 2; ++pp)
#pragma unroll
        for (int bb = 0; bb < B; ++bb)
          if (pp == p && bb == b) g = acc[2 * pp][bb], u = acc[2 * pp + 1][bb];
      g = bf(g), u = bf(u);  // the two nn.Linear outputs in bf16, then LlamaMLP: bf16(bf16(silu(g)) * u)
      const int j = (n0 >> 1) + p;
      out[static_cast<long long>(b) * half + j] = __float2bfloat16(bf(g / (1.f + __expf(-g))) * u);
    }
  }
}

template <int B, int R, bool ROPE>
int launch_gemv_r(const void* x, const void* W, const void* bias, const void* norm_w, const void* resid, void* out, int K, int N,
                  int prologue, float eps, int pair, int n_sm, cudaStream_t st, const RopeArgs& rope) {
  const size_t smem = static_cast<size_t>(B) * K * 2;
  static size_t configured[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 && smem > configured[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(gemv_kernel<B, R, ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured[dev] = smem;
  }
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 3);
  long long ctas = (static_cast<long long>(N) / R + (GEMV_THREADS / 32) - 1) / (GEMV_THREADS / 32);
  if (ctas > static_cast<long long>(n_sm) * per_sm) ctas = static_cast<long long>(n_sm) * per_sm;
  return static_cast<int>(launch_pdl(
      gemv_kernel<B, R, ROPE>, dim3(static_cast<unsigned>(ctas)), dim3(GEMV_THREADS), smem, st, static_cast<const __nv_bfloat16*>(x),
      static_cast<const __nv_bfloat16*>(W), static_cast<const __nv_bfloat16*>(bias), static_cast<const __nv_bfloat16*>(norm_w),
      static_cast<const __nv_bfloat16*>(resid), static_cast<__nv_bfloat16*>(out), K, N, prologue, eps, pair, rope));
}

template <int B>
int launch_gemv(const void* x, const void* W, const void* bias, const void* norm_w, const void* resid, void* out, int K, int N,
                int prologue, float eps, int pair, cudaStream_t st, const RopeArgs* rope = nullptr) {
  const int n_sm = device_sm_count();
  // narrow layers: two rows per warp pass, so that N / 2 warp passes still cover every SM a few times
  const RopeArgs none = {};
  if (rope) return launch_gemv_r<B, 2, true>(x, W, bias, norm_w, resid, out, K, N, prologue, eps, 0, n_sm, st, *rope);
  if (N / 4 < n_sm * 16) return launch_gemv_r<B, 2, false>(x, W, bias, norm_w, resid, out, K, N, prologue, eps, pair, n_sm, st, none);
  return launch_gemv_r<B, 4, false>(x, W, bias, norm_w, resid, out, K, N, prologue, eps, pair, n_sm, st, none);
}

}  // namespace

// ----------------------------------------------------------------------------------------------------------------
// Tensor-core form of the same layer for up to 8 prompts: D[16 rows x 8 prompts] += W[16 x 16] * x^T[16 x 8] with warp
// MMAs, only to keep the instruction count per weight byte independent of the batch (the op stays HBM-bound).  A dot
// product does not care about the order of k, so each lane loads one contiguous 16-byte chunk of a weight row and feeds
// it to two MMAs: within a block of 32 k's, lane (g, t) owns k = 8t..8t+7 of rows g and g+8; MMA 1 takes elements
// {0,1 | 2,3} of the chunk as its (k_lo | k_hi) pair, MMA 2 elements {4,5 | 6,7}; the input chunk of prompt g is split
// the same way.  Persistent CTAs stage the input once; the 8 warps of a CTA split K and reduce through shared memory.
// ----------------------------------------------------------------------------------------------------------------
namespace {

constexpr int MMA_GEMV_WARPS = 8;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(MMA_GEMV_WARPS * 32)
gemv_mma_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ bias,
                const __nv_bfloat16* __restrict__ norm_w, const __nv_bfloat16* __restrict__ resid, __nv_bfloat16* __restrict__ out,
                int B, int K, int N, int prologue, float eps, int pair_swiglu, int rows_staged) {
  // rows_staged: B input rows in shared memory, or 0 = the input is too large to stage: it is then read from global memory
  // through L1 / L2, prologue 0 only
  constexpr int THREADS = MMA_GEMV_WARPS * 32;
  extern __shared__ __align__(16) uint8_t dyn[];
  float* sred = reinterpret_cast<float*>(dyn);                                                  // [warps][32 lanes][4]
  // [rows_staged][K + 32]: rows are 64 bytes apart modulo 128, so the 8 lanes of a load phase — (g, t) and (g + 1, t) —
  // touch 8 different 16-byte bank groups
  __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(dyn + MMA_GEMV_WARPS * 32 * 4 * sizeof(float));
  const int KS = K + 32;
  __shared__ float red[8][MMA_GEMV_WARPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nvec = K / 8;
  const int half = N / 2;
  const int n_blocks = N / 16;
  // K split: warp w owns k-blocks (32 elements) [kb0, kb1)
  const int kblocks = K / 32;
  const int kb0 = static_cast<int>(static_cast<long long>(kblocks) * warp / MMA_GEMV_WARPS);
  const int kb1 = static_cast<int>(static_cast<long long>(kblocks) * (warp + 1) / MMA_GEMV_WARPS);
  auto row_ptr = [&](int blk, int r16) -> const uint4* {  // r16 in [0, 16): rows 0-7 / 8-15 of the block
    int row;
    if (pair_swiglu) row = (r16 >= 8 ? half : 0) + blk * 8 + (r16 & 7);  // 8 gate rows + their 8 up rows
    else row = blk * 16 + r16;
    return reinterpret_cast<const uint4*>(W + static_cast<long long>(row) * K) + t;
  };

  // ---- prologue: stage the input (rows >= B stay zero), optionally RMS-normalised or as silu(gate) * up ----
  for (int i = tid; i < rows_staged * nvec; i += THREADS) {
    const int b = i / nvec, v = i % nvec;
    uint4 a = make_uint4(0u, 0u, 0u, 0u);
    if (b < B) {
      if (prologue == PRO_SWIGLU) {
        float gg[8], uu[8];
        unpack8(reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * 2 * K)[v], gg);
        unpack8(reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * 2 * K + K)[v], uu);
#pragma unroll
        for (int e = 0; e < 8; ++e) gg[e] = bf(gg[e] / (1.f + __expf(-gg[e]))) * uu[e];
        a = pack8(gg);
      } else {
        a = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * K)[v];
      }
    }
    reinterpret_cast<uint4*>(sx + b * KS)[v] = a;
  }
  if (prologue == PRO_RMSNORM) {
    __syncthreads();
    for (int b = 0; b < B; ++b) {  // every thread sums a slice of the row, one block reduction per row
      float ss = 0.f;
      for (int v = tid; v < nvec; v += THREADS) {
        float f[8];
        unpack8(reinterpret_cast<const uint4*>(sx + b * KS)[v], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
      }
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) red[b][warp] = ss;
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < MMA_GEMV_WARPS; ++w) tot += red[b][w];
      const float rstd = rsqrtf(tot / static_cast<float>(K) + eps);
      for (int v = tid; v < nvec; v += THREADS) {
        float f[8], nw[8];
        unpack8(reinterpret_cast<const uint4*>(sx + b * KS)[v], f);
        unpack8(reinterpret_cast<const uint4*>(norm_w)[v], nw);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = nw[e] * bf(f[e] * rstd);  // LlamaRMSNorm.forward
        reinterpret_cast<uint4*>(sx + b * KS)[v] = pack8(f);
      }
    }
  }
  __syncthreads();

  // prompt g's chunks of this lane: shared memory when staged, else global (nullptr: no such prompt -> zeros)
  const uint4* sxv = rows_staged ? (g < rows_staged ? reinterpret_cast<const uint4*>(sx + static_cast<long long>(g) * KS) + t : nullptr)
                                 : (g < B ? reinterpret_cast<const uint4*>(x + static_cast<long long>(g) * K) + t : nullptr);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  constexpr int UN = 4;  // k-blocks per batch of loads: 8 independent 16-byte weight loads per lane and batch
  for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const uint4* wlo = row_ptr(blk, g);
    const uint4* whi = row_ptr(blk, g + 8);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    // Register double-buffering: the loads of batch i+1 are issued before batch i is consumed, so a batch of UN * 2
    // 16-byte loads per lane is always in flight.
    uint4 la[UN], ha[UN], lb[UN], hb[UN];
    auto load_batch = [&](uint4 (&lo)[UN], uint4 (&hi)[UN], int kb) {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        lo[u] = ld_stream(wlo + (kb + u) * 4);
        hi[u] = ld_stream(whi + (kb + u) * 4);
      }
    };
    auto use_batch = [&](const uint4 (&lo)[UN], const uint4 (&hi)[UN], int kb) {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const uint4 xv = sxv ? sxv[(kb + u) * 4] : zero4;
        mma_bf16_16816(c, lo[u].x, hi[u].x, lo[u].y, hi[u].y, xv.x, xv.y);
        mma_bf16_16816(c, lo[u].z, hi[u].z, lo[u].w, hi[u].w, xv.z, xv.w);
      }
    };
    int kb = kb0;
    const int kb_main = kb0 + (kb1 - kb0) / UN * UN;
    if (kb < kb_main) load_batch(la, ha, kb);
    for (; kb < kb_main; kb += 2 * UN) {
      const bool more1 = kb + UN < kb_main;
      if (more1) load_batch(lb, hb, kb + UN);
      use_batch(la, ha, kb);
      if (more1) {
        if (kb + 2 * UN < kb_main) load_batch(la, ha, kb + 2 * UN);
        use_batch(lb, hb, kb + UN);
      }
    }
    for (kb = kb_main; kb < kb1; ++kb) {
      const uint4 lo = ld_stream(wlo + kb * 4), hi = ld_stream(whi + kb * 4), xv = sxv ? sxv[kb * 4] : zero4;
      mma_bf16_16816(c, lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
      mma_bf16_16816(c, lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
    }
    // ---- reduce the K slices of the 8 warps, then the epilogue on warp 0's fragment layout ----
    *reinterpret_cast<float4*>(sred + (warp * 32 + lane) * 4) = make_float4(c[0], c[1], c[2], c[3]);
    __syncthreads();
    if (warp == 0) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < MMA_GEMV_WARPS; ++w) {
        const float4 p4 = *reinterpret_cast<const float4*>(sred + (w * 32 + lane) * 4);
        acc.x += p4.x, acc.y += p4.y, acc.z += p4.z, acc.w += p4.w;
      }
      // fragment: (row g, prompts 2t, 2t+1) = acc.x, acc.y; (row g + 8, same prompts) = acc.z, acc.w
      const float v[2][2] = {{acc.x, acc.y}, {acc.z, acc.w}};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int b = t * 2 + e;
        if (b >= B) continue;
        if (pair_swiglu) {
          const float gate = bf(v[0][e]), up = bf(v[1][e]);  // the two nn.Linear outputs in bf16, then LlamaMLP
          out[static_cast<long long>(b) * half + blk * 8 + g] = __float2bfloat16(bf(gate / (1.f + __expf(-gate))) * up);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = blk * 16 + g + h * 8;
            float sum = v[h][e];
            if (bias) sum += __bfloat162float(bias[n]);
            sum = bf(sum);
            if (resid) sum = bf(sum + __bfloat162float(resid[static_cast<long long>(b) * N + n]));
            out[static_cast<long long>(b) * N + n] = __float2bfloat16(sum);
          }
        }
      }
    }
    __syncthreads();
  }
}

int launch_gemv_mma(const void* x, const void* W, const void* bias, const void* norm_w, const void* resid, void* out, int B, int K,
                    int N, int prologue, float eps, int pair, cudaStream_t st) {
  const size_t red_bytes = MMA_GEMV_WARPS * 32 * 4 * sizeof(float);
  int rows_staged = B;
  if (static_cast<size_t>(rows_staged) * (K + 32) * 2 + red_bytes > 72 * 1024) {  // keep three CTAs per SM
    if (prologue != PRO_NONE && static_cast<size_t>(rows_staged) * (K + 32) * 2 + red_bytes > 200 * 1024) return SDAG_ERR_UNSUPPORTED;
    if (prologue == PRO_NONE) rows_staged = 0;
  }
  const size_t smem = static_cast<size_t>(rows_staged) * (K + 32) * 2 + red_bytes;
  static size_t configured[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 && smem > configured[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(gemv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured[dev] = smem;
  }
  const int n_sm = device_sm_count();
  const int per_sm = smem > 110 * 1024 ? 1 : (smem > 72 * 1024 ? 2 : 3);
  long long ctas = N / 16;
  if (ctas > static_cast<long long>(n_sm) * per_sm) ctas = static_cast<long long>(n_sm) * per_sm;
  gemv_mma_kernel<<<static_cast<unsigned>(ctas), MMA_GEMV_WARPS * 32, smem, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(W), static_cast<const __nv_bfloat16*>(bias),
      static_cast<const __nv_bfloat16*>(norm_w), static_cast<const __nv_bfloat16*>(resid), static_cast<__nv_bfloat16*>(out), B, K, N,
      prologue, eps, pair, rows_staged);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" int llm_gemv(const void* x, const void* weight, const void* bias, const void* norm_weight, const void* residual,
                        void* out, int B, int K, int N, int prologue, float eps, int pair_swiglu, void* stream) {
  if (!x || !weight || !out || B <= 0 || K <= 0 || N <= 0 || (K % 8) || (N % 4)) return SDAG_ERR_INVALID;
  if (prologue < PRO_NONE || prologue > PRO_SWIGLU || (prologue == PRO_RMSNORM && !norm_weight)) return SDAG_ERR_INVALID;
  if (pair_swiglu && (bias || residual)) return SDAG_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pair = pair_swiglu ? 1 : 0;
  // tensor-core form: any B <= 8 when the row blocks line up and the staged input (8 rows) fits in shared memory
  const bool mma_ok = B <= 8 && (K % 32) == 0 && (N % 16) == 0 &&
                      (prologue == PRO_NONE || static_cast<size_t>(B) * (K + 32) * 2 + 4096 <= 200 * 1024);
#ifdef SDAG_GEMV_FORCE_MMA
  if (mma_ok) return launch_gemv_mma(x, weight, bias, norm_weight, residual, out, B, K, N, prologue, eps, pair, st);
#endif
  // the FMA form keeps up with the stream for one or two prompts; SDAG_GEMV_FMA_MAX_B (3 or 4) widens it for experiments
  static int fma_max_b = -1;
  if (fma_max_b < 0) {
    const char* e = getenv("SDAG_GEMV_FMA_MAX_B");
    fma_max_b = e ? atoi(e) : 2;
    if (fma_max_b < 2 || fma_max_b > 4) fma_max_b = 2;
  }
  if (B > fma_max_b && mma_ok) return launch_gemv_mma(x, weight, bias, norm_weight, residual, out, B, K, N, prologue, eps, pair, st);
  if (B > 4 || static_cast<size_t>(B) * K * 2 > 200 * 1024) return SDAG_ERR_UNSUPPORTED;
  switch (B) {
    case 1: return launch_gemv<1>(x, weight, bias, norm_weight, residual, out, K, N, prologue, eps, pair, st);
    case 2: return launch_gemv<2>(x, weight, bias, norm_weight, residual, out, K, N, prologue, eps, pair, st);
    case 3: return launch_gemv<3>(x, weight, bias, norm_weight, residual, out, K, N, prologue, eps, pair, st);
    default: return launch_gemv<4>(x, weight, bias, norm_weight, residual, out, K, N, prologue, eps, pair, st);
  }
}

extern "C" int llm_gemv_rope(const void* x, const void* weight, const void* bias, const void* norm_weight, float eps,
                             const void* cos, const void* sin, void* q_out, void* k_out, void* v_out, int B, int K, int H,
                             int Hkv, int D, int64_t kv_stride_b, int64_t kv_stride_h, int pos0, const int32_t* pos0_dev,
                             int pos0_dev_stride, void* stream) {
  if (!x || !weight || !cos || !sin || !q_out || !k_out || !v_out) return SDAG_ERR_INVALID;
  if (B <= 0 || K <= 0 || (K % 8) || H <= 0 || Hkv <= 0 || D <= 0 || (D % 2)) return SDAG_ERR_INVALID;
  if (pos0_dev_stride != 0 && pos0_dev_stride != 1) return SDAG_ERR_INVALID;
  if (B > 4 || static_cast<size_t>(B) * K * 2 > 200 * 1024) return SDAG_ERR_UNSUPPORTED;
  RopeArgs r;
  r.cos = static_cast<const __nv_bfloat16*>(cos);
  r.sin = static_cast<const __nv_bfloat16*>(sin);
  r.q_out = static_cast<__nv_bfloat16*>(q_out);
  r.k_out = static_cast<__nv_bfloat16*>(k_out);
  r.v_out = static_cast<__nv_bfloat16*>(v_out);
  r.kv_stride_b = kv_stride_b;
  r.kv_stride_h = kv_stride_h;
  r.pos_dev = pos0_dev;
  r.pos_dev_stride = pos0_dev_stride;
  r.pos_host = pos0;
  r.H = H;
  r.Hkv = Hkv;
  r.D = D;
  const int N = (H + 2 * Hkv) * D;
  const int prologue = norm_weight ? PRO_RMSNORM : PRO_NONE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (B) {
    case 1: return launch_gemv<1>(x, weight, bias, norm_weight, nullptr, nullptr, K, N, prologue, eps, 0, st, &r);
    case 2: return launch_gemv<2>(x, weight, bias, norm_weight, nullptr, nullptr, K, N, prologue, eps, 0, st, &r);
    case 3: return launch_gemv<3>(x, weight, bias, norm_weight, nullptr, nullptr, K, N, prologue, eps, 0, st, &r);
    default: return launch_gemv<4>(x, weight, bias, norm_weight, nullptr, nullptr, K, N, prologue, eps, 0, st, &r);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// L2 prefetch of a weight matrix: launched on a side stream while the decode step's attention kernels run (HBM is
// nearly idle then), so that the o_proj GEMV that follows streams its 33 MB of weights out of the 126 MB L2.
// ----------------------------------------------------------------------------------------------------------------
namespace {
constexpr unsigned PREFETCH_CHUNK = 16 * 1024;
__global__ void __launch_bounds__(128) prefetch_l2_kernel(const uint8_t* __restrict__ base, unsigned long long bytes) {
  const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  const unsigned long long off = i * PREFETCH_CHUNK;
  if (off >= bytes) return;
  const unsigned long long left = bytes - off;
  const unsigned size = static_cast<unsigned>(left < PREFETCH_CHUNK ? (left & ~15ull) : PREFETCH_CHUNK);
  if (size) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(size) : "memory");
}
}  // namespace

extern "C" int llm_prefetch_l2(const void* ptr, size_t bytes, void* stream) {
  if (!ptr || (reinterpret_cast<uintptr_t>(ptr) & 15)) return SDAG_ERR_INVALID;
  if (bytes == 0) return SDAG_OK;
  const unsigned long long chunks = (bytes + PREFETCH_CHUNK - 1) / PREFETCH_CHUNK;
  prefetch_l2_kernel<<<static_cast<unsigned>((chunks + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(ptr), bytes);
  return static_cast<int>(cudaGetLastError());
}
