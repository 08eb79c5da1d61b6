// Prefill linear layers of the decoder (path A4): out[M, N] = x[M, K] . W[N, K]^T in bf16 with fp32 accumulation — the
// q/k/v, o, gate/up and down projections HF runs inside generate_with_custom_mask's prefill forward (SDAG.py:175-176 ->
// LlamaAttention / LlamaMLP, models/llama/modeling_llama.py).  Written for CTA pairs: one tcgen05.mma.cta_group::2 covers
// 256 token rows x 256 weight rows; each CTA of the pair keeps its own 128 token rows (A) and its own 128 of the 256
// weight rows (B) per 64-wide k-block in shared memory, the fp32 accumulators (2 x 256 columns, double buffered) live in
// each CTA's TMEM.  The epilogue converts straight out of TMEM:
//   LLM_GEMM_PLAIN   out = bf16(acc + bias)
//   LLM_GEMM_GELU    out = bf16(gelu_erf(bf16(acc + bias)))   (the encoder's BertIntermediate)
//   LLM_GEMM_SWIGLU  W = [gate; up] rows (N = 2*I): every MMA holds 128 gate rows and the 128 matching up rows, and
//                    out[M, I] = bf16(bf16(silu(g)) * u) with g, u rounded to bf16 first — the roundings of LlamaMLP with
//                    a bf16 gate_proj / up_proj — so the [M, 2*I] intermediate never reaches HBM.
// Same pipeline as the retrieval GEMM (ip_search.cu): warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA),
// warps 2-5 epilogue; the leader's producer announces the bytes of both CTAs so the stage loop has no remote arrive.
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int TMH = 128;   // token rows per CTA (the pair's MMA has M = 256)
constexpr int TNP = 256;   // MMA N: weight rows per pair tile, 128 per CTA
constexpr int KB = 64;     // k-block: one 128-byte swizzle row of bf16
constexpr int STAGES = 6;
constexpr int A_BYTES = TMH * KB * 2;
constexpr int BQ_BYTES = 64 * KB * 2;  // one 64-row box of weight rows; a CTA loads two per k-block
constexpr int STAGE_BYTES = A_BYTES + 2 * BQ_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int GROUP_N = 16;  // most weight tiles per raster group (one wave of 74 pairs then touches 16 weight and ~5 token tiles)

struct Params {
  const __nv_bfloat16* bias;  // [N] or null
  const __nv_bfloat16* resid; // [M, ldr] or null (plain epilogue): out = bf16(bf16(acc + bias) + resid); may alias out
  long long ldr;
  __nv_bfloat16* out;
  long long ldo;  // elements
  long long M;
  int n_out;      // output columns (N, or I for SwiGLU)
  int n_mt, n_nt; // pair tiles along M (256 rows) and N
  int kblocks;
  int up_offset;  // SwiGLU: first up row (= I)
  int group_n;    // weight tiles per raster group
  uint64_t hint_x, hint_w;  // L2 eviction priorities of the two operand streams
  // LLM_GEMM_ROPE: weight = fused q|k|v rows, token row m = (prompt m / L, position pos0 + m % L)
  const __nv_bfloat16* cos;  // [positions, D]
  const __nv_bfloat16* sin;
  __nv_bfloat16* q_out;      // [B, L, H, D]
  __nv_bfloat16* k_out;      // cache [B, Hkv, S_cap, D], element strides kv_stride_b / kv_stride_h
  __nv_bfloat16* v_out;
  long long kv_stride_b, kv_stride_h;
  int L, H, Hkv, D, pos0;
};
constexpr int LLM_GEMM_ROPE_EPI = 100;  // internal epilogue id of llm_gemm_rope

struct Unit {
  int mt, nt;
};
__device__ __forceinline__ Unit unit_of(long long u, const Params& p) {
  const long long per_group = static_cast<long long>(p.group_n) * p.n_mt;
  const int g = static_cast<int>(u / per_group);
  const int rem = static_cast<int>(u - g * per_group);
  const int gn = min(p.group_n, p.n_nt - g * p.group_n);
  return Unit{rem / gn, g * p.group_n + rem % gn};
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// plain epilogues: the linear output in bf16, for LLM_GEMM_GELU followed by the erf-GELU of that bf16 value (BertIntermediate)
template <int EPI>
__device__ __forceinline__ uint32_t pack_out(float a, float b) {
  if (EPI == LLM_GEMM_GELU) return pack_bf16x2(gelu_erf(bf_round(a)), gelu_erf(bf_round(b)));
  return pack_bf16x2(a, b);
}

// The residual row segment a thread will add to its accumulator row is known before the accumulators are: pull its 128-byte
// lines into L2 while the epilogue warp would otherwise only wait for the tile's last MMA.
__device__ __forceinline__ void prefetch_residual(const __nv_bfloat16* rrow, int col0, int n_cols, int n_out) {
  for (int c = 0; c < n_cols; c += 64)
    if (col0 + c < n_out) asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + col0 + c));
}
// 32 residual values (64 bytes) of this thread's row as four 16-byte loads, issued ahead of the accumulator chunk they meet
__device__ __forceinline__ void load_residual(uint4 (&w)[4], const __nv_bfloat16* rrow, int col0, int n_out) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = (col0 + 8 * i + 8 <= n_out) ? __ldg(reinterpret_cast<const uint4*>(rrow + col0 + 8 * i)) : make_uint4(0u, 0u, 0u, 0u);
}
// v <- bf16(bf16(v) + residual): the two roundings of `residual + linear(x)` with a bf16 linear output (LlamaDecoderLayer,
// BertSelfOutput / BertOutput); columns past n_out (ragged last tile, N % 8 == 0) are left alone
__device__ __forceinline__ void apply_residual(float (&v)[32], const uint4 (&w)[4], int col0, int n_out) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (col0 + 8 * i + 8 <= n_out) {
      const uint32_t ww[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[j]));
        v[8 * i + 2 * j] = bf_round(bf_round(v[8 * i + 2 * j]) + f.x);
        v[8 * i + 2 * j + 1] = bf_round(bf_round(v[8 * i + 2 * j + 1]) + f.y);
      }
    }
  }
}

// One 32-column bf16 chunk (64 bytes) of this thread's output row.  Lane pairs exchange halves so that every store
// instruction of the warp writes whole 32-byte sectors (even lane: first 16 bytes, odd lane: the other 16) instead of
// two half-written sectors per row.
__device__ __forceinline__ void store_chunk(__nv_bfloat16* row_even, __nv_bfloat16* row_odd, bool ok_even, bool ok_odd,
                                            int col0, int n_out, const uint32_t (&o)[16], int lane) {
  const bool odd = lane & 1;
  uint32_t give[8], got[8];  // even lane gives 16-byte pieces 1 and 3 of its row, odd lane pieces 0 and 2
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    give[i] = odd ? o[i] : o[4 + i];
    give[4 + i] = odd ? o[8 + i] : o[12 + i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) got[i] = __shfl_xor_sync(0xffffffffu, give[i], 1);
  // sector t of (row_even: t = 0, 1; row_odd: t = 2, 3): even lane writes bytes [32 (t&1), +16), odd lane the next 16
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const bool to_even_row = t < 2;
    const int sec = t & 1;
    uint4 v;
    if (to_even_row)  // even lane: own piece 2*sec; odd lane: partner's piece 2*sec + 1
      v = odd ? make_uint4(got[4 * sec], got[4 * sec + 1], got[4 * sec + 2], got[4 * sec + 3])
              : make_uint4(o[8 * sec], o[8 * sec + 1], o[8 * sec + 2], o[8 * sec + 3]);
    else              // even lane: partner's piece 2*sec; odd lane: own piece 2*sec + 1
      v = odd ? make_uint4(o[8 * sec + 4], o[8 * sec + 5], o[8 * sec + 6], o[8 * sec + 7])
              : make_uint4(got[4 * sec], got[4 * sec + 1], got[4 * sec + 2], got[4 * sec + 3]);
    const int col = col0 + 16 * sec + (odd ? 8 : 0);
    __nv_bfloat16* row = to_even_row ? row_even : row_odd;
    if ((to_even_row ? ok_even : ok_odd) && col + 8 <= n_out) *reinterpret_cast<uint4*>(row + col) = v;
  }
}

// EW: epilogue warps per CTA.  4 (one per TMEM lane quarter, a 256-column row each) is what the SwiGLU epilogue uses; at
// K = 1024 (the encoder's hidden size) a tile's MMAs take ~8 k cycles and a 256-column GELU or residual row does not fit
// under them (ncu: tensor pipe 38-52 % active), so the other epilogues run 8 warps of 128 columns (RoPE: of half the
// tile's heads) each.
template <int EPI, bool BIAS, int EW = 4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
llm_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);   // (leader's copy is live) the leader's producer expects the bytes of both CTAs
      mbar_init(&empty[s], 1);  // multicast commit of the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);   // multicast commit of the leader
      mbar_init(&acc_empty[a], 2 * EW);  // (leader's copy) one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    tmem_alloc_2cta(&tmem_base_s, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const long long n_units = static_cast<long long>(p.n_mt) * p.n_nt;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    // warp-uniform control flow; one elected lane issues (a divergent `if (lane == 0)` makes every TMA operand a
    // per-lane value and costs a uniform-register broadcast loop per instruction)
    const uint32_t lead_full0 = mapa_shared(smem_u32(&full[0]), 0);
    uint32_t it = 0;
    for (long long u = pair; u < n_units; u += n_pairs) {
      const Unit un = unit_of(u, p);
      const int m0 = un.mt * 2 * TMH + static_cast<int>(rank) * TMH;
      int w0, w1;  // first weight row of this CTA's two 64-row boxes
      if (EPI == LLM_GEMM_SWIGLU) {
        w0 = un.nt * 128 + static_cast<int>(rank) * 64;  // gate rows
        w1 = p.up_offset + w0;                            // the matching up rows
      } else {
        w0 = un.nt * TNP + static_cast<int>(rank) * 128;
        w1 = w0 + 64;
      }
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        const uint32_t lead_full = lead_full0 + s * 8;
        uint8_t* sa = smem + s * STAGE_BYTES;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
          // token rows stream through once per raster group; the group's weight tiles are re-read by every wave
          tma_load_2d_2sm_hint(sa, &tmX, lead_full, kb * KB, m0, p.hint_x);
          tma_load_2d_2sm_hint(sa + A_BYTES, &tmW, lead_full, kb * KB, w0, p.hint_w);
          tma_load_2d_2sm_hint(sa + A_BYTES + BQ_BYTES, &tmW, lead_full, kb * KB, w1, p.hint_w);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader only) =======================
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16_f32(2 * TMH, TNP, 0, 0);
      uint32_t it = 0, ui = 0;
      for (long long u = pair; u < n_units; u += n_pairs, ++ui) {
        const uint32_t a = ui & 1;
        mbar_wait(&acc_empty[a], ((ui >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = umma_smem_desc_sw128(sa, 16, 1024), db = umma_smem_desc_sw128(sa + A_BYTES, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk)
              umma_ss_2cta(tmem + a * TNP, da + ((kk * 32) >> 4), db + ((kk * 32) >> 4), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit_2cta(&empty[s], 3);
            if (kb == p.kblocks - 1) umma_commit_2cta(&acc_full[a], 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================= epilogue (warps 2.., both CTAs): TMEM -> bf16 rows =======================
    static_assert(EW == 4 || (EW == 8 && EPI != LLM_GEMM_SWIGLU) || (EW == 16 && (EPI == LLM_GEMM_PLAIN || EPI == LLM_GEMM_GELU)),
                  "8 epilogue warps: plain / GELU / RoPE rows; 16: plain / GELU rows");
    constexpr int CPW = (TNP / 32) / (EW / 4);  // 32-column chunks per warp
    const int c_lo = ((warp - 2) >> 2) * CPW;   // warps 2-5 take the first column half of their rows, warps 6-9 the second
    const int quarter = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ui = 0;
    for (long long u = pair; u < n_units; u += n_pairs, ++ui) {
      const Unit un = unit_of(u, p);
      const uint32_t a = ui & 1;
      // this lane's row, and the rows of its lane pair (even lane's, odd lane's) for the paired stores
      const long long m_even = static_cast<long long>(un.mt) * 2 * TMH + static_cast<long long>(rank) * TMH + quarter * 32 + (lane & ~1);
      const bool ok_even = m_even < p.M, ok_odd = m_even + 1 < p.M;
      __nv_bfloat16* row_even = p.out + (ok_even ? m_even : 0) * p.ldo;
      __nv_bfloat16* row_odd = p.out + (ok_odd ? m_even + 1 : 0) * p.ldo;
      const __nv_bfloat16* rrow = nullptr;  // this thread's residual row (plain epilogue with a residual only)
      if (EPI == LLM_GEMM_PLAIN && p.resid && m_even + (lane & 1) < p.M) {
        rrow = p.resid + (m_even + (lane & 1)) * p.ldr;
        prefetch_residual(rrow, un.nt * TNP + c_lo * 32, CPW * 32, p.n_out);
      }
      mbar_wait(&acc_full[a], (ui >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem + lane_addr + a * TNP;
      if (EPI == LLM_GEMM_ROPE_EPI) {
        // apply_rotary_pos_emb + cache write: a tile holds whole heads (256 / D of them); the partners d and d + D/2 of a
        // rotation sit D/2 accumulator columns apart in this thread's row.  Roundings as in rope_qkv_kernel (llm_glue.cu).
        const int D = p.D, half = D >> 1, heads_per_tile = TNP / D;
        const long long m_odd = m_even + 1;
        const int b_e = static_cast<int>(m_even / p.L), l_e = static_cast<int>(m_even % p.L);
        const int b_o = static_cast<int>(m_odd / p.L), l_o = static_cast<int>(m_odd % p.L);
        const int my_l = (lane & 1) ? l_o : l_e;
        const __nv_bfloat16* cosr = p.cos + static_cast<long long>(p.pos0 + my_l) * D;
        const __nv_bfloat16* sinr = p.sin + static_cast<long long>(p.pos0 + my_l) * D;
        const bool my_ok = (lane & 1) ? ok_odd : ok_even;
        const int hh_n = heads_per_tile / (EW / 4);  // heads per warp (a tile holds 2 or 4 whole heads)
        const int hh_lo = ((warp - 2) >> 2) * hh_n;
#pragma unroll 1
        for (int hh = hh_lo; hh < hh_lo + hh_n; ++hh) {
          const int head = un.nt * heads_per_tile + hh;
          if (head >= p.H + 2 * p.Hkv) break;
          __nv_bfloat16 *base_e, *base_o;
          if (head < p.H) {
            base_e = p.q_out + ((ok_even ? m_even : 0) * p.H + head) * D;
            base_o = p.q_out + ((ok_odd ? m_odd : 0) * p.H + head) * D;
          } else {
            const bool is_k = head < p.H + p.Hkv;
            __nv_bfloat16* cache = is_k ? p.k_out : p.v_out;
            const int hk = head - p.H - (is_k ? 0 : p.Hkv);
            base_e = cache + (ok_even ? b_e : 0) * p.kv_stride_b + hk * p.kv_stride_h + static_cast<long long>(p.pos0 + (ok_even ? l_e : 0)) * D;
            base_o = cache + (ok_odd ? b_o : 0) * p.kv_stride_b + hk * p.kv_stride_h + static_cast<long long>(p.pos0 + (ok_odd ? l_o : 0)) * D;
          }
          const bool rotate = head < p.H + p.Hkv;
#pragma unroll 1
          for (int c = 0; c < half / 32; ++c) {
            uint32_t xa[32], xb[32];
            const int d0 = c * 32;
            tmem_ld_32x32b_x32(tacc + hh * D + d0, xa);
            tmem_ld_32x32b_x32(tacc + hh * D + half + d0, xb);
            tmem_ld_wait();
            uint32_t o1[16], o2[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) {  // 8 elements at a time
              float x1[8], x2[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                x1[e] = __uint_as_float(xa[8 * v + e]);
                x2[e] = __uint_as_float(xb[8 * v + e]);
              }
              if (BIAS) {
                const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(p.bias + head * D + d0) + v);
                const uint4 b2 = __ldg(reinterpret_cast<const uint4*>(p.bias + head * D + half + d0) + v);
                const uint32_t w1[4] = {b1.x, b1.y, b1.z, b1.w}, w2[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w1[j]));
                  const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w2[j]));
                  x1[2 * j] += f1.x, x1[2 * j + 1] += f1.y, x2[2 * j] += f2.x, x2[2 * j + 1] += f2.y;
                }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) x1[e] = bf_round(x1[e]), x2[e] = bf_round(x2[e]);  // nn.Linear output in bf16
              float r1[8], r2[8];
              if (rotate && my_ok) {
                const uint4 cl = __ldg(reinterpret_cast<const uint4*>(cosr + d0) + v), ch = __ldg(reinterpret_cast<const uint4*>(cosr + half + d0) + v);
                const uint4 sl = __ldg(reinterpret_cast<const uint4*>(sinr + d0) + v), sh = __ldg(reinterpret_cast<const uint4*>(sinr + half + d0) + v);
                const uint32_t cw[4] = {cl.x, cl.y, cl.z, cl.w}, chw[4] = {ch.x, ch.y, ch.z, ch.w};
                const uint32_t sw[4] = {sl.x, sl.y, sl.z, sl.w}, shw[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 c1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cw[j]));
                  const float2 c2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&chw[j]));
                  const float2 s1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sw[j]));
                  const float2 s2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&shw[j]));
                  // rotate_half(x) = cat(-x[D/2:], x[:D/2])
                  r1[2 * j] = bf_round(x1[2 * j] * c1.x) + bf_round(-x2[2 * j] * s1.x);
                  r1[2 * j + 1] = bf_round(x1[2 * j + 1] * c1.y) + bf_round(-x2[2 * j + 1] * s1.y);
                  r2[2 * j] = bf_round(x2[2 * j] * c2.x) + bf_round(x1[2 * j] * s2.x);
                  r2[2 * j + 1] = bf_round(x2[2 * j + 1] * c2.y) + bf_round(x1[2 * j + 1] * s2.y);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) r1[e] = x1[e], r2[e] = x2[e];
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                o1[4 * v + j] = pack_bf16x2(r1[2 * j], r1[2 * j + 1]);
                o2[4 * v + j] = pack_bf16x2(r2[2 * j], r2[2 * j + 1]);
              }
            }
            store_chunk(base_e, base_o, ok_even, ok_odd, d0, D, o1, lane);
            store_chunk(base_e, base_o, ok_even, ok_odd, half + d0, D, o2, lane);
          }
        }
      } else if (EPI == LLM_GEMM_SWIGLU) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {  // (half, 32-column chunk): gate at half*128 + chunk*32, up 64 columns further
          uint32_t g[32], v[32];
          const int hc = (c >> 1) * 128 + (c & 1) * 32;
          tmem_ld_32x32b_x32(tacc + hc, g);
          tmem_ld_32x32b_x32(tacc + hc + 64, v);
          tmem_ld_wait();
          const int col0 = un.nt * 128 + (c >> 1) * 64 + (c & 1) * 32;
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float g0 = bf_round(__uint_as_float(g[2 * i])), g1 = bf_round(__uint_as_float(g[2 * i + 1]));
            const float u0 = bf_round(__uint_as_float(v[2 * i])), u1 = bf_round(__uint_as_float(v[2 * i + 1]));
            o[i] = pack_bf16x2(bf_round(g0 / (1.f + __expf(-g0))) * u0, bf_round(g1 / (1.f + __expf(-g1))) * u1);
          }
          store_chunk(row_even, row_odd, ok_even, ok_odd, col0, p.n_out, o, lane);
        }
      } else {
        uint4 rw[4];
        if (rrow) load_residual(rw, rrow, un.nt * TNP + c_lo * 32, p.n_out);
#pragma unroll 1
        for (int c = c_lo; c < c_lo + CPW; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tacc + c * 32, r);
          tmem_ld_wait();
          const int col0 = un.nt * TNP + c * 32;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (BIAS) {  // 8-column groups (N % 8 == 0); nothing is added past the edge of a ragged last tile
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (col0 + 8 * i + 8 <= p.n_out) {
                const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + 8 * i));
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 bf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[j]));
                  v[8 * i + 2 * j] += bf.x;
                  v[8 * i + 2 * j + 1] += bf.y;
                }
              }
            }
          }
          if (rrow) {  // the chunk's residual values were requested one iteration ago; request the next chunk's now
            apply_residual(v, rw, col0, p.n_out);
            if (c + 1 < c_lo + CPW) load_residual(rw, rrow, col0 + 32, p.n_out);
          }
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = pack_out<EPI>(v[2 * i], v[2 * i + 1]);
          store_chunk(row_even, row_odd, ok_even, ok_odd, col0, p.n_out, o, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[a]), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody may still read the peer's shared memory / TMEM when a CTA tears down
  if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// Wide-tile variant (experiment, SDAG_GEMM_WIDE=1; PLAIN without bias and SWIGLU only): a pair tile of 256 token rows x
// 512 weight rows.  The 128 token rows a CTA loads per k-block feed two MMAs (N = 256 each), so the L2 -> SM bytes per FLOP
// drop to 0.75x of the 256 x 256 tile.  All 512 TMEM columns hold one accumulator set; the two 256-column halves are handed
// to the epilogue and back separately, so the next tile's first MMAs (half 0) overlap the draining of half 1.
// ------------------------------------------------------------------------------------------------
constexpr int W_STAGES = 4;
constexpr int W_STAGE_BYTES = A_BYTES + 4 * BQ_BYTES;  // token rows + two MMAs' worth of weight rows (2 x 2 boxes of 64)
constexpr int W_SMEM_BYTES = W_STAGES * W_STAGE_BYTES + 1024;
constexpr int W_THREADS = 320;  // producer warp, MMA warp, eight epilogue warps

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W_THREADS, 1)
llm_gemm_wide_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[W_STAGES], empty[W_STAGES], acc_full, acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (tid == 0) {
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty[0], 8);
    mbar_init(&acc_empty[1], 8);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    tmem_alloc_2cta(&tmem_base_s, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const long long n_units = static_cast<long long>(p.n_mt) * p.n_nt;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    const uint32_t lead_full0 = mapa_shared(smem_u32(&full[0]), 0);
    uint32_t it = 0;
    for (long long u = pair; u < n_units; u += n_pairs) {
      const Unit un = unit_of(u, p);
      const int m0 = un.mt * 2 * TMH + static_cast<int>(rank) * TMH;
      int w[4];  // first weight rows of the four 64-row boxes: (MMA 0: box 0, 1), (MMA 1: box 0, 1)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (EPI == LLM_GEMM_SWIGLU) {
          w[2 * j] = un.nt * 256 + j * 128 + static_cast<int>(rank) * 64;  // gate rows
          w[2 * j + 1] = p.up_offset + w[2 * j];                           // the matching up rows
        } else {
          w[2 * j] = un.nt * 512 + j * 256 + static_cast<int>(rank) * 128;
          w[2 * j + 1] = w[2 * j] + 64;
        }
      }
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const int s = it % W_STAGES;
        mbar_wait(&empty[s], ((it / W_STAGES) & 1) ^ 1);
        const uint32_t lead_full = lead_full0 + s * 8;
        uint8_t* sa = smem + s * W_STAGE_BYTES;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[s], 2 * W_STAGE_BYTES);
          tma_load_2d_2sm_hint(sa, &tmX, lead_full, kb * KB, m0, p.hint_x);
#pragma unroll
          for (int b = 0; b < 4; ++b)
            tma_load_2d_2sm_hint(sa + A_BYTES + b * BQ_BYTES, &tmW, lead_full, kb * KB, w[b], p.hint_w);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16_f32(2 * TMH, TNP, 0, 0);
      uint32_t it = 0, ui = 0;
      for (long long u = pair; u < n_units; u += n_pairs, ++ui) {
        for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
          const int s = it % W_STAGES;
          mbar_wait(&full[s], (it / W_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * W_STAGE_BYTES);
          const uint64_t da = umma_smem_desc_sw128(sa, 16, 1024);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (kb == 0) {  // half j of the accumulator must have been drained by the previous tile's epilogue
              mbar_wait(&acc_empty[j], (ui & 1) ^ 1);
              tc_fence_after();
            }
            const uint64_t db = umma_smem_desc_sw128(sa + A_BYTES + j * 2 * BQ_BYTES, 16, 1024);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < KB / 16; ++kk)
                umma_ss_2cta(tmem + j * TNP, da + ((kk * 32) >> 4), db + ((kk * 32) >> 4), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            __syncwarp();
          }
          if (elect_one()) {
            umma_commit_2cta(&empty[s], 3);
            if (kb == p.kblocks - 1) umma_commit_2cta(&acc_full, 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // eight epilogue warps: warps 2-5 drain accumulator half 0, warps 6-9 half 1 (TMEM lane quarter = warp % 4), so both
    // halves return to the MMA issuer after half the drain time
    const int quarter = warp & 3;
    const int j = (warp - 2) >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ui = 0;
    for (long long u = pair; u < n_units; u += n_pairs, ++ui) {
      const Unit un = unit_of(u, p);
      const long long m_even = static_cast<long long>(un.mt) * 2 * TMH + static_cast<long long>(rank) * TMH + quarter * 32 + (lane & ~1);
      const bool ok_even = m_even < p.M, ok_odd = m_even + 1 < p.M;
      __nv_bfloat16* row_even = p.out + (ok_even ? m_even : 0) * p.ldo;
      __nv_bfloat16* row_odd = p.out + (ok_odd ? m_even + 1 : 0) * p.ldo;
      const __nv_bfloat16* rrow = nullptr;
      if (EPI == LLM_GEMM_PLAIN && p.resid && m_even + (lane & 1) < p.M) {
        rrow = p.resid + (m_even + (lane & 1)) * p.ldr;
        prefetch_residual(rrow, un.nt * 512 + j * TNP, TNP, p.n_out);
      }
      mbar_wait(&acc_full, ui & 1);
      tc_fence_after();
      {
        const uint32_t tacc = tmem + lane_addr + j * TNP;
        if (EPI == LLM_GEMM_SWIGLU) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t g[32], v[32];
            const int hc = (c >> 1) * 128 + (c & 1) * 32;
            tmem_ld_32x32b_x32(tacc + hc, g);
            tmem_ld_32x32b_x32(tacc + hc + 64, v);
            tmem_ld_wait();
            const int col0 = un.nt * 256 + j * 128 + (c >> 1) * 64 + (c & 1) * 32;
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float g0 = bf_round(__uint_as_float(g[2 * i])), g1 = bf_round(__uint_as_float(g[2 * i + 1]));
              const float u0 = bf_round(__uint_as_float(v[2 * i])), u1 = bf_round(__uint_as_float(v[2 * i + 1]));
              o[i] = pack_bf16x2(bf_round(g0 / (1.f + __expf(-g0))) * u0, bf_round(g1 / (1.f + __expf(-g1))) * u1);
            }
            store_chunk(row_even, row_odd, ok_even, ok_odd, col0, p.n_out, o, lane);
          }
        } else {
          uint4 rw[4];
          if (rrow) load_residual(rw, rrow, un.nt * 512 + j * TNP, p.n_out);
#pragma unroll 1
          for (int c = 0; c < TNP / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tacc + c * 32, r);
            tmem_ld_wait();
            const int col0 = un.nt * 512 + j * TNP + c * 32;
            uint32_t o[16];
            if (rrow) {
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
              apply_residual(v, rw, col0, p.n_out);
              if (c + 1 < TNP / 32) load_residual(rw, rrow, col0 + 32, p.n_out);
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            }
            store_chunk(row_even, row_odd, ok_even, ok_odd, col0, p.n_out, o, lane);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[j]), 0));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

// Shared host part: tensor maps, raster parameters, launch.  `p` arrives with the epilogue-specific fields filled in.
int run_gemm(const void* x, int64_t ldx, const void* weight, int64_t M, int N, int K, int epi, Params& p, cudaStream_t st) {
  CUtensorMap tmX, tmW;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t str[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t box[2] = {KB, TMH};
    if (make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box)) return SDAG_ERR_DRIVER;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {KB, 64};
    if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, weight, dims, str, box)) return SDAG_ERR_DRIVER;
  }
  static bool configured[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (!configured[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_PLAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_PLAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_SWIGLU, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_GELU, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_GELU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_PLAIN, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_GELU, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_PLAIN, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_PLAIN, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_GELU, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_GELU, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_ROPE_EPI, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY((cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_ROPE_EPI, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_ROPE_EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_kernel<LLM_GEMM_ROPE_EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_wide_kernel<LLM_GEMM_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES));
    SDAG_CUDA_TRY(cudaFuncSetAttribute(llm_gemm_wide_kernel<LLM_GEMM_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES));
    configured[dev] = true;
  }
  const int n_sm = device_sm_count();
  const bool swiglu = epi == LLM_GEMM_SWIGLU;
  // 256 x 512 pair tiles move 0.75x the L2 -> SM bytes per FLOP but cannot double-buffer the accumulator: measured faster
  // where the main loop is long enough to hide the exposed epilogue (K = 14336 down projection, plain epilogue: 1322 ->
  // 1395 TFLOP/s sustained, cuBLAS 1407) and slower at K = 4096 (o, gate/up).  With the residual epilogue the exposed part
  // grows and the order flips (same-box A/B inside the prefill step: down + residual 1.84 ms on 256 x 256 tiles, whose
  // double-buffered accumulators hide the epilogue, vs 1.91 ms on wide tiles), so a residual keeps the 256 x 256 tiles.
  // SDAG_GEMM_WIDE=0 / 1 forces one form (experiments).
  static int wide_env = -2;
  if (wide_env == -2) {
    const char* e = getenv("SDAG_GEMM_WIDE");
    wide_env = e ? (atoi(e) ? 1 : 0) : -1;
  }
  const bool wide_ok = swiglu || (epi == LLM_GEMM_PLAIN && !p.bias);
  const bool wide = wide_ok && (wide_env >= 0 ? wide_env == 1 : (!swiglu && !p.resid && K >= 8192 && M >= 2048));
  p.M = M;
  p.n_mt = static_cast<int>((M + 2 * TMH - 1) / (2 * TMH));
  p.n_nt = swiglu ? (p.n_out + 127) / 128 : (N + TNP - 1) / TNP;
  if (wide) p.n_nt = (p.n_nt + 1) / 2;
  p.kblocks = (K + KB - 1) / KB;
  p.up_offset = p.n_out;
  // raster group: as many weight tiles as stay L2-resident (~60 MB, loaded EVICT_LAST) while the token tiles stream past
  // them; EVICT_FIRST on the token rows was measured to hurt (bit 0 of the knob) and stays off
  static int group_env = -1, hint_env = -1;
  if (group_env < 0) {
    const char* e = getenv("SDAG_GEMM_GROUP_N");  // tuning knobs
    group_env = e ? atoi(e) : 0;
    const char* h = getenv("SDAG_GEMM_HINTS");
    hint_env = h ? atoi(h) : 2;
  }
  int gn = static_cast<int>((60ll << 20) / (static_cast<long long>(swiglu ? 256 : TNP) * (wide ? 2 : 1) * K * 2));
  gn = gn < 2 ? 2 : (gn > GROUP_N ? GROUP_N : gn);
  p.group_n = group_env > 0 ? group_env : gn;
  p.hint_x = (hint_env & 1) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  p.hint_w = (hint_env & 2) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
  const long long units = static_cast<long long>(p.n_mt) * p.n_nt;
  const long long pairs = units < n_sm / 2 ? units : n_sm / 2;
  const unsigned grid = static_cast<unsigned>(2 * pairs);
  // Eight epilogue warps for the plain / residual / GELU / RoPE rows: decisive for short main loops (K = 1024, the encoder: a
  // 256-column row does not fit under ~8 k cycles of MMAs), a consistent +0.3 % on the Llama prefill step (same-box ABAB:
  // o + residual 1254 -> 1263, q/k/v + RoPE 1290 -> 1301 TFLOP/s).  SDAG_GEMM_EW=4 / 8 forces one form.
  static int ew_env = -1;
  if (ew_env < 0) {
    const char* e = getenv("SDAG_GEMM_EW");
    ew_env = e ? atoi(e) : 0;
  }
  const bool ew8 = !wide && !swiglu && (ew_env ? ew_env >= 8 : true);
  // sixteen (two 32-column chunks per warp) for the biased plain / GELU rows of short main loops: experiment, SDAG_GEMM_EW=16
  const bool ew16 = !wide && ew_env == 16 && p.bias && (epi == LLM_GEMM_PLAIN || epi == LLM_GEMM_GELU);
  if (ew16 && epi == LLM_GEMM_GELU)
    llm_gemm_kernel<LLM_GEMM_GELU, true, 16><<<grid, 576, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew16)
    llm_gemm_kernel<LLM_GEMM_PLAIN, true, 16><<<grid, 576, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew8 && epi == LLM_GEMM_ROPE_EPI && p.bias)
    llm_gemm_kernel<LLM_GEMM_ROPE_EPI, true, 8><<<grid, 320, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew8 && epi == LLM_GEMM_ROPE_EPI)
    llm_gemm_kernel<LLM_GEMM_ROPE_EPI, false, 8><<<grid, 320, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew8 && epi == LLM_GEMM_GELU && p.bias)
    llm_gemm_kernel<LLM_GEMM_GELU, true, 8><<<grid, 320, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew8 && epi == LLM_GEMM_GELU)
    llm_gemm_kernel<LLM_GEMM_GELU, false, 8><<<grid, 320, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew8 && p.bias)
    llm_gemm_kernel<LLM_GEMM_PLAIN, true, 8><<<grid, 320, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (ew8)
    llm_gemm_kernel<LLM_GEMM_PLAIN, false, 8><<<grid, 320, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (wide && swiglu)
    llm_gemm_wide_kernel<LLM_GEMM_SWIGLU><<<grid, W_THREADS, W_SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (wide)
    llm_gemm_wide_kernel<LLM_GEMM_PLAIN><<<grid, W_THREADS, W_SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (swiglu)
    llm_gemm_kernel<LLM_GEMM_SWIGLU, false><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (epi == LLM_GEMM_ROPE_EPI && p.bias)
    llm_gemm_kernel<LLM_GEMM_ROPE_EPI, true><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (epi == LLM_GEMM_ROPE_EPI)
    llm_gemm_kernel<LLM_GEMM_ROPE_EPI, false><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (epi == LLM_GEMM_GELU && p.bias)
    llm_gemm_kernel<LLM_GEMM_GELU, true><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (epi == LLM_GEMM_GELU)
    llm_gemm_kernel<LLM_GEMM_GELU, false><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (p.bias)
    llm_gemm_kernel<LLM_GEMM_PLAIN, true><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  else
    llm_gemm_kernel<LLM_GEMM_PLAIN, false><<<grid, 192, SMEM_BYTES, st>>>(tmX, tmW, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

static int llm_gemm_impl(const void* x, int64_t ldx, const void* weight, const void* bias, const void* residual, int64_t ldr,
                         void* out, int64_t ldo, int64_t M, int N, int K, int epilogue, void* stream) {
  if (!x || !weight || !out || M < 0 || N <= 0 || K <= 0) return SDAG_ERR_INVALID;
  if (epilogue != LLM_GEMM_PLAIN && epilogue != LLM_GEMM_SWIGLU && epilogue != LLM_GEMM_GELU) return SDAG_ERR_INVALID;
  if (residual && (epilogue != LLM_GEMM_PLAIN || (ldr % 8) || ldr < N || (reinterpret_cast<uintptr_t>(residual) & 15)))
    return SDAG_ERR_INVALID;
  if ((K % 8) || (ldx % 8) || (ldo % 8) || ldx < K) return SDAG_ERR_INVALID;  // 16-byte rows for TMA and the stores
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out)) & 15)
    return SDAG_ERR_INVALID;
  const bool swiglu = epilogue == LLM_GEMM_SWIGLU;
  if (swiglu && ((N % 16) || bias)) return SDAG_ERR_INVALID;
  if (!swiglu && (N % 8)) return SDAG_ERR_INVALID;
  const int n_out = swiglu ? N / 2 : N;
  if (ldo < n_out) return SDAG_ERR_INVALID;
  if (M == 0) return SDAG_OK;
  if (M > (1ll << 31) - 512) return SDAG_ERR_UNSUPPORTED;
  Params p = {};
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.resid = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.n_out = n_out;
  return run_gemm(x, ldx, weight, M, N, K, epilogue, p, static_cast<cudaStream_t>(stream));
}

extern "C" int llm_gemm(const void* x, int64_t ldx, const void* weight, const void* bias, void* out, int64_t ldo, int64_t M,
                        int N, int K, int epilogue, void* stream) {
  return llm_gemm_impl(x, ldx, weight, bias, nullptr, 0, out, ldo, M, N, K, epilogue, stream);
}

extern "C" int llm_gemm_residual(const void* x, int64_t ldx, const void* weight, const void* bias, const void* residual,
                                 int64_t ldr, void* out, int64_t ldo, int64_t M, int N, int K, void* stream) {
  if (!residual) return SDAG_ERR_INVALID;
  return llm_gemm_impl(x, ldx, weight, bias, residual, ldr, out, ldo, M, N, K, LLM_GEMM_PLAIN, stream);
}

extern "C" int llm_gemm_rope(const void* x, int64_t ldx, const void* weight, const void* bias, const void* cos, const void* sin,
                             void* q_out, void* k_out, void* v_out, int B, int L, int H, int Hkv, int D, int K,
                             int64_t kv_stride_b, int64_t kv_stride_h, int pos0, void* stream) {
  if (!x || !weight || !cos || !sin || !q_out || !k_out || !v_out) return SDAG_ERR_INVALID;
  if (B < 0 || L <= 0 || H <= 0 || Hkv <= 0 || K <= 0 || pos0 < 0) return SDAG_ERR_INVALID;
  if ((K % 8) || (ldx % 8) || ldx < K || (kv_stride_b % 8) || (kv_stride_h % 8)) return SDAG_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(q_out) |
       reinterpret_cast<uintptr_t>(k_out) | reinterpret_cast<uintptr_t>(v_out) | reinterpret_cast<uintptr_t>(cos) |
       reinterpret_cast<uintptr_t>(sin) | reinterpret_cast<uintptr_t>(bias)) & 15)
    return SDAG_ERR_INVALID;
  // a 256-column tile must hold whole heads of one kind
  if (D != 64 && D != 128) return SDAG_ERR_UNSUPPORTED;
  const int heads_per_tile = TNP / D;
  if ((H % heads_per_tile) || (Hkv % heads_per_tile)) return SDAG_ERR_UNSUPPORTED;
  const int64_t M = static_cast<int64_t>(B) * L;
  if (M == 0) return SDAG_OK;
  if (M > (1ll << 31) - 512) return SDAG_ERR_UNSUPPORTED;
  Params p = {};
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.cos = static_cast<const __nv_bfloat16*>(cos);
  p.sin = static_cast<const __nv_bfloat16*>(sin);
  p.q_out = static_cast<__nv_bfloat16*>(q_out);
  p.k_out = static_cast<__nv_bfloat16*>(k_out);
  p.v_out = static_cast<__nv_bfloat16*>(v_out);
  p.kv_stride_b = kv_stride_b;
  p.kv_stride_h = kv_stride_h;
  p.L = L;
  p.H = H;
  p.Hkv = Hkv;
  p.D = D;
  p.pos0 = pos0;
  const int N = (H + 2 * Hkv) * D;
  p.n_out = N;
  return run_gemm(x, ldx, weight, M, N, K, LLM_GEMM_ROPE_EPI, p, static_cast<cudaStream_t>(stream));
}
