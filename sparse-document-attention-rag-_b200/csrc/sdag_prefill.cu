// SDAG prefill attention for sm_100a: block-sparse FlashAttention-style forward over the tile lists of
// sdag_blockmap.cu.  Replaces the masked SDPA call the reference reaches from SDAG.py:176
// (HF modeling_llama.py:199-221 / integrations/sdpa_attention.py:40-104).
//
// One CTA per (query tile, head, prompt); 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs.
//   warp 0      TMA producer: Q once, then K_j / V_j tiles (single-buffered each, separate barriers)
//   warp 1      tcgen05 issuer: S = Q K_j^T (SS, both K-major) -> TMEM; O += P_j V_j (TS: P in TMEM,
//               V MN-major straight from its natural [kv][d] layout)
//   warps 2-5   softmax: one query row per thread (TMEM lane), fp32 online softmax with lazy rescaling,
//               P written back over S as packed bf16; then the epilogue (O / l -> bf16 -> global)
// TMEM columns: [0,128) S / P, [128,128+D) O.
#include <math.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int BM = SDAG_TILE;
constexpr int BN = SDAG_TILE;
constexpr int KV_FLAG_BITMASK = 1;
constexpr int VSUM_CHUNKS = 32;
constexpr float RESCALE_TAU = 8.0f;  // log2 units: P stays <= 2^8 with a stale max

struct PrefillParams {
  __nv_bfloat16* out;
  long long so_b, so_l, so_h;
  const int32_t* blockmaps;
  long long bm_stride_ints;
  const float* vsum;  // [B][Hkv][VSUM_CHUNKS][D] partial column sums of V, then [B][Hkv][D] mean of V
  int B, H, Hkv, Lq, Skv;
  float scale_log2;
};

template <int D>
struct Smem {
  static constexpr int kQ = BM * D * 2;
  static constexpr int kK = BN * D * 2;
  static constexpr int kV = BN * D * 2;
  static constexpr int kTotal = kQ + kK + kV + 1024;  // + alignment slack
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

#ifdef SDAG_DEBUG_TIMING
__device__ unsigned long long* g_timing = nullptr;  // [grid][8] globaltimer stamps (bring-up only)
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define STAMP(slot)                                                                       \
  do {                                                                                    \
    if (g_timing && threadIdx.x == 64) g_timing[blockIdx.x * 8ull + (slot)] = gtime();    \
  } while (0)
#else
#define STAMP(slot) \
  do {              \
  } while (0)
#endif

template <int D>
__global__ void __launch_bounds__(192, 2)
sdag_prefill_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const PrefillParams p) {
  // ---- which (query tile, head, prompt) ----
  const int per_tile = p.H * p.B;
  const int t = blockIdx.x / per_tile;
  const int rem = blockIdx.x - t * per_tile;
  const int b = rem / p.H;
  const int h = rem - b * p.H;
  const int32_t* map = p.blockmaps + static_cast<long long>(b) * p.bm_stride_ints;
  STAMP(0);
  if (t >= map[0] || map[2] != 0) return;  // uniform: before any barrier / allocation
  const int4 qt = reinterpret_cast<const int4*>(map + SDAG_BLOCKMAP_HEADER)[t];
  const int q0 = qt.x, nrows = qt.y, n_kv = qt.w;
  const int4* kvt = reinterpret_cast<const int4*>(map + SDAG_BLOCKMAP_HEADER) + map[5] + qt.z;
  const int hkv = h / (p.H / p.Hkv);

  if (n_kv == 0) {
    // Gap rows only (no allowed key): every row is the mean of V over all keys.  No tensor work, no TMEM.
    const float* vm = p.vsum + (static_cast<long long>(p.B) * p.Hkv * VSUM_CHUNKS + static_cast<long long>(b) * p.Hkv + hkv) * D;
    for (int i = threadIdx.x; i < nrows * (D / 8); i += blockDim.x) {
      const int rr = i / (D / 8), ch = i % (D / 8);
      const float4 a = *reinterpret_cast<const float4*>(vm + ch * 8);
      const float4 c4 = *reinterpret_cast<const float4*>(vm + ch * 8 + 4);
      __nv_bfloat16* orow = p.out + b * p.so_b + static_cast<long long>(q0 + rr) * p.so_l + h * p.so_h;
      *reinterpret_cast<uint4*>(orow + ch * 8) =
          make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(c4.x, c4.y), pack_bf16x2(c4.z, c4.w));
    }
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Smem<D>::kQ;
  uint8_t* sV = sK + Smem<D>::kK;
  __shared__ uint64_t q_full, k_full, k_empty, v_full, v_empty, s_full, p_full, o_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&k_full, 1);
    mbar_init(&k_empty, 1);
    mbar_init(&v_full, 1);
    mbar_init(&v_empty, 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_full, 128);
    mbar_init(&o_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  STAMP(1);
  const uint32_t tmem_S = tmem;        // fp32 scores; P (bf16x2) is written over its first 64 columns
  const uint32_t tmem_O = tmem + 128;  // fp32 output accumulator, D columns

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0 && n_kv > 0) {
      mbar_expect_tx(&q_full, Smem<D>::kQ);
      for (int dh = 0; dh < D / 64; ++dh) tma_load_4d(sQ + dh * BM * 128, &tmQ, &q_full, dh * 64, q0, h, b);
      for (int j = 0; j < n_kv; ++j) {
        const int k0 = kvt[j].x;
        mbar_wait(&k_empty, (j & 1) ^ 1);
        mbar_expect_tx(&k_full, Smem<D>::kK);
        for (int dh = 0; dh < D / 64; ++dh) tma_load_4d(sK + dh * BN * 128, &tmK, &k_full, dh * 64, k0, hkv, b);
        mbar_wait(&v_empty, (j & 1) ^ 1);
        mbar_expect_tx(&v_full, Smem<D>::kV);
        for (int dh = 0; dh < D / 64; ++dh) tma_load_4d(sV + dh * BN * 128, &tmV, &v_full, dh * 64, k0, hkv, b);
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (n_kv > 0) {
      mbar_wait(&q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int ncols = kvt[j].y;
        const uint32_t n_eff = static_cast<uint32_t>((ncols + 15) & ~15);
        const int pv_steps = (ncols + 15) >> 4;
        mbar_wait(&k_full, j & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t idesc = umma_idesc_bf16_f32(BM, n_eff, 0, 0);
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = (ks >> 2) * (BM * 128) + (ks & 3) * 32;
            const uint32_t offk = (ks >> 2) * (BN * 128) + (ks & 3) * 32;
            umma_ss(tmem_S, umma_smem_desc_sw128(smem_u32(sQ) + off, 16, 1024),
                    umma_smem_desc_sw128(smem_u32(sK) + offk, 16, 1024), idesc, ks > 0);
          }
          umma_commit(&k_empty);  // K tile consumed -> producer may load K_{j+1}
          umma_commit(&s_full);   // scores ready (and every earlier MMA, incl. P_{j-1} V_{j-1}, retired)
        }
        __syncwarp();
        mbar_wait(&p_full, j & 1);
        mbar_wait(&v_full, j & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t idesc = umma_idesc_bf16_f32(BM, D, 0, 1);  // B = V, MN-major
          for (int ks = 0; ks < pv_steps; ++ks) {
            umma_ts(tmem_O, tmem_S + ks * 8, umma_smem_desc_sw128(smem_u32(sV) + ks * 16 * 128, BN * 128, 1024), idesc,
                    (j > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&v_empty);
          if (j == n_kv - 1) umma_commit(&o_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= softmax + epilogue (warps 2..5) =======================
    const int quarter = warp & 3;  // TMEM lanes this warp may touch: [32*quarter, 32*quarter+32)
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint32_t r[32], r2[32];

    for (int j = 0; j < n_kv; ++j) {
      const int4 kt = kvt[j];
      const int ncols = kt.y;
      const bool bitmask = (kt.w & KV_FLAG_BITMASK) != 0;
      int lim = min(ncols - 1, row + kt.z);  // allowed columns of this row: c <= lim
      // warp-uniform bounds of lim over the 32 rows of this warp: chunks at or below wlim_lo need no masking at
      // all, chunks above wlim_hi hold no allowed column for any lane
      int wlim_lo = min(ncols - 1, quarter * 32 + kt.z);
      int wlim_hi = min(ncols - 1, quarter * 32 + 31 + kt.z);
      uint32_t bits[4] = {0, 0, 0, 0};
      if (bitmask) {
        lim = ncols - 1;
        wlim_lo = -1;  // every chunk takes the masked path
        wlim_hi = ncols - 1;
        if (q0 + row < p.Lq) {
          const uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(map + map[10]) +
                                                          static_cast<long long>(q0 + row) * map[7] + (kt.x >> 5));
          bits[0] = w.x, bits[1] = w.y, bits[2] = w.z, bits[3] = w.w;
        }
      }
      const int nchunks = (ncols + 31) >> 5;
      mbar_wait(&s_full, j & 1);
      tc_fence_after();
      if (j == 0) STAMP(2);

      // Single pass over the scores (TMEM reads are the scarce resource): per 32-column chunk take the chunk
      // maximum, keep the running reference maximum m unless it is exceeded by more than RESCALE_TAU, and write
      // P = exp2(S*scale - m) straight back over the S columns already consumed.  When m does move, everything
      // accumulated against the old reference — l, O, and the P chunks of this tile already written — is
      // rescaled (rare after the first chunk of the first tile).
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      tmem_ld_32x32b_x32(tmem_S + lane_addr, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t* cur = (c & 1) ? r2 : r;
        uint32_t* nxt = (c & 1) ? r : r2;
        if (c < nchunks) {
          const bool full_chunk = wlim_lo >= c * 32 + 31;
          const bool empty_chunk = wlim_hi < c * 32;
          const uint32_t bw = bitmask ? bits[c] : 0xffffffffu;
          const int rel = lim - c * 32;  // columns i <= rel are allowed (masked chunks)
          uint32_t pk[16];
          if (!empty_chunk) {
            float mx = -INFINITY;
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const bool ok = (i <= rel) && ((bw >> i) & 1u);
                mx = fmaxf(mx, ok ? __uint_as_float(cur[i]) : -INFINITY);
              }
            }
            mx *= p.scale_log2;
            const bool need = mx > m + RESCALE_TAU;  // also the first finite maximum (m == -inf)
            if (__any_sync(0xffffffffu, need)) {
              const float alpha = need ? exp2f(m - mx) : 1.f;
              if (need) m = mx;
              l *= alpha;
              l0 *= alpha, l1 *= alpha, l2 *= alpha, l3 *= alpha;
              if (j > 0) {  // O holds P_0 V_0 .. P_{j-1} V_{j-1}, all retired (s_full commit)
                for (int oc = 0; oc < D / 32; ++oc) {
                  tmem_ld_32x32b_x32(tmem_O + lane_addr + oc * 32, nxt);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) nxt[i] = __float_as_uint(__uint_as_float(nxt[i]) * alpha);
                  tmem_st_32x32b_x32(tmem_O + lane_addr + oc * 32, nxt);
                }
              }
              for (int pc = 0; pc < c; ++pc) {  // P chunks of this tile written against the old reference
                tmem_ld_32x32b_x16(tmem_S + lane_addr + pc * 16, nxt);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&nxt[i]));
                  nxt[i] = pack_bf16x2(f.x * alpha, f.y * alpha);
                }
                tmem_st_32x32b_x16(tmem_S + lane_addr + pc * 16, nxt);
              }
            }
          }
          if (c + 1 < nchunks) tmem_ld_32x32b_x32(tmem_S + lane_addr + (c + 1) * 32, nxt);
          const float neg_m = (m == -INFINITY) ? 0.f : -m;
          if (empty_chunk) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          } else if (full_chunk) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float p0 = exp2f(fmaf(__uint_as_float(cur[i]), p.scale_log2, neg_m));
              const float p1 = exp2f(fmaf(__uint_as_float(cur[i + 1]), p.scale_log2, neg_m));
              const float p2 = exp2f(fmaf(__uint_as_float(cur[i + 2]), p.scale_log2, neg_m));
              const float p3 = exp2f(fmaf(__uint_as_float(cur[i + 3]), p.scale_log2, neg_m));
              l0 += p0, l1 += p1, l2 += p2, l3 += p3;
              pk[i >> 1] = pack_bf16x2(p0, p1);
              pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const bool ok0 = (i <= rel) && ((bw >> i) & 1u);
              const bool ok1 = (i + 1 <= rel) && ((bw >> (i + 1)) & 1u);
              const float p0 = ok0 ? exp2f(fmaf(__uint_as_float(cur[i]), p.scale_log2, neg_m)) : 0.f;
              const float p1 = ok1 ? exp2f(fmaf(__uint_as_float(cur[i + 1]), p.scale_log2, neg_m)) : 0.f;
              l0 += p0, l1 += p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
          }
          if (c + 1 < nchunks) tmem_ld_wait();  // S chunk c+1 is in registers before P chunk c lands on S columns
          tmem_st_32x32b_x16(tmem_S + lane_addr + c * 16, pk);
        }
      }
      l += (l0 + l1) + (l2 + l3);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full);
      if (j == 0) STAMP(3);
    }

    // ---- epilogue ----
    if (n_kv > 0) {
      mbar_wait(&o_full, 0);
      tc_fence_after();
    }
    STAMP(4);
    const bool valid = row < nrows;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    // Stage the bf16 row in shared memory (Q/K/V tiles are dead once o_full fired), then let the bulk-copy engine
    // write each row as one contiguous D*2-byte burst: 32 rows per warp would otherwise mean 32 partial lines per
    // store instruction.
    constexpr int ROW_BYTES = D * 2 + 16;  // +16 B: conflict-free 16-byte stores at a row stride
    uint8_t* srow = smem + row * ROW_BYTES;
    const float* vm = p.vsum + (static_cast<long long>(p.B) * p.Hkv * VSUM_CHUNKS + static_cast<long long>(b) * p.Hkv + hkv) * D;
    if (n_kv > 0) tmem_ld_32x32b_x32(tmem_O + lane_addr, r);
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t* cur = (c & 1) ? r2 : r;
      uint32_t* nxt = (c & 1) ? r : r2;
      if (n_kv > 0) {
        tmem_ld_wait();
        if (c + 1 < D / 32) tmem_ld_32x32b_x32(tmem_O + lane_addr + (c + 1) * 32, nxt);
      }
      uint32_t o[16];
      if (l > 0.f) {
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          o[i >> 1] = pack_bf16x2(__uint_as_float(cur[i]) * inv, __uint_as_float(cur[i + 1]) * inv);
      } else {
        // no allowed key in the whole row: uniform softmax over all Skv keys = mean of V
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 f = valid ? *reinterpret_cast<const float4*>(vm + c * 32 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          o[i >> 1] = pack_bf16x2(f.x, f.y);
          o[(i >> 1) + 1] = pack_bf16x2(f.z, f.w);
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(srow + c * 64);
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
      dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
    }
    fence_proxy_async_smem();
    if (valid) {
      __nv_bfloat16* orow = p.out + b * p.so_b + static_cast<long long>(q0 + row) * p.so_l + h * p.so_h;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(orow), "r"(smem_u32(srow)), "n"(D * 2)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem may be released once it has been read
  }

  STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
  STAMP(6);
}

// Partial column sums of V over VSUM_CHUNKS row ranges: vsum[b][hkv][chunk][d] (fp32, deterministic).
template <int D>
__global__ void __launch_bounds__(256) v_colsum_kernel(const __nv_bfloat16* __restrict__ v, long long sv_b, long long sv_h,
                                                       long long sv_l, int Skv, float* __restrict__ vsum,
                                                       const int32_t* __restrict__ blockmaps, long long bm_stride_ints) {
  const int chunk = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  if (blockmaps && blockmaps[static_cast<long long>(b) * bm_stride_ints + 3] == 0) return;  // no gap rows
  constexpr int TPR = D / 8;        // threads per row (16 B each)
  constexpr int RPI = 256 / TPR;    // rows per iteration
  const int tr = threadIdx.x / TPR, tc = threadIdx.x % TPR;
  const int rows_per_chunk = (Skv + VSUM_CHUNKS - 1) / VSUM_CHUNKS;
  const int r0 = chunk * rows_per_chunk, r1 = min(Skv, r0 + rows_per_chunk);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __nv_bfloat16* base = v + b * sv_b + hkv * sv_h;
  for (int rr = r0 + tr; rr < r1; rr += RPI) {
    const uint4 w = *reinterpret_cast<const uint4*>(base + rr * sv_l + tc * 8);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h2[i]);
      acc[2 * i] += f.x;
      acc[2 * i + 1] += f.y;
    }
  }
  __shared__ float red[RPI][D];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[tr][tc * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < D) {
    float s = 0.f;
    for (int q = 0; q < RPI; ++q) s += red[q][threadIdx.x];
    vsum[((static_cast<long long>(b) * gridDim.y + hkv) * VSUM_CHUNKS + chunk) * D + threadIdx.x] = s;
  }
}

// vmean[b][hkv][d] = (sum of the chunk partials, fixed order) / Skv, stored after the partials.
template <int D>
__global__ void __launch_bounds__(D) v_mean_kernel(float* __restrict__ vsum, int B, int Hkv, int Skv,
                                                   const int32_t* __restrict__ blockmaps, long long bm_stride_ints) {
  const int hkv = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  if (blockmaps && blockmaps[static_cast<long long>(b) * bm_stride_ints + 3] == 0) return;
  const float* part = vsum + (static_cast<long long>(b) * Hkv + hkv) * VSUM_CHUNKS * D;
  float s = 0.f;
  for (int ch = 0; ch < VSUM_CHUNKS; ++ch) s += part[ch * D + d];
  vsum[(static_cast<long long>(B) * Hkv * VSUM_CHUNKS + static_cast<long long>(b) * Hkv + hkv) * D + d] = s / static_cast<float>(Skv);
}

int make_map_4d(CUtensorMap* m, const void* base, int D, int rows, int heads, int B, int64_t s_l, int64_t s_h,
                int64_t s_b, int box_rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(rows), static_cast<uint64_t>(heads),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(s_l) * 2, static_cast<uint64_t>(s_h) * 2, static_cast<uint64_t>(s_b) * 2};
  uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), 1, 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box);
}

template <int D>
int launch_prefill(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const PrefillParams& p,
                   int cap_qtiles, cudaStream_t st) {
  static bool configured = false;
  const int smem = Smem<D>::kTotal > 100 * 1024 ? Smem<D>::kTotal : 100 * 1024;  // cap residency at 2 CTAs / SM (TMEM)
  if (!configured) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(sdag_prefill_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const long long grid = static_cast<long long>(cap_qtiles) * p.H * p.B;
  if (grid <= 0 || grid > 0x7fffffffll) return SDAG_ERR_INVALID;
  sdag_prefill_kernel<D><<<static_cast<unsigned>(grid), 192, smem, st>>>(tq, tk, tv, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" {

#ifdef SDAG_DEBUG_TIMING
int sdag_debug_set_timing(unsigned long long* buf) {
  return static_cast<int>(cudaMemcpyToSymbol(g_timing, &buf, sizeof(buf)));
}
#endif

size_t sdag_prefill_workspace_bytes(int B, int Hkv, int D) {
  if (B <= 0 || Hkv <= 0 || D <= 0) return 0;
  return sizeof(float) * static_cast<size_t>(B) * Hkv * (VSUM_CHUNKS + 1) * D;  // chunk partials + the mean
}

int sdag_prefill_fwd(const void* q, int64_t sq_b, int64_t sq_h, int64_t sq_l, const void* k, int64_t sk_b,
                     int64_t sk_h, int64_t sk_l, const void* v, int64_t sv_b, int64_t sv_h, int64_t sv_l, void* out,
                     int64_t so_b, int64_t so_l, int64_t so_h, int B, int H, int Hkv, int Lq, int Skv, int D,
                     float scaling, const int32_t* blockmaps_dev, size_t blockmap_stride_bytes, int cap_qtiles,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !out || !blockmaps_dev || !workspace) return SDAG_ERR_INVALID;
  if (B <= 0 || H <= 0 || Hkv <= 0 || Lq <= 0 || Skv <= 0 || H % Hkv || cap_qtiles <= 0) return SDAG_ERR_INVALID;
  if (D != 64 && D != 128) return SDAG_ERR_UNSUPPORTED;
  const int64_t strides[] = {sq_b, sq_h, sq_l, sk_b, sk_h, sk_l, sv_b, sv_h, sv_l, so_b, so_l, so_h};
  for (int64_t s : strides)
    if (s < 0 || (s % 8) != 0) return SDAG_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return SDAG_ERR_INVALID;
  if (blockmap_stride_bytes % 4) return SDAG_ERR_INVALID;
  if (workspace_bytes < sdag_prefill_workspace_bytes(B, Hkv, D)) return SDAG_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  CUtensorMap tq, tk, tv;
  // batch stride of a size-1 dim is irrelevant to addressing but must still be a legal stride
  auto bs = [](int64_t s, int64_t fallback) { return s > 0 ? s : fallback; };
  if (make_map_4d(&tq, q, D, Lq, H, B, sq_l, sq_h, bs(sq_b, sq_l * Lq), BM)) return SDAG_ERR_DRIVER;
  if (make_map_4d(&tk, k, D, Skv, Hkv, B, sk_l, sk_h, bs(sk_b, sk_l * Skv), BN)) return SDAG_ERR_DRIVER;
  if (make_map_4d(&tv, v, D, Skv, Hkv, B, sv_l, sv_h, bs(sv_b, sv_l * Skv), BN)) return SDAG_ERR_DRIVER;

  PrefillParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.so_b = so_b, p.so_l = so_l, p.so_h = so_h;
  p.blockmaps = blockmaps_dev;
  p.bm_stride_ints = static_cast<long long>(blockmap_stride_bytes / 4);
  p.vsum = static_cast<const float*>(workspace);
  p.B = B, p.H = H, p.Hkv = Hkv, p.Lq = Lq, p.Skv = Skv;
  p.scale_log2 = scaling * 1.4426950408889634f;

  const dim3 vgrid(VSUM_CHUNKS, Hkv, B);
  if (D == 128) {
    v_colsum_kernel<128><<<vgrid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(v), sv_b, sv_h, sv_l, Skv,
                                                static_cast<float*>(workspace), blockmaps_dev, p.bm_stride_ints);
    v_mean_kernel<128><<<dim3(Hkv, B), 128, 0, st>>>(static_cast<float*>(workspace), B, Hkv, Skv, blockmaps_dev, p.bm_stride_ints);
    SDAG_CUDA_TRY(cudaGetLastError());
    return launch_prefill<128>(tq, tk, tv, p, cap_qtiles, st);
  }
  v_colsum_kernel<64><<<vgrid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(v), sv_b, sv_h, sv_l, Skv,
                                             static_cast<float*>(workspace), blockmaps_dev, p.bm_stride_ints);
  v_mean_kernel<64><<<dim3(Hkv, B), 64, 0, st>>>(static_cast<float*>(workspace), B, Hkv, Skv, blockmaps_dev, p.bm_stride_ints);
  SDAG_CUDA_TRY(cudaGetLastError());
  return launch_prefill<64>(tq, tk, tv, p, cap_qtiles, st);
}

}  // extern "C"
