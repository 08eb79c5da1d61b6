// Exact max-inner-product top-k for sm_100a (path B2): replaces faiss IndexFlatIP.search as called from
// search_index, src/pipeline/retrieval/dense.py:68-84.
//
// faiss computes Q x N fp32 scores with blocked sgemm and keeps a k-heap per query.  Here the Q x N score
// matrix never exists:
//   1. a bf16 tcgen05 GEMM streams the bf16 shadow of the corpus (HBM-resident, read once per batch) against
//      the bf16 queries; its epilogue compares every score with a per-query threshold while it is still in
//      TMEM and appends the rare survivors (score, row) to a small per-query candidate list;
//   2. the corpus is swept in epochs of doubling size; after each epoch a "tighten" kernel raises every
//      query's threshold to (k-th best bf16 score so far) - 2e, where e >= |bf16 score - exact score| is a
//      rigorous bound (Cauchy-Schwarz on the rounding errors).  Any row whose exact score reaches the final
//      exact k-th best therefore always survives: the candidate list is a superset of the true top-k;
//   3. candidates are re-scored exactly in fp32 against the fp32 corpus and ordered (score descending, row
//      ascending) — so ids and scores are those of an exact fp32 search, up to fp32 summation order.
// A query whose list overflows is re-done by an exact fp32 scan (counted in stats).
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int TM = 128;      // queries per tile (UMMA M)
constexpr int TN = 256;      // corpus rows per tile (UMMA N)
constexpr int KB = 64;       // k-block (one 128-byte swizzle row of bf16)
#ifndef SDAG_IP_USE_2CTA
#define SDAG_IP_USE_2CTA 1
#endif
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * KB * 2;
constexpr int B_BYTES = TN * KB * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024;
constexpr int SORT_MAX = 2048;
constexpr int SCAN_SLICES = 32;
constexpr long long IP_FAST_GROWTH_ROWS = 65536;  // the swept prefix grows x4 per epoch below this many rows, x2 above
constexpr uint32_t ROW_EXACT = 0x80000000u;  // Cand.row bit: .score already is the exact fp32 score (rows are < 2^31)
constexpr int SELECT_MAX_K = 32;  // tighten: k-th best by k max-extraction rounds up to this k, bitonic sort above
// |bf16(x).bf16(q) - x.q| <= (2u + u^2) |x||q| with u = 2^-8, plus fp32 accumulation slack
constexpr float ERR_COEF = 0.0078125f + 0.00048828125f;

struct Cand {
  float score;
  int32_t row;
};

struct SearchState {   // device arrays, all indexed by query
  float* thr;          // current emission threshold (bf16-score domain)
  int32_t* cnt;        // candidates appended (may exceed cap -> overflow)
  int32_t* cnt_done;   // cnt as the last tighten left it (stable while the next epoch's filter appends behind it)
  int32_t* overflow;   // sticky flag
  float* qnorm;        // |q|
  Cand* cand;          // [Q][cap]
  int cap;
};

__device__ __forceinline__ uint32_t orderable(float f) {  // monotone float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorderable(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// ------------------------------------------------------------------------------------------------
// fp32 rows -> bf16 rows (+ row norms).  One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, float* __restrict__ max_norm,
                 float* __restrict__ row_norm, long long n, long long n_pad, int d) {
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  float ss = 0.f;
  for (int j = lane * 4; j < d; j += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) v = *reinterpret_cast<const float4*>(x + row * d + j);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&lo);
    w.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + row * d + j) = w;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0 && row < n) {
    const float nr = sqrtf(ss) * 1.0000002f;  // round up: the bound must not be under-estimated
    if (row_norm) row_norm[row] = nr;
    if (max_norm) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(nr));  // nr >= 0: int order == float order
  }
}

// ------------------------------------------------------------------------------------------------
// GEMM + threshold filter
// ------------------------------------------------------------------------------------------------
struct GemmParams {
  SearchState st;
  int Q, n_qt;           // queries, query tiles
  long long row_begin;   // first corpus row of this epoch (multiple of TN)
  long long row_end;     // one past the last
  int n_ct;              // corpus tiles in this epoch
  int kblocks;           // d / 64
  int dump;              // 1: positional dump of every score (first epoch)
};

__global__ void __launch_bounds__(192, 1)
ip_gemm_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const long long n_units = static_cast<long long>(p.n_ct) * p.n_qt;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int ct = static_cast<int>(u / p.n_qt), qt = static_cast<int>(u % p.n_qt);
        const int row0 = static_cast<int>(p.row_begin + static_cast<long long>(ct) * TN);
        for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* sa = smem + s * STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full[s], kb * KB, qt * TM);
          tma_load_2d(sa + A_BYTES, &tmB, &full[s], kb * KB, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = umma_idesc_bf16_f32(TM, TN, 0, 0);
    uint32_t it = 0, ui = 0;
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      const uint32_t a = ui & 1;
      mbar_wait(&acc_empty[a], ((ui >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < KB / 16; ++kk)
            umma_ss(tmem + a * TN, umma_smem_desc_sw128(sa + kk * 32, 16, 1024),
                    umma_smem_desc_sw128(sa + A_BYTES + kk * 32, 16, 1024), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (kb == p.kblocks - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= epilogue: threshold filter (warps 2..5) =======================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ui = 0;
    uint32_t r[32];
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      const int ct = static_cast<int>(u / p.n_qt), qt = static_cast<int>(u % p.n_qt);
      const uint32_t a = ui & 1;
      const int q = qt * TM + row;
      const bool q_ok = q < p.Q;
      const float thr = q_ok ? p.st.thr[q] : INFINITY;
      const long long col0 = p.row_begin + static_cast<long long>(ct) * TN;
      Cand* cq = p.st.cand + static_cast<long long>(q_ok ? q : 0) * p.st.cap;
      mbar_wait(&acc_full[a], (ui >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < TN / 32; ++c) {
        tmem_ld_32x32b_x32(tmem + lane_addr + a * TN + c * 32, r);
        tmem_ld_wait();
        const long long cb = col0 + c * 32;
        if (p.dump) {
          if (q_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const long long rr = cb + i;
              if (rr < p.row_end) cq[rr - p.row_begin] = Cand{__uint_as_float(r[i]), static_cast<int32_t>(rr)};
            }
          }
        } else {
          float mx = __uint_as_float(r[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          if (mx >= thr) {  // rare: fully unrolled so r[] stays in registers
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float s = __uint_as_float(r[i]);
              if (s >= thr && cb + i < p.row_end) {
                const int slot = atomicAdd(&p.st.cnt[q], 1);
                if (slot < p.st.cap) cq[slot] = Cand{s, static_cast<int32_t>(cb + i)};
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// Same GEMM + filter on CTA pairs (cta_group::2): one MMA covers 256 queries x 256 corpus rows, each CTA of the pair holds
// its own 128 queries (A) and its own 128 of the 256 corpus rows (B) per k-block, so the shared-memory operand traffic per
// SM and MMA drops from 128 + 256 to 128 + 128 rows — the SS-mode operand fetch is what keeps the single-CTA kernel at
// ~75 % of the tensor pipe.  The leader (rank 0) issues the MMAs; both CTAs run TMA producers and filter epilogues.
// ------------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6;
constexpr int B2_BYTES = (TN / 2) * KB * 2;
constexpr int STAGE2_BYTES = A_BYTES + B2_BYTES;
constexpr int GEMM2_SMEM = STAGES2 * STAGE2_BYTES + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
ip_gemm_filter_kernel_2cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES2], empty[STAGES2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full[s], 1);   // (leader's copy is the live one) the leader's producer expects the bytes of both CTAs
      mbar_init(&empty[s], 1);  // multicast commit of the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);     // multicast commit of the leader
      mbar_init(&acc_empty[a], 8);    // (leader's copy) one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
  }
  if (warp == 1) {
    tmem_alloc_2cta(&tmem_base_s, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_qt2 = (p.n_qt + 1) / 2;
  const long long n_units = static_cast<long long>(p.n_ct) * n_qt2;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    // warp-uniform control flow, one elected lane issues (no per-lane operands -> no uniform-register broadcast loops)
    const uint32_t lead_full0 = mapa_shared(smem_u32(&full[0]), 0);
    uint32_t it = 0;
    for (long long u = pair; u < n_units; u += n_pairs) {
      const int ct = static_cast<int>(u / n_qt2), qt2 = static_cast<int>(u % n_qt2);
      const int row0 = static_cast<int>(p.row_begin + static_cast<long long>(ct) * TN) + static_cast<int>(rank) * (TN / 2);
      const int q0 = (qt2 * 2 + static_cast<int>(rank)) * TM;
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const int s = it % STAGES2;
        mbar_wait(&empty[s], ((it / STAGES2) & 1) ^ 1);
        // The peer's bytes may complete on the leader's barrier before the leader has announced them (the
        // transaction count goes negative meanwhile); no remote arrive, hence no cluster-scope release per stage.
        const uint32_t lead_full = lead_full0 + s * 8;
        uint8_t* sa = smem + s * STAGE2_BYTES;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[s], 2 * STAGE2_BYTES);
          tma_load_2d_2sm(sa, &tmA, lead_full, kb * KB, q0);
          tma_load_2d_2sm(sa + A_BYTES, &tmBh, lead_full, kb * KB, row0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader only) =======================
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16_f32(2 * TM, TN, 0, 0);
      uint32_t it = 0, ui = 0;
      for (long long u = pair; u < n_units; u += n_pairs, ++ui) {
        const uint32_t a = ui & 1;
        mbar_wait(&acc_empty[a], ((ui >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
          const int s = it % STAGES2;
          mbar_wait(&full[s], (it / STAGES2) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE2_BYTES);
          const uint64_t da = umma_smem_desc_sw128(sa, 16, 1024), db = umma_smem_desc_sw128(sa + A_BYTES, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk)
              umma_ss_2cta(tmem + a * TN, da + ((kk * 32) >> 4), db + ((kk * 32) >> 4), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit_2cta(&empty[s], 3);
            if (kb == p.kblocks - 1) umma_commit_2cta(&acc_full[a], 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================= epilogue: threshold filter (warps 2..5, both CTAs) =======================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ui = 0;
    uint32_t r[32];
    for (long long u = pair; u < n_units; u += n_pairs, ++ui) {
      const int ct = static_cast<int>(u / n_qt2), qt2 = static_cast<int>(u % n_qt2);
      const uint32_t a = ui & 1;
      const int q = (qt2 * 2 + static_cast<int>(rank)) * TM + row;
      const bool q_ok = q < p.Q;
      const float thr = q_ok ? p.st.thr[q] : INFINITY;
      const long long col0 = p.row_begin + static_cast<long long>(ct) * TN;
      Cand* cq = p.st.cand + static_cast<long long>(q_ok ? q : 0) * p.st.cap;
      mbar_wait(&acc_full[a], (ui >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < TN / 32; ++c) {
        tmem_ld_32x32b_x32(tmem + lane_addr + a * TN + c * 32, r);
        tmem_ld_wait();
        const long long cb = col0 + c * 32;
        if (p.dump) {
          if (q_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const long long rr = cb + i;
              if (rr < p.row_end) cq[rr - p.row_begin] = Cand{__uint_as_float(r[i]), static_cast<int32_t>(rr)};
            }
          }
        } else {
          float mx = __uint_as_float(r[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          if (mx >= thr) {  // rare: fully unrolled so r[] stays in registers
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float s = __uint_as_float(r[i]);
              if (s >= thr && cb + i < p.row_end) {
                const int slot = atomicAdd(&p.st.cnt[q], 1);
                if (slot < p.st.cap) cq[slot] = Cand{s, static_cast<int32_t>(cb + i)};
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[a]), 0));  // the leader's barrier gates both CTAs' accumulators
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody may still be reading the peer's shared memory / TMEM when a CTA tears down
  if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// bitonic sort of 64-bit keys in shared memory, descending; n_pow2 <= SORT_MAX, blockDim = 256
// ------------------------------------------------------------------------------------------------
__device__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int pow2_at_least(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// After an epoch: raise thr[q] to (k-th best bf16 score) - 2e and drop candidates below it.
__global__ void __launch_bounds__(256)
tighten_kernel(SearchState st, int k, int force_cnt, const float* __restrict__ max_norm) {
  __shared__ unsigned long long keys[SORT_MAX];
  __shared__ int n_keep_s;
  const int q = blockIdx.x;
  int n = force_cnt >= 0 ? force_cnt : st.cnt[q];
  if (n > st.cap) {
    if (threadIdx.x == 0) st.overflow[q] = 1;
    n = st.cap;
  }
  Cand* cq = st.cand + static_cast<long long>(q) * st.cap;
  const bool select = k <= SELECT_MAX_K;  // small k: k rounds of block-wide max instead of a full sort
  const int P = select ? n : pow2_at_least(n < 1 ? 1 : n);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long key = 0;
    if (i < n) {
      const Cand c = cq[i];
      key = (static_cast<unsigned long long>(orderable(c.score)) << 32) | static_cast<uint32_t>(c.row);
    }
    keys[i] = key;
  }
  if (threadIdx.x == 0) n_keep_s = 0;
  __syncthreads();
  float thr = st.thr[q];
  unsigned long long kth_key = 0;
  if (select) {
    // keys are distinct (one row appears once): the r-th largest is the largest key below the (r-1)-th largest
    __shared__ unsigned long long round_max[2];
    if (n >= k) {
      unsigned long long below = ~0ull;
      for (int r = 0; r < k; ++r) {
        if (threadIdx.x == 0) round_max[r & 1] = 0;
        __syncthreads();
        unsigned long long loc = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const unsigned long long key = keys[i];
          if (key < below && key > loc) loc = key;
        }
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, loc, o);
          loc = other > loc ? other : loc;
        }
        if ((threadIdx.x & 31) == 0 && loc) atomicMax(&round_max[r & 1], loc);
        __syncthreads();
        below = round_max[r & 1];  // the other slot is reset at the top of the next round, after this barrier
      }
      kth_key = below;
    }
  } else {
    bitonic_sort_desc(keys, P);
    if (n >= k) kth_key = keys[k - 1];
  }
  if (n >= k) {
    const float kth = unorderable(static_cast<uint32_t>(kth_key >> 32));
    const float e = ERR_COEF * st.qnorm[q] * (*max_norm);
    thr = fmaxf(thr, kth - 2.f * e);
  }
  if (select) {  // survivors in any order (every consumer re-orders them)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float s = unorderable(static_cast<uint32_t>(keys[i] >> 32));
      if (s >= thr) cq[atomicAdd(&n_keep_s, 1)] = Cand{s, static_cast<int32_t>(keys[i] & 0xffffffffu)};
    }
  } else {
    int keep = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float s = unorderable(static_cast<uint32_t>(keys[i] >> 32));
      if (s >= thr) {
        cq[i] = Cand{s, static_cast<int32_t>(keys[i] & 0xffffffffu)};
        ++keep;
      }
    }
    if (keep) atomicAdd(&n_keep_s, keep);  // sorted descending: the kept entries form a prefix
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st.thr[q] = thr;
    st.cnt[q] = n_keep_s;
    st.cnt_done[q] = n_keep_s;
  }
}

// Same step for the epochs after the first, one WARP per query: by then a list holds a few dozen survivors, and a
// 256-thread block per query with 2k block-wide barriers (the kernel above) spends ~60 us per epoch on 4096 queries —
// the fixed cost that limits the sharded search.  Up to 256 candidates are held in registers (8 per lane); the k-th best
// comes from k rounds of a warp-wide 64-bit max over the keys below the previous round's, survivors are compacted with
// ballots.  Longer lists (rare: clustered data) take the same rounds straight from global memory, chunk by chunk.
constexpr int TW_PER_LANE = 8;
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other > v ? other : v;
  }
  return v;
}
__global__ void __launch_bounds__(256)
tighten_warp_kernel(SearchState st, int Q, int k, const float* __restrict__ max_norm) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  int n = st.cnt[q];
  if (n > st.cap) {
    if (lane == 0) st.overflow[q] = 1;
    n = st.cap;
  }
  Cand* cq = st.cand + static_cast<long long>(q) * st.cap;
  float thr = st.thr[q];
  auto key_of = [](const Cand c) {
    return (static_cast<unsigned long long>(orderable(c.score)) << 32) | static_cast<uint32_t>(c.row);
  };
  if (n <= 32 * TW_PER_LANE) {
    unsigned long long keys[TW_PER_LANE];
#pragma unroll
    for (int j = 0; j < TW_PER_LANE; ++j) {
      const int i = lane + 32 * j;
      keys[j] = i < n ? key_of(cq[i]) : 0ull;
    }
    if (n >= k) {
      unsigned long long below = ~0ull;
      for (int r = 0; r < k; ++r) {  // keys are distinct: the r-th largest is the largest key below the (r-1)-th largest
        unsigned long long loc = 0;
#pragma unroll
        for (int j = 0; j < TW_PER_LANE; ++j)
          if (keys[j] < below && keys[j] > loc) loc = keys[j];
        below = warp_max_u64(loc);
      }
      const float kth = unorderable(static_cast<uint32_t>(below >> 32));
      thr = fmaxf(thr, kth - 2.f * ERR_COEF * st.qnorm[q] * (*max_norm));
    }
    int keep = 0;
    __syncwarp();  // every lane holds its candidates in registers before the list is rewritten in place
#pragma unroll
    for (int j = 0; j < TW_PER_LANE; ++j) {
      const float s = unorderable(static_cast<uint32_t>(keys[j] >> 32));
      const bool ok = (lane + 32 * j) < n && s >= thr;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) cq[keep + __popc(m & ((1u << lane) - 1u))] = Cand{s, static_cast<int32_t>(keys[j] & 0xffffffffu)};
      keep += __popc(m);
    }
    if (lane == 0) {
      st.thr[q] = thr;
      st.cnt[q] = keep;
      st.cnt_done[q] = keep;
    }
    return;
  }
  // long list: the same selection with the keys re-read from global memory (L2) in every round
  if (n >= k) {
    unsigned long long below = ~0ull;
    for (int r = 0; r < k; ++r) {
      unsigned long long loc = 0;
      for (int i = lane; i < n; i += 32) {
        const unsigned long long key = key_of(cq[i]);
        if (key < below && key > loc) loc = key;
      }
      below = warp_max_u64(loc);
    }
    const float kth = unorderable(static_cast<uint32_t>(below >> 32));
    thr = fmaxf(thr, kth - 2.f * ERR_COEF * st.qnorm[q] * (*max_norm));
  }
  int keep = 0;
  for (int base = 0; base < n; base += 32) {  // in-place compaction: chunk c is read whole before its survivors are
    const int i = base + lane;                // written, and they land at or below the chunk's own first index
    Cand c = Cand{-INFINITY, -1};
    if (i < n) c = cq[i];
    const bool ok = i < n && c.score >= thr;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    __syncwarp();
    if (ok) cq[keep + __popc(m & ((1u << lane) - 1u))] = c;
    keep += __popc(m);
    __syncwarp();
  }
  if (lane == 0) {
    st.thr[q] = thr;
    st.cnt[q] = keep;
    st.cnt_done[q] = keep;
  }
}

// ------------------------------------------------------------------------------------------------
// exact fp32 dot of one corpus row with the query held in shared memory (one warp)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_dot_f32(const float* __restrict__ row, const float* sq, int d, int lane) {
  float s = 0.f;
  for (int j = lane * 4; j < d; j += 128) {
    const float4 a = *reinterpret_cast<const float4*>(row + j);
    const float4 b = *reinterpret_cast<const float4*>(sq + j);
    s = fmaf(a.x, b.x, s);
    s = fmaf(a.y, b.y, s);
    s = fmaf(a.z, b.z, s);
    s = fmaf(a.w, b.w, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
__device__ __forceinline__ float warp_dot_bf16(const __nv_bfloat16* __restrict__ row, const float* sq, int d, int lane) {
  float s = 0.f;
  for (int j = lane * 4; j < d; j += 128) {
    const uint2 w = *reinterpret_cast<const uint2*>(row + j);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
    const float4 b = *reinterpret_cast<const float4*>(sq + j);
    s = fmaf(a.x, b.x, s);
    s = fmaf(a.y, b.y, s);
    s = fmaf(c.x, b.z, s);
    s = fmaf(c.y, b.w, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Overflowed queries only: exact scan of a slice of the corpus, slice-local top-k -> cand[q][slice*k ..].
__global__ void __launch_bounds__(256)
exact_scan_kernel(SearchState st, int Q, const float* __restrict__ xq, const float* __restrict__ corpus_f32,
                  const __nv_bfloat16* __restrict__ corpus_bf16, long long N, int d, int k) {
  extern __shared__ __align__(16) uint8_t dyn[];
  float* sq = reinterpret_cast<float*>(dyn);  // [d]
  __shared__ unsigned long long keys[SORT_MAX];
  __shared__ int fill_s;
  __shared__ float thr_s;
  // grid = (slices, query slots): a block walks the queries of its slot and works only on the overflowed ones, so the
  // common case (no overflow anywhere) costs a few thousand flag reads instead of Q x slices block launches (72 us at
  // Q = 4096 when every block merely looked at its flag and left)
  const int slice = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int q = blockIdx.y; q < Q; q += gridDim.y) {
  if (!st.overflow[q]) continue;  // block-uniform
  __syncthreads();                // the previous query of this block is completely written out
  for (int j = tid; j < d; j += blockDim.x) sq[j] = xq[static_cast<long long>(q) * d + j];
  if (tid == 0) {
    fill_s = 0;
    thr_s = -INFINITY;
  }
  __syncthreads();
  const long long per = (N + SCAN_SLICES - 1) / SCAN_SLICES;
  const long long lo = slice * per, hi = (lo + per < N) ? lo + per : N;
  constexpr int BATCH = 8 * 64;  // rows per round (64 per warp)
  for (long long base = lo; base < hi; base += BATCH) {
    for (int t = 0; t < 64; ++t) {
      const long long rr = base + warp * 64 + t;
      if (rr < hi) {
        const float s = corpus_f32 ? warp_dot_f32(corpus_f32 + rr * d, sq, d, lane)
                                   : warp_dot_bf16(corpus_bf16 + rr * d, sq, d, lane);
        if (lane == 0 && s >= thr_s) {
          const int slot = atomicAdd(&fill_s, 1);  // <= 512 new entries per round + k kept: fits SORT_MAX
          keys[slot] = (static_cast<unsigned long long>(orderable(s)) << 32) | (0xffffffffu - static_cast<uint32_t>(rr));
        }
      }
    }
    __syncthreads();
    const int n = fill_s;
    if (n > k) {
      const int P = pow2_at_least(n);
      for (int i = n + tid; i < P; i += blockDim.x) keys[i] = 0;
      __syncthreads();
      bitonic_sort_desc(keys, P);
      if (tid == 0) {
        fill_s = k;
        thr_s = unorderable(static_cast<uint32_t>(keys[k - 1] >> 32));
      }
    }
    __syncthreads();
  }
  // slice result: exact scores, rows recovered from the key
  const int n = fill_s < k ? fill_s : k;
  if (fill_s <= k) {  // not sorted yet (fewer than k rows in the slice) — order is irrelevant here
  }
  Cand* out = st.cand + static_cast<long long>(q) * st.cap + slice * k;
  for (int i = tid; i < k; i += blockDim.x) {
    if (i < n)
      out[i] = Cand{unorderable(static_cast<uint32_t>(keys[i] >> 32)),
                    static_cast<int32_t>(0xffffffffu - static_cast<uint32_t>(keys[i] & 0xffffffffu))};
    else
      out[i] = Cand{-FLT_MAX, -1};
  }
  if (tid == 0 && slice == 0) st.cnt[q] = SCAN_SLICES * k;
  }
}

// Exact re-scoring ahead of time: the candidates that survived the tighten before the LAST epoch (about seven eighths of
// the final list) get their exact fp32 score now, on a second stream, while the last epoch's GEMM — which leaves HBM ~95 %
// idle — sweeps the rest of the corpus and appends behind cnt_done.  The exact score replaces the bf16 one in place and
// the row is tagged ROW_EXACT; the last tighten may treat such a score like any other (|stored - exact| <= e holds for
// both kinds, which is all its bound needs), and rescore_topk below only computes what is still untagged.  Re-scoring is
// 4 KB of row per candidate, ~100 candidates per query: 275 us per 4096-query search when it ran after the sweep, whatever
// the shard size (profiles/r02_topk_dram.md).
__global__ void __launch_bounds__(256)
rescore_partial_kernel(SearchState st, int Q, const float* __restrict__ xq, const float* __restrict__ corpus_f32,
                       const __nv_bfloat16* __restrict__ corpus_bf16, int d) {
  // A small persistent grid (two blocks per SM, 4 KB of shared memory each) walks the queries: it must never hold the
  // thread slots or the shared memory the GEMM's CTAs (193 KB, one per SM) need — a first version with a block per query
  // flooded the SMs ahead of the GEMM and cost more than it saved.
  extern __shared__ __align__(16) uint8_t dyn[];
  float* sq = reinterpret_cast<float*>(dyn);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
    __syncthreads();  // sq of the previous query is no longer read
    if (st.overflow[q]) continue;  // redone by exact_scan anyway (block-uniform)
    int n = st.cnt_done[q];
    if (n > st.cap) n = st.cap;
    if (n == 0) continue;
    for (int j = tid; j < d; j += blockDim.x) sq[j] = xq[static_cast<long long>(q) * d + j];
    __syncthreads();
    Cand* cq = st.cand + static_cast<long long>(q) * st.cap;
    for (int i = warp; i < n; i += blockDim.x / 32) {
      const Cand c = cq[i];
      if (c.row < 0) continue;  // already exact
      const float s = corpus_f32 ? warp_dot_f32(corpus_f32 + static_cast<long long>(c.row) * d, sq, d, lane)
                                 : warp_dot_bf16(corpus_bf16 + static_cast<long long>(c.row) * d, sq, d, lane);
      if (lane == 0) cq[i] = Cand{s, static_cast<int32_t>(static_cast<uint32_t>(c.row) | ROW_EXACT)};
    }
  }
}

// Exact re-scoring of the candidates and the final ordered top-k.
__global__ void __launch_bounds__(256)
rescore_topk_kernel(SearchState st, const float* __restrict__ xq, const float* __restrict__ corpus_f32,
                    const __nv_bfloat16* __restrict__ corpus_bf16, int d, int k, long long id_offset,
                    float* __restrict__ out_scores, long long* __restrict__ out_ids, int32_t* __restrict__ stats) {
  extern __shared__ __align__(16) uint8_t dyn[];
  float* sq = reinterpret_cast<float*>(dyn);
  __shared__ unsigned long long keys[SORT_MAX];
  const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int j = tid; j < d; j += blockDim.x) sq[j] = xq[static_cast<long long>(q) * d + j];
  int n = st.cnt[q];
  if (n > st.cap) n = st.cap;
  const int P = pow2_at_least(n < 1 ? 1 : n);
  for (int i = n + tid; i < P; i += blockDim.x) keys[i] = 0;
  __syncthreads();
  const Cand* cq = st.cand + static_cast<long long>(q) * st.cap;
  for (int i = warp; i < n; i += blockDim.x / 32) {
    const Cand c = cq[i];
    unsigned long long key = 0;
    if (c.row != -1) {  // -1: padding written by exact_scan
      const uint32_t row = static_cast<uint32_t>(c.row) & ~ROW_EXACT;
      float s = c.score;  // tagged: scored by rescore_partial_kernel while the last epoch was running
      if (c.row >= 0)
        s = corpus_f32 ? warp_dot_f32(corpus_f32 + static_cast<long long>(row) * d, sq, d, lane)
                       : warp_dot_bf16(corpus_bf16 + static_cast<long long>(row) * d, sq, d, lane);
      // descending key order = score descending, then row ascending
      key = (static_cast<unsigned long long>(orderable(s)) << 32) | (0xffffffffu - row);
    }
    if (lane == 0) keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(keys, P);
  for (int i = tid; i < k; i += blockDim.x) {
    float s = -FLT_MAX;
    long long id = -1;
    if (i < n && keys[i] != 0) {
      s = unorderable(static_cast<uint32_t>(keys[i] >> 32));
      id = static_cast<long long>(0xffffffffu - static_cast<uint32_t>(keys[i] & 0xffffffffu)) + id_offset;
    }
    out_scores[static_cast<long long>(q) * k + i] = s;
    out_ids[static_cast<long long>(q) * k + i] = id;
  }
  if (stats && tid == 0) {
    atomicAdd(&stats[1], n);
    if (st.overflow[q]) atomicAdd(&stats[0], 1);
  }
}

__global__ void init_state_kernel(SearchState st, int Q, int32_t* stats) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < Q) {
    st.thr[q] = -INFINITY;
    st.cnt[q] = 0;
    st.cnt_done[q] = 0;
    st.overflow[q] = 0;
  }
  if (q < 4 && stats) stats[q] = 0;
}

// merge of per-shard top-k lists: order (score desc, id asc), id < 0 (padding) last
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, int G, int Q, int k,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids) {
  extern __shared__ __align__(16) uint8_t dyn[];
  const int n = G * k;
  float* s = reinterpret_cast<float*>(dyn);
  long long* id = reinterpret_cast<long long*>(dyn + ((n * 4 + 15) & ~15));
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int g = i / k, j = i % k;
    s[i] = scores[(static_cast<long long>(g) * Q + q) * k + j];
    id[i] = ids[(static_cast<long long>(g) * Q + q) * k + j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float si = s[i];
    const long long ii = id[i];
    const bool pad_i = ii < 0;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = s[j];
      const long long ij = id[j];
      const bool pad_j = ij < 0;
      bool better;
      if (pad_i != pad_j)
        better = pad_i;  // real entries before padding
      else if (sj != si)
        better = sj > si;
      else if (ij != ii)
        better = ij < ii;
      else
        better = j < i;
      rank += better;
    }
    if (rank < k) {
      out_scores[static_cast<long long>(q) * k + rank] = pad_i ? -FLT_MAX : si;
      out_ids[static_cast<long long>(q) * k + rank] = pad_i ? -1 : ii;
    }
  }
}

// Local results as 12-byte records {f32 score bits, id lo, id hi} so that the shards exchange ONE buffer (one all-gather
// instead of one for the scores and one for the ids), and the merge of the gathered records [G][Q][k].
__global__ void __launch_bounds__(256)
topk_pack_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, long long n, int32_t* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const long long id = ids[i];
  out[3 * i] = __float_as_int(scores[i]);
  out[3 * i + 1] = static_cast<int32_t>(id & 0xffffffffll);
  out[3 * i + 2] = static_cast<int32_t>(id >> 32);
}

__global__ void __launch_bounds__(128)
topk_merge_packed_kernel(const int32_t* __restrict__ rec, int G, int Q, int k, float* __restrict__ out_scores,
                         long long* __restrict__ out_ids) {
  extern __shared__ __align__(16) uint8_t dyn[];
  const int n = G * k;
  float* s = reinterpret_cast<float*>(dyn);
  long long* id = reinterpret_cast<long long*>(dyn + ((n * 4 + 15) & ~15));
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int g = i / k, j = i % k;
    const int32_t* r = rec + ((static_cast<long long>(g) * Q + q) * k + j) * 3;
    s[i] = __int_as_float(r[0]);
    id[i] = static_cast<long long>(static_cast<uint32_t>(r[1])) | (static_cast<long long>(r[2]) << 32);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float si = s[i];
    const long long ii = id[i];
    const bool pad_i = ii < 0;
    int rank = 0;
    for (int j = 0; j < n; ++j) {  // same order rule as topk_merge_kernel: real before padding, score desc, id asc
      const float sj = s[j];
      const long long ij = id[j];
      const bool pad_j = ij < 0;
      bool better;
      if (pad_i != pad_j)
        better = pad_i;
      else if (sj != si)
        better = sj > si;
      else if (ij != ii)
        better = ij < ii;
      else
        better = j < i;
      rank += better;
    }
    if (rank < k) {
      out_scores[static_cast<long long>(q) * k + rank] = pad_i ? -FLT_MAX : si;
      out_ids[static_cast<long long>(q) * k + rank] = pad_i ? -1 : ii;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The exchange step of the sharded search as ONE kernel over peer memory (NVLink / NVSwitch), replacing
// topk_pack -> ncclAllGather -> topk_merge_packed: every rank stores its packed local top-k records straight into the
// receive buffers of ALL ranks (remote stores through the peer mappings of a symmetric allocation), raises one flag per
// (source rank, block) in the receiver's flag array with release semantics at system scope, waits for the G flags of its
// own block and merges them.  Block b of every rank only ever waits for block b of the other ranks, and those raise
// their flags before they wait for anything, so the kernel cannot deadlock whatever order the blocks are scheduled in;
// flags carry the call's epoch (they are never reset), the receive area is double-buffered by epoch parity (a rank can
// be at most one call ahead of a peer: it needs that peer's flag of call n + 1 to get past n + 1).
// Layout of a rank's symmetric buffer: uint32 flags[G][n_blocks] | pad to 256 B | int32 recv[2][G][Q_cap][k][3].
// ------------------------------------------------------------------------------------------------
constexpr int XCHG_QB = 8;  // queries per block
struct ExchangeParams {
  const float* scores;
  const long long* ids;
  float* out_scores;
  long long* out_ids;
  unsigned long long* peers;  // device array of G base addresses (this rank's mapping of every rank's buffer)
  long long recv_off;         // byte offset of recv[parity] inside a buffer
  int Q, Q_cap, k, rank, G, n_blocks;
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(128)
topk_exchange_merge_kernel(const ExchangeParams p) {
  extern __shared__ __align__(16) uint8_t dyn[];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = b * XCHG_QB;
  const int nq = min(XCHG_QB, p.Q - q0);
  const int words = nq * p.k * 3;  // int32 words this block contributes per destination
  // phase 1: my records of these queries -> slot [rank] of every rank's receive area (self included)
  for (int peer = 0; peer < p.G; ++peer) {
    int32_t* dst = reinterpret_cast<int32_t*>(p.peers[peer] + p.recv_off) +
                   (static_cast<long long>(p.rank) * p.Q_cap + q0) * p.k * 3;
    for (int i = tid; i < words; i += blockDim.x) {
      const int rec = i / 3, w = i - rec * 3;
      const long long src = static_cast<long long>(q0) * p.k + rec;
      const long long id = p.ids[src];
      dst[i] = w == 0 ? __float_as_int(p.scores[src]) : (w == 1 ? static_cast<int32_t>(id & 0xffffffffll) : static_cast<int32_t>(id >> 32));
    }
  }
  __threadfence_system();  // every thread's remote stores are ordered before the flags below
  __syncthreads();
  // phase 2: tell every rank that slot [rank] of block b holds call `epoch`
  if (tid < p.G) {
    uint32_t* flags = reinterpret_cast<uint32_t*>(p.peers[tid]);
    st_release_sys_u32(flags + static_cast<long long>(p.rank) * p.n_blocks + b, p.epoch);
  }
  // phase 3: wait for block b of every rank
  if (tid < p.G) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(p.peers[p.rank]) + static_cast<long long>(tid) * p.n_blocks + b;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys_u32(mine) - p.epoch) < 0) {
      if (clock64() - t0 > SDAG_SPIN_TRAP_CYCLES) __trap();  // a rank that never calls: fail instead of hanging the box
      __nanosleep(64);
    }
  }
  __syncthreads();
  // phase 4: merge G x k records per query (a warp per query): order (score desc, id asc), padding ids (-1) last
  const int n = p.G * p.k;
  float* s_all = reinterpret_cast<float*>(dyn);
  long long* id_all = reinterpret_cast<long long*>(dyn + ((4 * n * 4 + 15) & ~15));
  float* s = s_all + warp * n;
  long long* id = id_all + warp * n;
  const int32_t* recv = reinterpret_cast<const int32_t*>(p.peers[p.rank] + p.recv_off);
  for (int qq = warp; qq < nq; qq += 4) {
    const int q = q0 + qq;
    for (int i = lane; i < n; i += 32) {
      const int g = i / p.k, j = i - g * p.k;
      const int32_t* r = recv + ((static_cast<long long>(g) * p.Q_cap + q) * p.k + j) * 3;
      // written by other GPUs: read past L1 (the same addresses held the records of two calls ago)
      s[i] = __int_as_float(__ldcv(r));
      id[i] = static_cast<long long>(static_cast<uint32_t>(__ldcv(r + 1))) | (static_cast<long long>(__ldcv(r + 2)) << 32);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const float si = s[i];
      const long long ii = id[i];
      const bool pad_i = ii < 0;
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float sj = s[j];
        const long long ij = id[j];
        const bool pad_j = ij < 0;
        bool better;
        if (pad_i != pad_j)
          better = pad_i;
        else if (sj != si)
          better = sj > si;
        else if (ij != ii)
          better = ij < ii;
        else
          better = j < i;
        rank += better;
      }
      if (rank < p.k) {
        p.out_scores[static_cast<long long>(q) * p.k + rank] = pad_i ? -FLT_MAX : si;
        p.out_ids[static_cast<long long>(q) * p.k + rank] = pad_i ? -1 : ii;
      }
    }
    __syncwarp();
  }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WorkspaceLayout {
  size_t xq_bf16, qnorm, thr, cnt, cnt_done, overflow, cand, total;
  int q_pad;
};
inline WorkspaceLayout plan_workspace(int Q, int d, int cap) {
  WorkspaceLayout w;
  w.q_pad = (Q + TM - 1) / TM * TM;
  size_t off = 0;
  w.xq_bf16 = off, off = align_up(off + static_cast<size_t>(w.q_pad) * d * 2, 256);
  w.qnorm = off, off = align_up(off + static_cast<size_t>(w.q_pad) * 4, 256);
  w.thr = off, off = align_up(off + static_cast<size_t>(Q) * 4, 256);
  w.cnt = off, off = align_up(off + static_cast<size_t>(Q) * 4, 256);
  w.cnt_done = off, off = align_up(off + static_cast<size_t>(Q) * 4, 256);
  w.overflow = off, off = align_up(off + static_cast<size_t>(Q) * 4, 256);
  w.cand = off, off = align_up(off + static_cast<size_t>(Q) * cap * sizeof(Cand), 256);
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int ip_cast_rows_bf16(const float* x, void* x_bf16, float* max_norm_dev, int64_t n, int d, void* stream) {
  if (!x || !x_bf16 || n < 0 || d <= 0 || (d % 4)) return SDAG_ERR_INVALID;
  if (n == 0) return SDAG_OK;
  const long long blocks = (n * 32 + 255) / 256;
  cast_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(x_bf16), max_norm_dev, nullptr, n, n, d);
  return static_cast<int>(cudaGetLastError());
}

size_t ip_search_workspace_bytes(int Q, int d, int k, int cand_cap) {
  if (Q <= 0 || d <= 0 || k <= 0 || cand_cap <= 0) return 0;
  return plan_workspace(Q, d, cand_cap).total;
}

int ip_search(const float* xq, int Q, const float* corpus_f32, const void* corpus_bf16, int64_t N, int d,
              int64_t id_offset, const float* max_norm_dev, int k, float* out_scores, int64_t* out_ids,
              int32_t* stats_dev, void* workspace, size_t workspace_bytes, int cand_cap, void* stream) {
  if (!xq || !corpus_bf16 || !max_norm_dev || !out_scores || !out_ids || !workspace) return SDAG_ERR_INVALID;
  if (Q <= 0 || N < 0 || d <= 0 || k <= 0) return SDAG_ERR_INVALID;
  if (d % KB || d > 8192) return SDAG_ERR_UNSUPPORTED;       // whole 128-byte swizzle rows; query fits in smem
  if (cand_cap > SORT_MAX || cand_cap < 256 || (cand_cap & (cand_cap - 1))) return SDAG_ERR_INVALID;
  if (k > cand_cap / SCAN_SLICES || N > 0x7fffff00ll) return SDAG_ERR_UNSUPPORTED;
  const WorkspaceLayout w = plan_workspace(Q, d, cand_cap);
  if (workspace_bytes < w.total) return SDAG_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  SearchState state;
  state.thr = reinterpret_cast<float*>(ws + w.thr);
  state.cnt = reinterpret_cast<int32_t*>(ws + w.cnt);
  state.cnt_done = reinterpret_cast<int32_t*>(ws + w.cnt_done);
  state.overflow = reinterpret_cast<int32_t*>(ws + w.overflow);
  state.qnorm = reinterpret_cast<float*>(ws + w.qnorm);
  state.cand = reinterpret_cast<Cand*>(ws + w.cand);
  state.cap = cand_cap;
  __nv_bfloat16* xq_bf16 = reinterpret_cast<__nv_bfloat16*>(ws + w.xq_bf16);

  init_state_kernel<<<(Q + 255) / 256, 256, 0, st>>>(state, Q, stats_dev);
  {
    const long long blocks = (static_cast<long long>(w.q_pad) * 32 + 255) / 256;
    cast_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(xq, xq_bf16, nullptr, state.qnorm, Q, w.q_pad, d);
  }
  SDAG_CUDA_TRY(cudaGetLastError());

  if (N > 0) {
    CUtensorMap tmA, tmB;
    {
      uint64_t dims[2] = {static_cast<uint64_t>(d), static_cast<uint64_t>(w.q_pad)};
      uint64_t str[1] = {static_cast<uint64_t>(d) * 2};
      uint32_t box[2] = {KB, TM};
      if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xq_bf16, dims, str, box)) return SDAG_ERR_DRIVER;
    }
    {
      uint64_t dims[2] = {static_cast<uint64_t>(d), static_cast<uint64_t>(N)};
      uint64_t str[1] = {static_cast<uint64_t>(d) * 2};
      uint32_t box[2] = {KB, TN};
      if (make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, corpus_bf16, dims, str, box)) return SDAG_ERR_DRIVER;
    }
    CUtensorMap tmBh;  // half corpus tiles for the CTA-pair kernel
    {
      uint64_t dims[2] = {static_cast<uint64_t>(d), static_cast<uint64_t>(N)};
      uint64_t str[1] = {static_cast<uint64_t>(d) * 2};
      uint32_t box[2] = {KB, TN / 2};
      if (make_tmap(&tmBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, corpus_bf16, dims, str, box)) return SDAG_ERR_DRIVER;
    }
    static bool configured[SDAG_MAX_DEVICES] = {};
    const int dev = device_slot();
    if (dev < 0) return SDAG_ERR_UNSUPPORTED;
    if (!configured[dev]) {
      SDAG_CUDA_TRY(cudaFuncSetAttribute(ip_gemm_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
      SDAG_CUDA_TRY(cudaFuncSetAttribute(ip_gemm_filter_kernel_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM));
      configured[dev] = true;
    }
    const int n_sm = device_sm_count();
    GemmParams gp;
    gp.st = state;
    gp.Q = Q;
    gp.n_qt = w.q_pad / TM;
    gp.kblocks = d / KB;
    // first epoch: positional dump of up to cap/2 rows (multiple of TN); then the swept prefix grows x4 per epoch while
    // the epochs are too small to fill the machine (below 64 k rows the launch + tighten cost dominates, and a looser
    // threshold only means ~100 survivors per query), x2 afterwards (survivors make the filter epilogue take its slow path)
    long long begin = 0;
    long long span = cand_cap / 2 / TN * TN;
    if (span < TN) span = TN;
    bool first = true;
    int epochs = 0;
    // second stream for the early exact re-scoring (rescore_partial_kernel): forked before the last epoch's GEMM, joined
    // before the tighten that follows it.  One per device, created on first use; calls are expected from one host thread.
    static cudaStream_t aux[SDAG_MAX_DEVICES] = {};
    static cudaEvent_t ev_fork[SDAG_MAX_DEVICES] = {}, ev_join[SDAG_MAX_DEVICES] = {};
    if (!aux[dev]) {
      SDAG_CUDA_TRY(cudaStreamCreateWithFlags(&aux[dev], cudaStreamNonBlocking));
      SDAG_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
      SDAG_CUDA_TRY(cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming));
    }
    static int early_env = -1;
    if (early_env < 0) {
      // Measured and left OFF (same box, 1.25 M-row shard, Q = 4096): 8.49-8.59 ms without, 8.68-8.72 ms with a block per
      // query on the second stream (its 4096 blocks take the thread slots ahead of the GEMM's CTAs), 8.94-8.96 ms with two
      // persistent blocks per SM (latency-bound: it outlasts the last epoch and the join waits); the chip runs at its power
      // cap during the GEMM, so the HBM traffic moved under it is not free either.  SDAG_IP_EARLY_RESCORE=1 turns it on.
      const char* e = getenv("SDAG_IP_EARLY_RESCORE");
      early_env = (e && atoi(e)) ? 1 : 0;
    }
    static long long fast_growth_rows = -1;  // SDAG_IP_FAST_GROWTH_ROWS: tuning knob for the epoch schedule
    if (fast_growth_rows < 0) {
      const char* e = getenv("SDAG_IP_FAST_GROWTH_ROWS");
      fast_growth_rows = (e && atoll(e) > 0) ? atoll(e) : IP_FAST_GROWTH_ROWS;
    }
    while (begin < N) {
      long long end = begin + span < N ? begin + span : N;
      const bool fork = early_env && !first && end == N;  // last epoch, and a tighten has already run
      if (fork) {
        SDAG_CUDA_TRY(cudaEventRecord(ev_fork[dev], st));
        SDAG_CUDA_TRY(cudaStreamWaitEvent(aux[dev], ev_fork[dev], 0));
      }
      gp.row_begin = begin;
      gp.row_end = end;
      gp.n_ct = static_cast<int>((end - begin + TN - 1) / TN);
      gp.dump = first ? 1 : 0;
      const long long units = static_cast<long long>(gp.n_ct) * gp.n_qt;
      // CTA pairs (256 x 256 MMAs) once a batch has at least two query tiles; the single-CTA kernel otherwise
      if (SDAG_IP_USE_2CTA && gp.n_qt >= 2) {
        const long long units2 = static_cast<long long>(gp.n_ct) * ((gp.n_qt + 1) / 2);
        const long long pairs = units2 < n_sm / 2 ? units2 : n_sm / 2;
        ip_gemm_filter_kernel_2cta<<<static_cast<unsigned>(2 * pairs), 192, GEMM2_SMEM, st>>>(tmA, tmBh, gp);
      } else {
        const int grid = static_cast<int>(units < n_sm ? units : n_sm);
        ip_gemm_filter_kernel<<<grid, 192, GEMM_SMEM, st>>>(tmA, tmB, gp);
      }
      if (fork) {  // launched after the GEMM so that the GEMM's CTAs take their SMs first; the re-scoring blocks fit beside them
        rescore_partial_kernel<<<Q < 2 * n_sm ? Q : 2 * n_sm, 256, static_cast<size_t>(d) * 4, aux[dev]>>>(
            state, Q, xq, corpus_f32, static_cast<const __nv_bfloat16*>(corpus_bf16), d);
        SDAG_CUDA_TRY(cudaEventRecord(ev_join[dev], aux[dev]));
        SDAG_CUDA_TRY(cudaStreamWaitEvent(st, ev_join[dev], 0));
      }
      if (first)  // every list holds the whole dumped span: block per query
        tighten_kernel<<<Q, 256, 0, st>>>(state, k, static_cast<int>(end - begin), max_norm_dev);
      else        // a few dozen survivors per list: warp per query
        tighten_warp_kernel<<<(Q + 7) / 8, 256, 0, st>>>(state, Q, k, max_norm_dev);
      SDAG_CUDA_TRY(cudaGetLastError());
      first = false;
      begin = end;
      span = begin < fast_growth_rows ? 3 * begin : begin;
      ++epochs;
    }
    (void)epochs;
    exact_scan_kernel<<<dim3(SCAN_SLICES, Q < 64 ? Q : 64), 256, static_cast<size_t>(d) * 4, st>>>(
        state, Q, xq, corpus_f32, static_cast<const __nv_bfloat16*>(corpus_bf16), N, d, k);
  }
  rescore_topk_kernel<<<Q, 256, static_cast<size_t>(d) * 4, st>>>(state, xq, corpus_f32,
                                                                 static_cast<const __nv_bfloat16*>(corpus_bf16), d, k,
                                                                 id_offset, out_scores, reinterpret_cast<long long*>(out_ids),
                                                                 stats_dev);
  return static_cast<int>(cudaGetLastError());
}

int topk_pack(const float* scores, const int64_t* ids, int64_t n, void* records, void* stream) {
  if (!scores || !ids || !records || n < 0) return SDAG_ERR_INVALID;
  if (n == 0) return SDAG_OK;
  topk_pack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, reinterpret_cast<const long long*>(ids), n, static_cast<int32_t*>(records));
  return static_cast<int>(cudaGetLastError());
}

int topk_merge_packed(const void* records, int G, int Q, int k, float* out_scores, int64_t* out_ids, void* stream) {
  if (!records || !out_scores || !out_ids || G <= 0 || Q <= 0 || k <= 0) return SDAG_ERR_INVALID;
  const size_t n = static_cast<size_t>(G) * k;
  const size_t smem = ((n * 4 + 15) & ~size_t(15)) + n * 8;
  if (smem > 48 * 1024) return SDAG_ERR_UNSUPPORTED;
  topk_merge_packed_kernel<<<Q, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const int32_t*>(records), G, Q, k, out_scores, reinterpret_cast<long long*>(out_ids));
  return static_cast<int>(cudaGetLastError());
}

size_t topk_exchange_buffer_bytes(int G, int Q_cap, int k) {
  if (G <= 0 || Q_cap <= 0 || k <= 0) return 0;
  const size_t n_blocks = (static_cast<size_t>(Q_cap) + XCHG_QB - 1) / XCHG_QB;
  return align_up(static_cast<size_t>(G) * n_blocks * 4, 256) + 2 * static_cast<size_t>(G) * Q_cap * k * 12;
}

int topk_exchange_merge(const float* scores, const int64_t* ids, int Q, int Q_cap, int k, int rank, int G,
                        const uint64_t* peer_buffers_dev, uint32_t epoch, float* out_scores, int64_t* out_ids, void* stream) {
  if (!scores || !ids || !peer_buffers_dev || !out_scores || !out_ids) return SDAG_ERR_INVALID;
  if (Q <= 0 || Q > Q_cap || k <= 0 || G <= 0 || G > 128 || rank < 0 || rank >= G || epoch == 0) return SDAG_ERR_INVALID;
  const size_t n = static_cast<size_t>(G) * k;
  const size_t smem = ((4 * n * 4 + 15) & ~size_t(15)) + 4 * n * 8;
  if (smem > 48 * 1024) return SDAG_ERR_UNSUPPORTED;
  ExchangeParams p;
  p.scores = scores;
  p.ids = reinterpret_cast<const long long*>(ids);
  p.out_scores = out_scores;
  p.out_ids = reinterpret_cast<long long*>(out_ids);
  p.peers = reinterpret_cast<unsigned long long*>(const_cast<uint64_t*>(peer_buffers_dev));
  p.Q = Q, p.Q_cap = Q_cap, p.k = k, p.rank = rank, p.G = G;
  p.n_blocks = (Q_cap + XCHG_QB - 1) / XCHG_QB;
  p.recv_off = static_cast<long long>(align_up(static_cast<size_t>(G) * p.n_blocks * 4, 256)) +
               static_cast<long long>(epoch & 1) * G * Q_cap * k * 12;
  p.epoch = epoch;
  const int blocks = (Q + XCHG_QB - 1) / XCHG_QB;
  topk_exchange_merge_kernel<<<blocks, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return static_cast<int>(cudaGetLastError());
}

int topk_merge(const float* scores, const int64_t* ids, int G, int Q, int k, float* out_scores, int64_t* out_ids,
               void* stream) {
  if (!scores || !ids || !out_scores || !out_ids || G <= 0 || Q <= 0 || k <= 0) return SDAG_ERR_INVALID;
  const size_t n = static_cast<size_t>(G) * k;
  const size_t smem = ((n * 4 + 15) & ~size_t(15)) + n * 8;
  if (smem > 48 * 1024) return SDAG_ERR_UNSUPPORTED;
  topk_merge_kernel<<<Q, 128, smem, static_cast<cudaStream_t>(stream)>>>(scores, reinterpret_cast<const long long*>(ids), G,
                                                                        Q, k, out_scores, reinterpret_cast<long long*>(out_ids));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
