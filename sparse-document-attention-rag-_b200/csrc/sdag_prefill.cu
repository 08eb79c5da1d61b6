// SDAG prefill attention for sm_100a: block-sparse FlashAttention-style forward over the tile lists of
// sdag_blockmap.cu.  Replaces the masked SDPA call the reference reaches from SDAG.py:176
// (HF modeling_llama.py:199-221 / integrations/sdpa_attention.py:40-104); also the CARG causal prefill and, under a
// bidirectional block map, the encoder's attention.
//
// Persistent kernel, 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs; work items = (query tile, head,
// prompt), longest first, pulled from a global counter.
//   warps 0-3   softmax: one query row per thread (TMEM lane), fp32 online softmax with lazy rescaling, P written back
//               over S as packed bf16; the row sums of a finished item go to the epilogue warpgroup through smem
//   warp 4      scheduler + TMA producer: publishes decoded items through smem, loads Q once per item and K as 64-row
//               halves through a two-slot ring
//   warp 5      tcgen05 issuer: S = Q K_g^T (SS, both K-major) -> TMEM; O += P_g V_g (TS: P in TMEM, V MN-major straight
//               from its natural [kv][d] layout); issued under elect.sync from warp-uniform descriptors
//   warp 6      TMA producer for V (64-row halves, two-slot ring)
//   warps 8-11  epilogue: O out of TMEM (releasing it for the next item), O / l -> bf16 -> staging -> coalesced stores,
//               concurrently with the softmax warps' next item
// TMEM columns: [0,64) / [64,128) score buffers S0 / S1 (P aliased onto the first 32 columns of each), [128,128+D) O.
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
#ifndef SDAG_PAIR_POLY
#define SDAG_PAIR_POLY 1  // pair kernel: half of the exponentials as a cubic on the FMA pipe (0: all on the exp unit)
#endif
#ifndef SDAG_PREFILL_PAIR_DEFAULT
#define SDAG_PREFILL_PAIR_DEFAULT 0  // which kernel eligible shapes take when neither a flag nor SDAG_ATTN_PAIR says
#endif

struct PrefillParams {
  __nv_bfloat16* out;
  long long so_b, so_l, so_h;
  const int32_t* blockmaps;
  long long bm_stride_ints;
  const float* vsum;  // [B][Hkv][VSUM_CHUNKS][D] partial column sums of V, then [B][Hkv][D] mean of V
  int* next_item;     // work-stealing counter (zeroed by v_mean_kernel before every launch)
  int B, H, Hkv, Lq, Skv;
  float scale_log2;
};

template <int D>
struct Smem {
  static constexpr int kQ = BM * D * 2;
  static constexpr int kK = BN * D * 2;
  static constexpr int kV = BN * D * 2;
  static constexpr int kStage = 4 * 32 * 80;              // per softmax warp: 32 rows x (64 B + 16 B pad) of output
  static constexpr int kTotal = kQ + kK + kV + kStage + 1024;  // + alignment slack
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): two lanes of a 64-bit register pair per instruction.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long pack2u(uint32_t lo, uint32_t hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
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
__device__ unsigned long long* g_timing = nullptr;  // [n_items][8] globaltimer stamps (bring-up only)
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define STAMP(item, slot)                                                                  \
  do {                                                                                     \
    if (g_timing && (threadIdx.x == 0 || threadIdx.x == 256)) g_timing[(item) * 8ull + (slot)] = gtime(); \
  } while (0)
#else
#define STAMP(item, slot) \
  do {                    \
  } while (0)
#endif

#ifdef SDAG_DEBUG_TIMING
// Sub-step trace (bring-up only): the first TRACE_CTAS CTAs record (clock64 << 8 | event) for softmax warp 0
// (role 0) and the MMA issuer (role 1); slot 0 of each role holds %smid.
constexpr int TRACE_CTAS = 8, TRACE_EV = 8192;
__device__ unsigned long long* g_trace = nullptr;
struct Tracer {
  unsigned long long* base;
  int n;
  __device__ __forceinline__ void init(int role, bool on) {
    base = (g_trace && on && blockIdx.x < TRACE_CTAS) ? g_trace + (blockIdx.x * 2 + role) * static_cast<long long>(TRACE_EV) : nullptr;
    n = 1;
    if (base) {
      uint32_t sm;
      asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
      base[0] = sm;
    }
  }
  __device__ __forceinline__ void ev(int e) {
    if (base && n < TRACE_EV) base[n++] = (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(e);
  }
};
#define TRACE_INIT(tr, role, on) tr.init(role, on)
#define TRACE(tr, e) tr.ev(e)
#else
struct Tracer {};
#define TRACE_INIT(tr, role, on) \
  do {                          \
  } while (0)
#define TRACE(tr, e) \
  do {               \
  } while (0)
#endif

#ifdef SDAG_DEBUG_TIMING
__device__ __forceinline__ bool g_timing_enabled() { return g_timing != nullptr; }
__device__ __forceinline__ void g_timing_write(long long item, int slot, long long v) { g_timing[item * 8ull + slot] = v; }
#else
__device__ __forceinline__ bool g_timing_enabled() { return false; }
__device__ __forceinline__ void g_timing_write(long long, int, long long) {}
#endif

struct Item {
  int q0, nrows, n_kv, h, b, hkv;
  int pack, R;  // GQA packing: `pack` query heads of one kv group share the tile, R = 128 / pack rows each
  const int4* kvt;
  const int32_t* map;
  bool valid;
};

// Work item i -> (query tile, head, prompt).  Tiles are the slow index so the longest tiles of every head run
// first and the G heads of a kv group run side by side (K/V shared through L2).
__device__ __forceinline__ Item fetch_item(const PrefillParams& p, long long i) {
  Item it;
  const unsigned per_tile = static_cast<unsigned>(p.H * p.B), iu = static_cast<unsigned>(i);  // n_items < 2^31
  const int t = static_cast<int>(iu / per_tile);
  const int rem = static_cast<int>(iu - static_cast<unsigned>(t) * per_tile);
  it.b = rem / p.H;
  it.h = rem - it.b * p.H;
  it.hkv = it.h / (p.H / p.Hkv);
  it.map = p.blockmaps + static_cast<long long>(it.b) * p.bm_stride_ints;
  it.valid = t < it.map[0] && it.map[2] == 0;
  it.q0 = it.nrows = it.n_kv = 0;
  it.kvt = nullptr;
  it.pack = 1, it.R = BM;
  if (it.valid) {
    const int4 qt = reinterpret_cast<const int4*>(it.map + SDAG_BLOCKMAP_HEADER)[t];
    it.q0 = qt.x, it.nrows = qt.y, it.n_kv = qt.w;
    it.kvt = reinterpret_cast<const int4*>(it.map + SDAG_BLOCKMAP_HEADER) + it.map[5] + qt.z;
    // Short query tiles (question block, system block): stack 4 or 2 query heads of the same kv group into the
    // 128-row MMA tile — K/V tiles are shared by the group, so one pass serves all of them.
    const int G = p.H / p.Hkv;
    if (it.n_kv > 0) {
      if (it.nrows <= 32 && (G % 4) == 0) it.pack = 4;
      else if (it.nrows <= 64 && (G % 2) == 0) it.pack = 2;
    }
    it.R = BM >> (it.pack >> 1);
    if (it.h % it.pack) it.valid = false;  // handled by the item of the group's first head
  }
  return it;
}

// What the scheduler (producer warp) publishes per item: the decoded item and its first two kv tile descriptors, so
// that the other roles start from shared memory instead of a chain of dependent global loads.
struct __align__(16) SchedSlot {
  int4 a;  // q0, nrows, n_kv, h
  int4 b;  // b, hkv, pack, state (1 = valid item, 0 = nothing to do, -1 = stop)
  int4 kv0, kv1;
  const int4* kvt;
  long long i;
};

__device__ __forceinline__ Item item_from_slot(const PrefillParams& p, const SchedSlot& sl) {
  const int4 a = sl.a, b = sl.b;
  Item it;
  it.q0 = a.x, it.nrows = a.y, it.n_kv = a.z, it.h = a.w;
  it.b = b.x, it.hkv = b.y, it.pack = b.z, it.valid = b.w == 1;
  it.R = BM >> (it.pack >> 1);
  it.kvt = sl.kvt;
  it.map = p.blockmaps + static_cast<long long>(it.b) * p.bm_stride_ints;
  return it;
}

// Every role derives the same item sequence from the scheduler's ring; mbarrier parities are tracked with per-role use
// counters.  The producers run ahead across item boundaries (Q after the item's last QK^T, K after each QK^T, V after
// each PV), and the first QK^T of the next item is issued while the softmax warps are still in the epilogue of the
// current one.  384 threads at 80 registers; setmaxnreg moves them to 136 (softmax) / 48 (producers, issuer) / 56 (epilogue).
template <int D>
__global__ void __launch_bounds__(384, 2)
sdag_prefill_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ64,
                    const __grid_constant__ CUtensorMap tmQ32, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const PrefillParams p, const long long n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Smem<D>::kQ;
  uint8_t* sV = sK + Smem<D>::kK;
  uint8_t* sStage = sV + Smem<D>::kV;
  __shared__ uint64_t q_full, q_empty, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], pv_done, o_full, o_empty, l_full[2], l_empty[2];
  __shared__ uint64_t sched_full[2], sched_empty[2];
  __shared__ SchedSlot sched[2];
  __shared__ float l_buf[2][BM];  // row sums handed from the softmax warpgroup to the epilogue warpgroup
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&q_empty, 1);
    for (int x = 0; x < 2; ++x) {
      mbar_init(&k_full[x], 1);
      mbar_init(&k_empty[x], 1);
      mbar_init(&v_full[x], 1);
      mbar_init(&v_empty[x], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&p_full[x], 128);
    }
    mbar_init(&pv_done, 1);
    mbar_init(&o_full, 1);
    mbar_init(&o_empty, 128);
    for (int x = 0; x < 2; ++x) {
      mbar_init(&l_full[x], 128);   // softmax threads: row sums of an item are in l_buf
      mbar_init(&l_empty[x], 128);  // epilogue threads: read them
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&sched_full[x], 1);
      mbar_init(&sched_empty[x], 32 + 128 + 1 + 128);  // MMA-warp lanes, softmax threads, the V producer, epilogue threads
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmQ32);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 5) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: two 64-column score buffers S0 / S1 (P, as bf16 pairs, is written over the first 32 columns of
  // the buffer it came from), then the fp32 output accumulator (D columns).
  const uint32_t tmem_S = tmem;
  const uint32_t tmem_O = tmem + 128;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ======================= epilogue warpgroup (warps 8..11) =======================
    // O / l -> bf16 -> per-warp staging rows -> coalesced 16-byte stores, off the softmax warps' critical path: while this
    // runs, they are already in the next item.  O is read out of TMEM first (o_empty), so the next item's first P.V —
    // issued only after that item's first softmax sub-step — normally never waits for it.
    const int quarter = warp - 8;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t nq = 0;
    for (uint32_t ns = 0;; ++ns) {
      const uint32_t slot = ns & 1;
      mbar_wait(&sched_full[slot], (ns >> 1) & 1);
      const Item it = item_from_slot(p, sched[slot]);
      const int state = sched[slot].b.w;
      const long long i = sched[slot].i;
      mbar_arrive(&sched_empty[slot]);
      if (state < 0) break;
      if (!it.valid || it.n_kv == 0) continue;
      const float* vm = p.vsum + (static_cast<long long>(p.B) * p.Hkv * VSUM_CHUNKS + static_cast<long long>(it.b) * p.Hkv + it.hkv) * D;
      const int q0 = it.q0;
      const int rr = row & (it.R - 1);
      const bool valid = rr < it.nrows;
      constexpr int SROW = 64 + 16;  // +16 B spreads the row-strided 16-byte stores over the banks
      uint8_t* wbase = sStage + quarter * 32 * SROW;
      uint8_t* srow = wbase + lane * SROW;
      const int wr = lane >> 2, ch = lane & 3;  // write-out role: row within a group of 8, 16-byte chunk of the row
      // destination of the four rows this lane writes out per 32-column round (nullptr: beyond the query tile)
      __nv_bfloat16* orow[4];
      {
        const int rshift = 7 - (it.pack >> 1);  // log2(R)
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int trow = quarter * 32 + g8 * 8 + wr;  // tile row
          const int orr = trow & (it.R - 1);
          orow[g8] = orr < it.nrows ? p.out + it.b * p.so_b + static_cast<long long>(q0 + orr) * p.so_l +
                                          (it.h + (trow >> rshift)) * p.so_h + ch * 8
                                    : nullptr;
        }
      }
      mbar_wait(&l_full[nq & 1], (nq >> 1) & 1);
      const float l = l_buf[nq & 1][row];
      mbar_arrive(&l_empty[nq & 1]);
      const float inv = l > 0.f ? 1.f / l : 0.f;
      mbar_wait(&o_full, nq & 1);
      tc_fence_after();
      STAMP(i, 3);
      uint32_t r[16];
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[16];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // two 16-column loads per 32-column staging round (register budget)
          tmem_ld_32x32b_x16(tmem_O + lane_addr + c * 32 + hh * 16, r);
          tmem_ld_wait();
          if (c == D / 32 - 1 && hh == 1) {
            // O is in registers now: the next item's first P.V may overwrite it
            tc_fence_before();
            mbar_arrive(&o_empty);
          }
          if (l > 0.f) {
#pragma unroll
            for (int e = 0; e < 16; e += 2)
              o[hh * 8 + (e >> 1)] = pack_bf16x2(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
          } else {
            // no allowed key in the whole row: uniform softmax over all Skv keys = mean of V
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 f = valid ? *reinterpret_cast<const float4*>(vm + c * 32 + hh * 16 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
              o[hh * 8 + (e >> 1)] = pack_bf16x2(f.x, f.y);
              o[hh * 8 + (e >> 1) + 1] = pack_bf16x2(f.z, f.w);
            }
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(srow);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
        dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
        __syncwarp();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          if (orow[g8]) {
            const uint4 val = *reinterpret_cast<const uint4*>(wbase + (g8 * 8 + wr) * SROW + ch * 16);
            *reinterpret_cast<uint4*>(orow[g8] + c * 32) = val;
          }
        }
        __syncwarp();
      }
      STAMP(i, 4);
      if (g_timing_enabled() && lane == 0 && quarter == 0) g_timing_write(i, 5, it.n_kv), g_timing_write(i, 6, blockIdx.x);
      ++nq;
    }
  } else if (warp >= 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
   if (warp == 4) {
    // ======================= TMA producer (Q, K) + scheduler =======================
    // The whole warp walks the loop (uniform control flow, uniform registers); single lanes are elected only to issue.
    {
      uint32_t nq = 0, nk = 0;  // loads issued so far per buffer
      // Dynamic longest-first scheduling: the first item is blockIdx.x, later ones come from a global counter; the
      // producer publishes each item index to the other roles through a two-slot ring (so it can run one item ahead).
      auto next_index = [&]() {
        int t = 0;
        if (lane == 0) t = atomicAdd(p.next_item, 1);
        return static_cast<long long>(gridDim.x) + __shfl_sync(0xffffffffu, t, 0);
      };
      long long i = blockIdx.x;
      long long i_next = next_index();
      for (uint32_t ns = 0;; ++ns) {
        const uint32_t slot = ns & 1;
        Item it;
        it.valid = false, it.n_kv = 0, it.q0 = it.nrows = it.h = it.b = it.hkv = 0, it.pack = 1, it.R = BM, it.kvt = nullptr;
        // items without tensor work are never handed out: tiles beyond a prompt's count, heads covered by a packed item of
        // their group, and key-less tiles (written by v_mean_kernel)
        while (i < n_items) {
          it = fetch_item(p, i);
          if (it.valid && it.n_kv > 0) break;
          i = i_next;
          i_next = next_index();
        }
        int4 kv0 = make_int4(0, 0, 0, 0), kv1 = kv0;
        if (it.valid && it.n_kv > 0) kv0 = it.kvt[0];
        if (it.valid && it.n_kv > 1) kv1 = it.kvt[1];
        mbar_wait(&sched_empty[slot], ((ns >> 1) & 1) ^ 1);
        if (lane == 0) {
          SchedSlot& sl = sched[slot];
          sl.a = make_int4(it.q0, it.nrows, it.n_kv, it.h);
          sl.b = make_int4(it.b, it.hkv, it.pack, i < n_items ? (it.valid ? 1 : 0) : -1);
          sl.kv0 = kv0, sl.kv1 = kv1;
          sl.kvt = it.kvt;
          sl.i = i;
          mbar_arrive(&sched_full[slot]);
        }
        __syncwarp();
        if (i >= n_items) break;
        // The item after this one is already known: pull its Q and its first K / V halves into L2 now, so that the
        // ring loads issued a few microseconds from now do not pay HBM latency (short items are latency-bound).
        if (i_next < n_items) {
          const Item nx = fetch_item(p, i_next);
          if (nx.valid && nx.n_kv > 0) {
            const CUtensorMap* tmq = nx.pack == 1 ? &tmQ : (nx.pack == 2 ? &tmQ64 : &tmQ32);
            if (elect_one()) {
              for (int g = 0; g < nx.pack; ++g)
                for (int dh = 0; dh < D / 64; ++dh) tma_prefetch_l2_4d(tmq, dh * 64, nx.q0, nx.h + g, nx.b);
            }
            int halves = 0;
            for (int j = 0; j < nx.n_kv && halves < 6; ++j) {
              const int4 kt = nx.kvt[j];
              for (int sub = 0; sub < (kt.y > 64 ? 2 : 1) && halves < 6; ++sub, ++halves) {
                if (elect_one()) {
                  for (int dh = 0; dh < D / 64; ++dh) {
                    tma_prefetch_l2_4d(&tmK, dh * 64, kt.x + sub * 64, nx.hkv, nx.b);
                    tma_prefetch_l2_4d(&tmV, dh * 64, kt.x + sub * 64, nx.hkv, nx.b);
                  }
                }
              }
            }
          }
        }
        i = i_next;
        i_next = next_index();
        if (!it.valid || it.n_kv == 0) continue;
        mbar_wait(&q_empty, (nq & 1) ^ 1);
        {
          const CUtensorMap* tm = it.pack == 1 ? &tmQ : (it.pack == 2 ? &tmQ64 : &tmQ32);
          if (elect_one()) {
            mbar_expect_tx(&q_full, Smem<D>::kQ);
            for (int g = 0; g < it.pack; ++g)
              for (int dh = 0; dh < D / 64; ++dh)
                tma_load_4d(sQ + dh * BM * 128 + g * it.R * 128, tm, &q_full, dh * 64, it.q0, it.h + g, it.b);
          }
        }
        int4 kt_next = kv0;
        for (int j = 0; j < it.n_kv; ++j) {
          const int4 kt = kt_next;
          if (j + 1 < it.n_kv) kt_next = j == 0 ? kv1 : it.kvt[j + 1];
          const int nsub = kt.y > 64 ? 2 : 1;
          for (int sub = 0; sub < nsub; ++sub) {
            // K travels as 64-row halves through a two-slot ring, so the load of sub-step u+1 is in flight while
            // sub-step u is being computed — also across tile and item boundaries.  (V has its own producer warp:
            // its slots free up later, after P.V, and must not hold back the K loads.)
            const uint32_t ks = nk & 1;
            mbar_wait(&k_empty[ks], ((nk >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&k_full[ks], Smem<D>::kK / 2);
              for (int dh = 0; dh < D / 64; ++dh)
                tma_load_4d(sK + ks * (Smem<D>::kK / 2) + dh * 64 * 128, &tmK, &k_full[ks], dh * 64, kt.x + sub * 64, it.hkv, it.b);
            }
            ++nk;
          }
        }
        ++nq;
      }
    }
  } else if (warp == 6) {
    // ======================= TMA producer for V =======================
    {
      uint32_t nv = 0;
      for (uint32_t ns = 0;; ++ns) {
        const uint32_t slot = ns & 1;
        mbar_wait(&sched_full[slot], (ns >> 1) & 1);
        const Item it = item_from_slot(p, sched[slot]);
        const int state = sched[slot].b.w;
        int4 kt_next = sched[slot].kv0;
        const int4 kv1 = sched[slot].kv1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sched_empty[slot]);
        if (state < 0) break;
        if (!it.valid || it.n_kv == 0) continue;
        for (int j = 0; j < it.n_kv; ++j) {
          const int4 kt = kt_next;
          if (j + 1 < it.n_kv) kt_next = j == 0 ? kv1 : it.kvt[j + 1];
          const int nsub = kt.y > 64 ? 2 : 1;
          for (int sub = 0; sub < nsub; ++sub) {
            const uint32_t vs = nv & 1;
            mbar_wait(&v_empty[vs], ((nv >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&v_full[vs], Smem<D>::kV / 2);
              for (int dh = 0; dh < D / 64; ++dh)
                tma_load_4d(sV + vs * (Smem<D>::kV / 2) + dh * 64 * 128, &tmV, &v_full[vs], dh * 64, kt.x + sub * 64, it.hkv, it.b);
            }
            ++nv;
          }
        }
      }
    }
  } else if (warp == 5) {
    // ======================= MMA issuer =======================
    // Each 128-column kv tile is processed as one or two 64-column sub-steps.  Scores of sub-step g go to S[g & 1],
    // so Q.K^T of sub-steps g+1 / g+2 is issued before the softmax of g has finished: issue order is
    //   QK(0) QK(1) | PV(0) QK(2) | PV(1) QK(3) | ...   (QK(g+2) overwrites the buffer whose P was consumed by PV(g);
    // the tensor pipe executes in issue order).  Issue is cheap (uniform descriptors, elect.sync), so PV(g) + QK(g+2)
    // fit underneath the softmax of g+1.
    uint32_t nq = 0;           // items with tensor work done so far
    uint32_t gq = 0, gp = 0;   // QK / PV sub-steps issued so far (S-buffer index and phases)
    Tracer tr;
    TRACE_INIT(tr, 1, lane == 0);
    for (uint32_t ns = 0;; ++ns) {
      const uint32_t slot = ns & 1;
      mbar_wait(&sched_full[slot], (ns >> 1) & 1);
      const Item it = item_from_slot(p, sched[slot]);
      const int state = sched[slot].b.w;
      const int4 kv0 = sched[slot].kv0, kv1 = sched[slot].kv1;
      mbar_arrive(&sched_empty[slot]);
      if (state < 0) break;
      if (!it.valid || it.n_kv == 0) continue;
      mbar_wait(&q_full, nq & 1);
      int qj = 0, qs = 0, pj = 0, ps = 0;  // cursors: (tile, sub-step) of the next QK / PV to issue
      bool first_pv = true;
      int4 kt_q = make_int4(0, 0, 0, 0), kt_p = kt_q;
      int4 kt_qnext = kv0, kt_pnext = kv0;
      const uint64_t descQ0 = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t descK0 = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t descV0 = umma_smem_desc_sw128(smem_u32(sV), 64 * 128, 1024);

      auto issue_qk = [&]() {
        if (qs == 0) {
          kt_q = kt_qnext;
          if (qj + 1 < it.n_kv) kt_qnext = qj == 0 ? kv1 : it.kvt[qj + 1];
        }
        const int ncols = kt_q.y;
        const int nsub = ncols > 64 ? 2 : 1;
        const int sub_cols = min(64, ncols - qs * 64);
        const uint32_t slot = gq & 1;  // K ring slot == S buffer index: both advance once per sub-step
        TRACE(tr, 10);
        mbar_wait_spin(&k_full[slot], (gq >> 1) & 1);
        TRACE(tr, 12);
        tc_fence_after();
        const bool last_sub = qs == nsub - 1;
        // descriptors are built in warp-uniform code (uniform registers); only the issue itself is single-lane
        const uint32_t idesc = umma_idesc_bf16_f32(BM, static_cast<uint32_t>((sub_cols + 15) & ~15), 0, 0);
        const uint64_t dk = descK0 + ((slot * (Smem<D>::kK / 2)) >> 4);
        const uint32_t tS = tmem_S + slot * 64;
        const bool last_qk = last_sub && qj == it.n_kv - 1;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = (ks >> 2) * (BM * 128) + (ks & 3) * 32;
            const uint32_t offk = (ks >> 2) * (64 * 128) + (ks & 3) * 32;
            umma_ss(tS, descQ0 + (off >> 4), dk + (offk >> 4), idesc, ks > 0);
          }
          umma_commit(&k_empty[slot]);         // K half consumed -> producer may refill the slot
          if (last_qk) umma_commit(&q_empty);  // Q consumed -> next item's Q may load
          umma_commit(&s_full[slot]);
        }
        __syncwarp();
        ++gq;
        if (last_sub) {
          ++qj;
          qs = 0;
        } else {
          ++qs;
        }
      };
      auto issue_pv = [&]() {
        if (ps == 0) {
          kt_p = kt_pnext;
          if (pj + 1 < it.n_kv) kt_pnext = pj == 0 ? kv1 : it.kvt[pj + 1];
        }
        const int ncols = kt_p.y;
        const int nsub = ncols > 64 ? 2 : 1;
        const int sub_cols = min(64, ncols - ps * 64);
        const uint32_t slot = gp & 1;
        TRACE(tr, 20);
        mbar_wait_spin(&p_full[slot], (gp >> 1) & 1);
        TRACE(tr, 21);
        mbar_wait_spin(&v_full[slot], (gp >> 1) & 1);
        TRACE(tr, 22);
        if (first_pv && nq > 0) mbar_wait(&o_empty, (nq - 1) & 1);  // previous item's epilogue has read O out of TMEM
        tc_fence_after();
        const bool last_sub = ps == nsub - 1;
        constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, D, 0, 1);  // B = V, MN-major
        const int pv_steps = (sub_cols + 15) >> 4;
        const uint64_t dv = descV0 + ((slot * (Smem<D>::kV / 2)) >> 4);
        const uint32_t tP = tmem_S + slot * 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < pv_steps) umma_ts(tmem_O, tP + ks * 8, dv + ((ks * 16 * 128) >> 4), idesc, (!first_pv || ks > 0) ? 1u : 0u);
          umma_commit(&v_empty[slot]);
          umma_commit(&pv_done);
          if (last_sub && pj == it.n_kv - 1) umma_commit(&o_full);
        }
        __syncwarp();
        first_pv = false;
        ++gp;
        if (last_sub) {
          ++pj;
          ps = 0;
        } else {
          ++ps;
        }
      };

      issue_qk();
      if (qj < it.n_kv) issue_qk();
      while (pj < it.n_kv) {
        issue_pv();
        if (qj < it.n_kv) issue_qk();
      }
      ++nq;
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    // ======================= softmax (warps 0..3) =======================
    const int quarter = warp;  // TMEM lanes this warp may touch: [32*quarter, 32*quarter+32)
    const int row = quarter * 32 + lane;
    const int stid = tid;      // 0..127 among the softmax threads
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t r[32], r2[32];
    uint32_t nq = 0, g = 0;  // items with tensor work / 64-column sub-steps processed so far
    Tracer tr;
    TRACE_INIT(tr, 0, tid == 0);

    for (uint32_t ns = 0;; ++ns) {
      const uint32_t slot = ns & 1;
      mbar_wait(&sched_full[slot], (ns >> 1) & 1);
      const Item it = item_from_slot(p, sched[slot]);
      const int state = sched[slot].b.w;
      const long long i = sched[slot].i;
      int4 kt_next = sched[slot].kv0;
      mbar_arrive(&sched_empty[slot]);
      if (state < 0) break;
      if (!it.valid) continue;
      STAMP(i, 0);
      const float* vm = p.vsum + (static_cast<long long>(p.B) * p.Hkv * VSUM_CHUNKS + static_cast<long long>(it.b) * p.Hkv + it.hkv) * D;
      if (it.n_kv == 0) {
        // Gap rows only (no allowed key): every row is the mean of V over all keys.  No tensor work.
        for (int e = stid; e < it.nrows * (D / 8); e += 128) {
          const int rr = e / (D / 8), ch = e % (D / 8);
          const float4 a = *reinterpret_cast<const float4*>(vm + ch * 8);
          const float4 c4 = *reinterpret_cast<const float4*>(vm + ch * 8 + 4);
          __nv_bfloat16* orow = p.out + it.b * p.so_b + static_cast<long long>(it.q0 + rr) * p.so_l + it.h * p.so_h;
          *reinterpret_cast<uint4*>(orow + ch * 8) =
              make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(c4.x, c4.y), pack_bf16x2(c4.z, c4.w));
        }
        continue;
      }
      const int q0 = it.q0;
      const int32_t* map = it.map;
      float m = -INFINITY, l = 0.f;
      const uint32_t g_item0 = g;  // first sub-step of this item
      // with GQA packing the tile row splits into (head within the group, row within the query tile)
      const int rr = row & (it.R - 1);
      const int wrow0 = (quarter * 32) & (it.R - 1);  // first query-tile row handled by this warp

      bool first_sub = true;
#pragma unroll 1
      for (int j = 0; j < it.n_kv; ++j) {
        const int4 kt = kt_next;
        if (j + 1 < it.n_kv) kt_next = it.kvt[j + 1];  // descriptor of the next tile: off the dependency chain
        const bool bitmask = (kt.w & KV_FLAG_BITMASK) != 0;
        unsigned long long bits_lo = ~0ull, bits_hi = ~0ull;  // allow bits of columns [0,64) / [64,128) of the tile
        if (bitmask) {
          uint4 w = make_uint4(0u, 0u, 0u, 0u);
          if (q0 + rr < p.Lq)
            w = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(map + map[10]) +
                                                static_cast<long long>(q0 + rr) * map[7] + (kt.x >> 5));
          bits_lo = w.x | (static_cast<unsigned long long>(w.y) << 32);
          bits_hi = w.z | (static_cast<unsigned long long>(w.w) << 32);
        }
        const int nsub = kt.y > 64 ? 2 : 1;
#pragma unroll 1
        for (int sub = 0; sub < nsub; ++sub, ++g) {
          const int ncols = min(64, kt.y - sub * 64);  // valid columns of this 64-column sub-step
          const int doff = kt.z - sub * 64;            // allowed: c <= r + doff (c relative to the sub-step)
          // this row: columns [0, lim] are allowed (lim < 0: none); bit-mask tiles carry no diagonal, only bits
          const int lim = bitmask ? ncols - 1 : min(ncols - 1, rr + doff);
          // warp-uniform bounds of lim over the 32 rows of this warp
          const int wlim_lo = min(ncols - 1, wrow0 + doff);
          const int wlim_hi = min(ncols - 1, wrow0 + 31 + doff);
          const bool full = !bitmask && wlim_lo >= 63;   // every row of the warp sees all 64 columns: no masking
          const bool none = !bitmask && wlim_hi < 0;     // no row of the warp sees any column
          const bool dead1 = !bitmask && wlim_hi < 32;   // columns [32,64) are masked for every row of the warp
          const uint32_t tS = tmem_S + lane_addr + (g & 1) * 64;
          const uint32_t tP = tS;
          TRACE(tr, 1);
          mbar_wait_spin(&s_full[g & 1], (g >> 1) & 1);
          TRACE(tr, 2);
          tc_fence_after();
          if (first_sub) STAMP(i, 1);
          first_sub = false;
          if (none) {
            // P = 0 for these rows (the MMA still spans all 128 rows of the tile)
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = 0u;
            tmem_st_32x32b_x32(tP, r);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[g & 1]);
            continue;
          }

          // Single pass over the 64 scores of this sub-step: maximum over the allowed columns, keep the running
          // reference maximum m unless it is exceeded by more than RESCALE_TAU (then l and O, accumulated against
          // the old reference, are rescaled — rare), and write P = exp2(S*scale - m) as bf16 pairs over the first
          // 32 columns of the same buffer.  Columns that are not allowed are replaced by -inf up front, so the
          // arithmetic below exists once.
          tmem_ld_32x32b_x32(tS, r);
          if (!dead1) tmem_ld_32x32b_x32(tS + 32, r2);
          tmem_ld_wait();
          TRACE(tr, 3);
          if (!full) {
            unsigned long long allow = lim < 0 ? 0ull : (lim >= 63 ? ~0ull : ((2ull << lim) - 1ull));
            allow &= sub ? bits_hi : bits_lo;
            const uint32_t alo = static_cast<uint32_t>(allow), ahi = static_cast<uint32_t>(allow >> 32);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              r[e] = (alo >> e) & 1u ? r[e] : 0xff800000u;
              r2[e] = (ahi >> e) & 1u ? r2[e] : 0xff800000u;
            }
          }
          float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            mxa = fmax3(mxa, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
            mxb = fmax3(mxb, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          }
          if (!dead1) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              mxa = fmax3(mxa, __uint_as_float(r2[e]), __uint_as_float(r2[e + 1]));
              mxb = fmax3(mxb, __uint_as_float(r2[e + 2]), __uint_as_float(r2[e + 3]));
            }
          }
          const float mx = fmaxf(mxa, mxb) * p.scale_log2;
          const bool need = mx > m + RESCALE_TAU;  // also the first finite maximum (m == -inf)
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? exp2f(m - mx) : 1.f;
            if (need) m = mx;
            l *= alpha;
            if (g != g_item0) {
              // O holds the P.V of every earlier sub-step of this item; the latest one may still be in flight
              mbar_wait(&pv_done, (g - 1) & 1);
              tc_fence_after();
              uint32_t t[32];
#pragma unroll 1
              for (int oc = 0; oc < D / 32; ++oc) {
                tmem_ld_32x32b_x32(tmem_O + lane_addr + oc * 32, t);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) t[e] = __float_as_uint(__uint_as_float(t[e]) * alpha);
                tmem_st_32x32b_x32(tmem_O + lane_addr + oc * 32, t);
              }
            }
          }
          const float neg_m = (m == -INFINITY) ? 0.f : -m;
          const unsigned long long scale2 = pack2(p.scale_log2, p.scale_log2), negm2 = pack2(neg_m, neg_m);
          unsigned long long lsA = pack2(0.f, 0.f), lsB = lsA;
          uint32_t pk[32];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float x0, x1, x2, x3;
            unpack2(fma2(pack2u(r[e], r[e + 1]), scale2, negm2), x0, x1);
            unpack2(fma2(pack2u(r[e + 2], r[e + 3]), scale2, negm2), x2, x3);
            const float p0 = exp2f(x0), p1 = exp2f(x1), p2 = exp2f(x2), p3 = exp2f(x3);
            lsA = add2(lsA, pack2(p0, p1));
            lsB = add2(lsB, pack2(p2, p3));
            pk[e >> 1] = pack_bf16x2(p0, p1);
            pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          if (!dead1) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              float x0, x1, x2, x3;
              unpack2(fma2(pack2u(r2[e], r2[e + 1]), scale2, negm2), x0, x1);
              unpack2(fma2(pack2u(r2[e + 2], r2[e + 3]), scale2, negm2), x2, x3);
              const float p0 = exp2f(x0), p1 = exp2f(x1), p2 = exp2f(x2), p3 = exp2f(x3);
              lsA = add2(lsA, pack2(p0, p1));
              lsB = add2(lsB, pack2(p2, p3));
              pk[16 + (e >> 1)] = pack_bf16x2(p0, p1);
              pk[16 + (e >> 1) + 1] = pack_bf16x2(p2, p3);
            }
          } else {
#pragma unroll
            for (int e = 16; e < 32; ++e) pk[e] = 0u;
          }
          tmem_st_32x32b_x32(tP, pk);
          {
            float a0, a1;
            unpack2(add2(lsA, lsB), a0, a1);
            l += a0 + a1;
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[g & 1]);
          TRACE(tr, 4);
        }
      }

      // ---- hand the row sums to the epilogue warpgroup and move on to the next item ----
      STAMP(i, 2);
      TRACE(tr, 8);
      if (nq >= 2) mbar_wait(&l_empty[nq & 1], ((nq >> 1) - 1) & 1);
      l_buf[nq & 1][row] = l;
      mbar_arrive(&l_full[nq & 1]);
      ++nq;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 256);
}

// Partial column sums of V over VSUM_CHUNKS row ranges: vsum[b][hkv][chunk][d] (fp32, deterministic).
template <int D>
__global__ void __launch_bounds__(256) v_colsum_kernel(const __nv_bfloat16* __restrict__ v, long long sv_b, long long sv_h,
                                                       long long sv_l, int Skv, float* __restrict__ vsum,
                                                       const int32_t* __restrict__ blockmaps, long long bm_stride_ints) {
  const int chunk = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  if (blockmaps && blockmaps[static_cast<long long>(b) * bm_stride_ints + 3] == 0) return;  // no gap rows
  // ragged batches: a prompt's uniform rows average over its own L keys (header word 4), not the padded length
  if (blockmaps) Skv = min(Skv, blockmaps[static_cast<long long>(b) * bm_stride_ints + 4]);
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

// vmean[b][hkv][d] = (sum of the chunk partials, fixed order) / Skv, stored after the partials.  The same block also
// writes the output of every query tile that has no kv tile at all (gap rows: uniform softmax = mean of V) for the G
// query heads of its kv group, so the attention kernel's scheduler never has to hand those tiles out.
template <int D>
__global__ void __launch_bounds__(D) v_mean_kernel(float* __restrict__ vsum, int B, int Hkv, int Skv,
                                                   const int32_t* __restrict__ blockmaps, long long bm_stride_ints,
                                                   __nv_bfloat16* __restrict__ out, long long so_b, long long so_l, long long so_h,
                                                   int H, int Lq) {
  const int hkv = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  if (hkv == 0 && b == 0 && d == 0)  // work-stealing counter of the attention kernel that follows on the stream
    *reinterpret_cast<int*>(vsum + static_cast<long long>(B) * Hkv * (VSUM_CHUNKS + 1) * D) = 0;
  if (blockmaps && blockmaps[static_cast<long long>(b) * bm_stride_ints + 3] == 0) return;
  if (blockmaps) Skv = min(Skv, blockmaps[static_cast<long long>(b) * bm_stride_ints + 4]);
  const float* part = vsum + (static_cast<long long>(b) * Hkv + hkv) * VSUM_CHUNKS * D;
  float s = 0.f;
  for (int ch = 0; ch < VSUM_CHUNKS; ++ch) s += part[ch * D + d];
  const float mean = s / static_cast<float>(Skv);
  vsum[(static_cast<long long>(B) * Hkv * VSUM_CHUNKS + static_cast<long long>(b) * Hkv + hkv) * D + d] = mean;
  if (!blockmaps) return;
  const int32_t* map = blockmaps + static_cast<long long>(b) * bm_stride_ints;
  if (map[2] != 0) return;
  const __nv_bfloat16 mb = __float2bfloat16(mean);
  const int4* qtiles = reinterpret_cast<const int4*>(map + SDAG_BLOCKMAP_HEADER);
  const int G = H / Hkv;
  // the threads scan the tile list in parallel and collect the key-less tiles; then every thread writes its column
  __shared__ int2 gaps[128];
  __shared__ int n_gaps;
  if (d == 0) n_gaps = 0;
  __syncthreads();
  for (int t0 = 0; t0 < map[0]; t0 += 128 * D / D) {  // rounds of up to 128 collected tiles keep the list bounded
    for (int t = t0 + d; t < min(map[0], t0 + 128); t += D) {
      const int4 qt = qtiles[t];
      if (qt.w == 0) gaps[atomicAdd(&n_gaps, 1)] = make_int2(qt.x, min(qt.x + qt.y, Lq));
    }
    __syncthreads();
    for (int i = 0; i < n_gaps; ++i)
      for (int r = gaps[i].x; r < gaps[i].y; ++r)
        for (int gh = 0; gh < G; ++gh) out[b * so_b + r * so_l + (hkv * G + gh) * so_h + d] = mb;
    __syncthreads();
    if (d == 0) n_gaps = 0;
    __syncthreads();
  }
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
int launch_prefill(const CUtensorMap& tq, const CUtensorMap& tq64, const CUtensorMap& tq32, const CUtensorMap& tk,
                   const CUtensorMap& tv, const PrefillParams& p, int cap_qtiles, cudaStream_t st) {
  static bool configured[SDAG_MAX_DEVICES] = {};
  const int smem = Smem<D>::kTotal > 100 * 1024 ? Smem<D>::kTotal : 100 * 1024;  // cap residency at 2 CTAs / SM (TMEM)
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (!configured[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(sdag_prefill_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = true;
  }
  const int n_sm = device_sm_count();
  const long long n_items = static_cast<long long>(cap_qtiles) * p.H * p.B;
  if (n_items <= 0) return SDAG_ERR_INVALID;
  const long long grid = n_items < 2ll * n_sm ? n_items : 2ll * n_sm;
  sdag_prefill_kernel<D><<<static_cast<unsigned>(grid), 384, smem, st>>>(tq, tq64, tq32, tk, tv, p, n_items);
  return static_cast<int>(cudaGetLastError());
}


// ================================================================================================================
// CTA-pair variant (head_dim 128, an even number of query heads per kv head).
//
// The single-CTA kernel above is bound by what a cta_group::1 SS-mode MMA costs: a 128 x 64 x 16 Q.K^T instruction takes 80
// cycles against 32 cycles of math because the 128 x 16 slice of Q is re-read from shared memory every time
// (probe/mma_rate_probe.cu), which holds dense tiles at ~0.55 of the tensor pipe.  Two query heads of one GQA group share
// their K / V tiles, so a pair of CTAs on neighbouring SMs can run them as ONE tcgen05.mma.cta_group::2 of M = 256: CTA r
// of the pair owns head h0 + r (its 128 query rows, its S / P / O rows in its own TMEM, its softmax and epilogue), loads
// only HALF of every K tile (64 of the 128 kv rows: its half of the N dimension of S = Q.K^T) and HALF of every V tile (64
// of the 128 d columns: its half of the N dimension of O += P.V), and the leader CTA issues every MMA for both.  Per SM
// and per 128 x 128 tile the operand traffic from shared memory drops from 2 x (Q slices + K) + V to Q slices + K/2 + V/2
// and the MMAs run at N = 128, the shape that reaches the full rate on pairs (probe/mma2_rate_probe.cu).
//
// One CTA per SM, 512 TMEM columns: S0 / S1 (two 128-column score tiles; P, as bf16 pairs, is written over the first 64
// columns of the tile it came from), O0 / O1 (the output accumulator is double-buffered across work items, so an item's
// first P.V never waits for the previous item's epilogue).  Shared memory: Q double-buffered (2 x 32 KiB), four stages of
// K halves and V halves (16 KiB each), the epilogue's staging rows.  Work items (query tile, head pair, prompt) are dealt
// round-robin to the clusters, longest first — both CTAs of a pair and every role derive the same sequence, nothing is
// exchanged for scheduling.  Cross-CTA protocol (same as the retrieval / linear-layer pair kernels): every TMA load
// completes on the LEADER's `full` barrier (the leader announces the bytes of both CTAs), tcgen05.commit multicasts to the
// barriers of both CTAs, and the peer's softmax / epilogue warps signal the leader's `p_full` / `o_empty` barriers with
// one remote arrive per warp.
//   warps 0-15  softmax, four threads per query row (one 32-column chunk of the 128-column score tile each): maximum over
//               the allowed columns, exchanged between the four threads, lazy rescaling, P = exp2(S * scale - m) as bf16
//               (half of the exponentials on the exp unit, half as a cubic on the FMA pipe)
//   warp 16     TMA producer (both CTAs): Q, then K half / V half per tile
//   warp 17     MMA issuer (leader CTA only)
//   warps 20-23 epilogue: O / l -> bf16 -> staged, coalesced stores
// ================================================================================================================
namespace pairk {

// 2^x for two values on the FMA pipe (no MUFU): x = n + f with n = rint(x), f in [-0.5, 0.5]; 2^f by the cubic
// 1 + f (ln2 + f (ln2^2/2 + f ln2^3/6)) (relative error <= 4e-4, a fifth of a bf16 half-ulp — P is stored as bf16), 2^n by
// adding n to the exponent field.  Inputs are clamped to >= -125 so the exponent field never underflows; masked columns
// (x = -inf) therefore come out as 2^-125 ~ 2e-38 instead of exactly 0, which is below anything a row sum can see.  The
// exp unit does 16 results per clock per SM: for a 128 x 128 score tile per SM that is 1024 cycles — as long as the tile's
// tensor-core work on a CTA pair — so half of every row's exponentials are computed here, on otherwise idle FMA slots.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& o0, float& o1) {
  const unsigned long long x2 = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const unsigned long long t2 = add2(x2, pack2(12582912.f, 12582912.f));  // 1.5 * 2^23: the low mantissa bits now hold rint(x)
  const unsigned long long n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const unsigned long long f2 = fma2(n2, pack2(-1.f, -1.f), x2);
  unsigned long long q2 = fma2(f2, pack2(0.0555041087f, 0.0555041087f), pack2(0.2402265070f, 0.2402265070f));
  q2 = fma2(q2, f2, pack2(0.6931471806f, 0.6931471806f));
  q2 = fma2(q2, f2, pack2(1.f, 1.f));
  float t0, t1, q0, q1;
  unpack2(t2, t0, t1);
  unpack2(q2, q0, q1);
  o0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  o1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

constexpr int KV_STAGES = 4;
constexpr int Q_BYTES = BM * 128 * 2;  // 128 query rows x 128 d
constexpr int K_BYTES = 64 * 128 * 2;  // this CTA's 64 of the tile's 128 kv rows x 128 d
constexpr int V_BYTES = 128 * 64 * 2;  // 128 kv rows x this CTA's 64 of the 128 d columns
constexpr int OUT_STAGE = 4 * 32 * 80;
constexpr int SMEM_TOTAL = 2 * Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + OUT_STAGE + 1024;
constexpr int D = 128;
constexpr int SPLIT = 4;          // softmax threads per query row: each takes one 32-column chunk of the 128-column score tile
constexpr int THREADS = 768;      // warps 0-15 softmax (warp = part * 4 + row quarter), 16 producer, 17 issuer, 20-23 epilogue

struct PItem {
  int q0, nrows, n_kv, h0, b, hkv;
  int pack, R;  // query heads per CTA (1, or 2 for tiles of <= 64 rows), rows per head in the 128-row tile
  const int4* kvt;
  const int32_t* map;
  bool valid;
};

// Work item i -> (query tile, head-pair slot, prompt); tiles are the slow index (longest kv lists first).  A slot covers
// query heads h0 .. h0 + 2 * pack - 1 of one kv group: CTA r takes heads h0 + r * pack ... (+ pack - 1).
__device__ __forceinline__ PItem fetch_pitem(const PrefillParams& p, long long i) {
  PItem it;
  const unsigned slots = static_cast<unsigned>(p.H >> 1);
  const unsigned per_tile = slots * static_cast<unsigned>(p.B), iu = static_cast<unsigned>(i);
  const int t = static_cast<int>(iu / per_tile);
  const int rem = static_cast<int>(iu - static_cast<unsigned>(t) * per_tile);
  it.b = rem / static_cast<int>(slots);
  const int slot = rem - it.b * static_cast<int>(slots);
  it.h0 = 2 * slot;
  const int G = p.H / p.Hkv;
  it.hkv = it.h0 / G;
  it.map = p.blockmaps + static_cast<long long>(it.b) * p.bm_stride_ints;
  it.valid = t < it.map[0] && it.map[2] == 0;
  it.q0 = it.nrows = it.n_kv = 0;
  it.kvt = nullptr;
  it.pack = 1, it.R = BM;
  if (it.valid) {
    const int4 qt = reinterpret_cast<const int4*>(it.map + SDAG_BLOCKMAP_HEADER)[t];
    it.q0 = qt.x, it.nrows = qt.y, it.n_kv = qt.w;
    it.kvt = reinterpret_cast<const int4*>(it.map + SDAG_BLOCKMAP_HEADER) + it.map[5] + qt.z;
    if (it.n_kv > 0 && it.nrows <= 64 && (G % 4) == 0) {
      it.pack = 2, it.R = 64;
      if (slot & 1) it.valid = false;  // heads h0 .. h0 + 3 are covered by the item of the even slot
    }
  }
  return it;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
sdag_prefill_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ64,
                         const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const PrefillParams p,
                         const long long n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // [2][Q_BYTES]
  uint8_t* sK = sQ + 2 * Q_BYTES;                    // [KV_STAGES][K_BYTES]
  uint8_t* sV = sK + KV_STAGES * K_BYTES;            // [KV_STAGES][V_BYTES]
  uint8_t* sStage = sV + KV_STAGES * V_BYTES;
  __shared__ uint64_t q_full[2], q_empty[2], k_full[KV_STAGES], k_empty[KV_STAGES], v_full[KV_STAGES], v_empty[KV_STAGES];
  __shared__ uint64_t s_full[2], p_full[2], pv_done, o_full[2], o_empty[2], l_full[2], l_empty[2];
  __shared__ float l_buf[2][SPLIT][BM];   // [O buffer][column part][row]: partial row sums handed to the epilogue
  __shared__ float mx_buf[2][SPLIT][BM];  // [tile parity][column part][row]: the threads of a row exchange their maxima
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const long long cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(&q_full[x], 1);    // (leader's copy is live) the leader's producer announces the bytes of both CTAs
      mbar_init(&q_empty[x], 1);   // multicast commit
      mbar_init(&s_full[x], 1);    // multicast commit
      mbar_init(&p_full[x], 8 * SPLIT);  // (leader's copy) one arrival per softmax warp of both CTAs
      mbar_init(&o_full[x], 1);    // multicast commit
      mbar_init(&o_empty[x], 8);   // (leader's copy) one arrival per epilogue warp of both CTAs
      mbar_init(&l_full[x], 128 * SPLIT);  // every softmax thread of every row
      mbar_init(&l_empty[x], 128);
    }
    for (int x = 0; x < KV_STAGES; ++x) {
      mbar_init(&k_full[x], 1);
      mbar_init(&k_empty[x], 1);
      mbar_init(&v_full[x], 1);
      mbar_init(&v_empty[x], 1);
    }
    mbar_init(&pv_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17) {
    tmem_alloc_2cta(&tmem_base_s, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit / peer TMA completion
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_S = tmem;        // + buf * 128
  const uint32_t tmem_O = tmem + 256;  // + ob * 128

  // Registers: the CTA's pool is what it was launched with, 768 x 80; setmaxnreg re-deals it 16 x 32 x 88 (softmax) +
  // 4 x 32 x 56 (producer, issuer, two idle warps) + 4 x 32 x 72 (epilogue) = 61 440 exactly — an `inc` beyond the pool blocks forever.
  if (warp >= 20) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    // ======================= epilogue warpgroup =======================
    const int quarter = warp - 20;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t lead_oempty0 = mapa_shared(smem_u32(&o_empty[0]), 0);
    uint32_t nq = 0;
    for (long long i = cid; i < n_items; i += n_clusters) {
      const PItem it = fetch_pitem(p, i);
      if (!it.valid || it.n_kv == 0) continue;
      const uint32_t ob = nq & 1;
      const float* vm = p.vsum + (static_cast<long long>(p.B) * p.Hkv * VSUM_CHUNKS + static_cast<long long>(it.b) * p.Hkv + it.hkv) * D;
      const int hb = it.h0 + static_cast<int>(rank) * it.pack;
      const int rr = row & (it.R - 1);
      const bool valid = rr < it.nrows;
      constexpr int SROW = 64 + 16;
      uint8_t* wbase = sStage + quarter * 32 * SROW;
      uint8_t* srow = wbase + lane * SROW;
      const int wr = lane >> 2, ch = lane & 3;
      __nv_bfloat16* orow[4];
      {
        const int rshift = 7 - (it.pack >> 1);  // log2(R)
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int trow = quarter * 32 + g8 * 8 + wr;
          const int orr = trow & (it.R - 1);
          orow[g8] = orr < it.nrows ? p.out + it.b * p.so_b + static_cast<long long>(it.q0 + orr) * p.so_l +
                                          (hb + (trow >> rshift)) * p.so_h + ch * 8
                                    : nullptr;
        }
      }
      mbar_wait(&l_full[ob], (nq >> 1) & 1);
      float l = 0.f;
#pragma unroll
      for (int sp = 0; sp < SPLIT; ++sp) l += l_buf[ob][sp][row];
      mbar_arrive(&l_empty[ob]);
      const float inv = l > 0.f ? 1.f / l : 0.f;
      mbar_wait(&o_full[ob], (nq >> 1) & 1);
      tc_fence_after();
      uint32_t r[16];
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[16];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld_32x32b_x16(tmem_O + ob * 128 + lane_addr + c * 32 + hh * 16, r);
          tmem_ld_wait();
          if (c == D / 32 - 1 && hh == 1) {
            // every lane of this warp holds its last O values: O[ob] may be overwritten by the item after the next
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_oempty0 + ob * 8);
          }
          if (l > 0.f) {
#pragma unroll
            for (int e = 0; e < 16; e += 2)
              o[hh * 8 + (e >> 1)] = pack_bf16x2(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
          } else {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 f = valid ? *reinterpret_cast<const float4*>(vm + c * 32 + hh * 16 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
              o[hh * 8 + (e >> 1)] = pack_bf16x2(f.x, f.y);
              o[hh * 8 + (e >> 1) + 1] = pack_bf16x2(f.z, f.w);
            }
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(srow);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
        dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
        __syncwarp();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          if (orow[g8]) {
            const uint4 val = *reinterpret_cast<const uint4*>(wbase + (g8 * 8 + wr) * SROW + ch * 16);
            *reinterpret_cast<uint4*>(orow[g8] + c * 32) = val;
          }
        }
        __syncwarp();
      }
      ++nq;
    }
  } else if (warp >= 16) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // warps 16-19 (two of them idle) hand registers to the softmax warps
   if (warp == 16) {
    // ======================= TMA producer (both CTAs) =======================
    const uint32_t lead_qfull0 = mapa_shared(smem_u32(&q_full[0]), 0);
    const uint32_t lead_kfull0 = mapa_shared(smem_u32(&k_full[0]), 0);
    const uint32_t lead_vfull0 = mapa_shared(smem_u32(&v_full[0]), 0);
    uint32_t nq = 0, nkv = 0;
    for (long long i = cid; i < n_items; i += n_clusters) {
      const PItem it = fetch_pitem(p, i);
      if (!it.valid || it.n_kv == 0) continue;
      const uint32_t qb = nq & 1;
      const int hb = it.h0 + static_cast<int>(rank) * it.pack;
      mbar_wait(&q_empty[qb], ((nq >> 1) & 1) ^ 1);
      {
        const CUtensorMap* tm = it.pack == 1 ? &tmQ : &tmQ64;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&q_full[qb], 2 * Q_BYTES);
          for (int g = 0; g < it.pack; ++g)
            for (int dh = 0; dh < 2; ++dh)
              tma_load_4d_2sm(sQ + qb * Q_BYTES + dh * BM * 128 + g * it.R * 128, tm, lead_qfull0 + qb * 8, dh * 64, it.q0, hb + g, it.b);
        }
        __syncwarp();
      }
      int4 kt_next = it.kvt[0];
      for (int j = 0; j < it.n_kv; ++j, ++nkv) {
        const int4 kt = kt_next;
        if (j + 1 < it.n_kv) kt_next = it.kvt[j + 1];
        const uint32_t s = nkv % KV_STAGES;
        const uint32_t par = ((nkv / KV_STAGES) & 1) ^ 1;
        mbar_wait(&k_empty[s], par);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&k_full[s], 2 * K_BYTES);
          for (int dh = 0; dh < 2; ++dh)
            tma_load_4d_2sm(sK + s * K_BYTES + dh * 64 * 128, &tmK, lead_kfull0 + s * 8, dh * 64, kt.x + static_cast<int>(rank) * 64, it.hkv, it.b);
        }
        __syncwarp();
        mbar_wait(&v_empty[s], par);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&v_full[s], 2 * V_BYTES);
          tma_load_4d_2sm(sV + s * V_BYTES, &tmV, lead_vfull0 + s * 8, static_cast<int>(rank) * 64, kt.x, it.hkv, it.b);
        }
        __syncwarp();
      }
      ++nq;
    }
   } else if (warp == 17) {
    // ======================= MMA issuer (leader CTA only) =======================
    if (leader) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16_f32(256, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(256, 128, 0, 1);  // B = V, MN-major
      uint32_t nq = 0, gq = 0, gp = 0;
      Tracer tr;
      TRACE_INIT(tr, 1, lane == 0);
      for (long long i = cid; i < n_items; i += n_clusters) {
        const PItem it = fetch_pitem(p, i);
        if (!it.valid || it.n_kv == 0) continue;
        const uint32_t qb = nq & 1, ob = nq & 1;
        TRACE(tr, 30);
        mbar_wait(&q_full[qb], (nq >> 1) & 1);
        TRACE(tr, 31);
        const uint64_t descQ0 = umma_smem_desc_sw128(smem_u32(sQ + qb * Q_BYTES), 16, 1024);
        int qj = 0, pj = 0;
        bool first_pv = true;
        int4 kt_p = it.kvt[0];

        auto issue_qk = [&]() {
          const uint32_t s = gq % KV_STAGES, buf = gq & 1;
          TRACE(tr, 10);
          mbar_wait_spin(&k_full[s], (gq / KV_STAGES) & 1);
          TRACE(tr, 12);
          tc_fence_after();
          const uint64_t dk = umma_smem_desc_sw128(smem_u32(sK + s * K_BYTES), 16, 1024);
          const uint32_t tS = tmem_S + buf * 128;
          const bool last_qk = qj == it.n_kv - 1;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t off = (ks >> 2) * (BM * 128) + (ks & 3) * 32;
              const uint32_t offk = (ks >> 2) * (64 * 128) + (ks & 3) * 32;
              umma_ss_2cta(tS, descQ0 + (off >> 4), dk + (offk >> 4), idesc_qk, ks > 0);
            }
            umma_commit_2cta(&k_empty[s], 3);
            if (last_qk) umma_commit_2cta(&q_empty[qb], 3);
            umma_commit_2cta(&s_full[buf], 3);
          }
          __syncwarp();
          ++gq;
          ++qj;
        };
        auto issue_pv = [&]() {
          const int4 kt = kt_p;
          if (pj + 1 < it.n_kv) kt_p = it.kvt[pj + 1];
          const uint32_t s = gp % KV_STAGES, buf = gp & 1;
          TRACE(tr, 20);
          mbar_wait_spin(&p_full[buf], (gp >> 1) & 1);
          TRACE(tr, 21);
          mbar_wait_spin(&v_full[s], (gp / KV_STAGES) & 1);
          if (first_pv && nq >= 2) mbar_wait(&o_empty[ob], ((nq >> 1) - 1) & 1);  // the epilogues of item nq - 2 have read O[ob]
          TRACE(tr, 22);
          tc_fence_after();
          const int pv_steps = (min(kt.y, BN) + 15) >> 4;
          const uint64_t dv = umma_smem_desc_sw128(smem_u32(sV + s * V_BYTES), 64 * 128, 1024);
          const uint32_t tP = tmem_S + buf * 128;
          const bool last_pv = pj == it.n_kv - 1;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BN / 16; ++ks)
              if (ks < pv_steps)
                umma_ts_2cta(tmem_O + ob * 128, tP + ks * 8, dv + ((ks * 16 * 128) >> 4), idesc_pv, (!first_pv || ks > 0) ? 1u : 0u);
            umma_commit_2cta(&v_empty[s], 3);
            umma_commit_2cta(&pv_done, 3);
            if (last_pv) umma_commit_2cta(&o_full[ob], 3);
          }
          __syncwarp();
          first_pv = false;
          ++gp;
          ++pj;
        };

        issue_qk();
        if (qj < it.n_kv) issue_qk();
        while (pj < it.n_kv) {
          issue_pv();
          if (qj < it.n_kv) issue_qk();
        }
        ++nq;
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ======================= softmax: SPLIT threads per query row, one 32-column chunk of every score tile each ==========
    // (the instruction stream of one row is a chain of dependent ALU / MUFU / TMEM operations: a lone warp per scheduler ran
    // its 762 instructions per tile at an IPC of 0.3, two warps at 0.5; four warps per scheduler hide each other's latencies)
    const int quarter = warp & 3;
    const int part = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t lead_pfull0 = mapa_shared(smem_u32(&p_full[0]), 0);
    uint32_t nq = 0, g = 0;
    Tracer tr;
    TRACE_INIT(tr, 0, tid == 0);
    for (long long i = cid; i < n_items; i += n_clusters) {
      const PItem it = fetch_pitem(p, i);
      if (!it.valid || it.n_kv == 0) continue;
      const uint32_t ob = nq & 1;
      const int q0 = it.q0;
      const int32_t* map = it.map;
      float m = -INFINITY, l = 0.f;
      const uint32_t g_item0 = g;
      const int rr = row & (it.R - 1);
      const int wrow0 = (quarter * 32) & (it.R - 1);
      int4 kt_next = it.kvt[0];
#pragma unroll 1
      for (int j = 0; j < it.n_kv; ++j, ++g) {
        const int4 kt = kt_next;
        if (j + 1 < it.n_kv) kt_next = it.kvt[j + 1];
        const bool bitmask = (kt.w & KV_FLAG_BITMASK) != 0;
        uint4 bits = make_uint4(~0u, ~0u, ~0u, ~0u);  // allow bits of the tile's four 32-column chunks
        if (bitmask) {
          bits = make_uint4(0u, 0u, 0u, 0u);
          if (q0 + rr < p.Lq)
            bits = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(map + map[10]) +
                                                   static_cast<long long>(q0 + rr) * map[7] + (kt.x >> 5));
        }
        const int ncols = min(BN, kt.y);
        const int doff = kt.z;
        const int lim = bitmask ? ncols - 1 : min(ncols - 1, rr + doff);           // this row: columns [0, lim] (< 0: none)
        const int wlim_lo = min(ncols - 1, wrow0 + doff);                            // warp-uniform bounds of lim
        const int wlim_hi = bitmask ? ncols - 1 : min(ncols - 1, wrow0 + 31 + doff);
        const bool full = !bitmask && wlim_lo >= BN - 1;
        const uint32_t buf = g & 1;
        const uint32_t tS = tmem_S + buf * 128 + lane_addr;
        TRACE(tr, 1);
        mbar_wait_spin(&s_full[buf], (g >> 1) & 1);
        TRACE(tr, 2);
        tc_fence_after();
        auto chunk_mask = [&](int c) -> uint32_t {  // allowed columns of chunk c for this row
          const int rel = lim - 32 * c;
          const uint32_t up_to = rel < 0 ? 0u : (rel >= 31 ? ~0u : ((2u << rel) - 1u));
          return up_to & (c == 0 ? bits.x : (c == 1 ? bits.y : (c == 2 ? bits.z : bits.w)));
        };
        const int n_chunks = wlim_hi < 0 ? 0 : (wlim_hi >> 5) + 1;  // chunks some row of this warp can see (warp-uniform)
        const int c = part;  // this thread's chunk: columns [32 c, 32 c + 32)
        const bool live = c < n_chunks;  // warp-uniform: some row of this warp sees a column of the chunk

        uint32_t r[32];
        if (live) tmem_ld_32x32b_x32(tS + c * 32, r);
        tmem_ld_wait();
        TRACE(tr, 3);
        float mxa = -INFINITY, mxb = -INFINITY;
        if (live) {
          if (!full) {
            const uint32_t am = chunk_mask(c);
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = (am >> e) & 1u ? r[e] : 0xff800000u;
          }
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            mxa = fmax3(mxa, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
            mxb = fmax3(mxb, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          }
        }
        // row maximum = max over the SPLIT chunks: exchange through shared memory (buffer by tile parity, one barrier per
        // tile among the warps of this row quarter; it also orders this tile's P writes, which land on other threads' score
        // columns, after everybody's loads)
        mx_buf[g & 1][part][row] = fmaxf(mxa, mxb);
        TRACE(tr, 5);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * SPLIT) : "memory");
        TRACE(tr, 6);
        float mx = mx_buf[g & 1][0][row];
#pragma unroll
        for (int sp = 1; sp < SPLIT; ++sp) mx = fmaxf(mx, mx_buf[g & 1][sp][row]);
        mx *= p.scale_log2;
        const bool need = mx > m + RESCALE_TAU;  // also the first finite maximum (m == -inf); identical in all threads of a row
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? exp2f(m - mx) : 1.f;
          if (need) m = mx;
          l *= alpha;
          if (g != g_item0) {
            mbar_wait(&pv_done, (g - 1) & 1);  // O holds every earlier P.V of this item once the latest one has landed
            tc_fence_after();
            uint32_t t[32];  // this thread's share of the row's O columns
            const uint32_t to = tmem_O + ob * 128 + lane_addr + part * 32;
            tmem_ld_32x32b_x32(to, t);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) t[e] = __float_as_uint(__uint_as_float(t[e]) * alpha);
            tmem_st_32x32b_x32(to, t);
          }
        }
        const float neg_m = (m == -INFINITY) ? 0.f : -m;
        const unsigned long long scale2 = pack2(p.scale_log2, p.scale_log2), negm2 = pack2(neg_m, neg_m);
        unsigned long long lsA = pack2(0.f, 0.f), lsB = lsA;
        const int pv_chunks = (ncols + 31) >> 5;  // chunks the P.V steps read (16 columns per step)
        if (c < pv_chunks) {
          uint32_t pk[16];
          if (live) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              float x0, x1, x2, x3, p2, p3;
              unpack2(fma2(pack2u(r[e], r[e + 1]), scale2, negm2), x0, x1);
              unpack2(fma2(pack2u(r[e + 2], r[e + 3]), scale2, negm2), x2, x3);
              const float p0 = exp2f(x0), p1 = exp2f(x1);  // exp unit
#if SDAG_PAIR_POLY
              exp2_poly2(x2, x3, p2, p3);                  // FMA pipe
#else
              p2 = exp2f(x2), p3 = exp2f(x3);
#endif
              lsA = add2(lsA, pack2(p0, p1));
              lsB = add2(lsB, pack2(p2, p3));
              pk[e >> 1] = pack_bf16x2(p0, p1);
              pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = 0u;
          }
          tmem_st_32x32b_x16(tS + c * 16, pk);
        }
        {
          float a0, a1;
          unpack2(add2(lsA, lsB), a0, a1);
          l += a0 + a1;
        }
        TRACE(tr, 7);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_pfull0 + buf * 8);
        TRACE(tr, 4);
      }
      // hand the partial row sums to the epilogue warpgroup and move on
      if (nq >= 2) mbar_wait(&l_empty[ob], ((nq >> 1) - 1) & 1);
      l_buf[ob][part][row] = (m == -INFINITY) ? 0.f : l;  // a row without any allowed key: the epilogue writes mean(V)
      mbar_arrive(&l_full[ob]);
      ++nq;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody may still address the peer's shared memory / TMEM when a CTA tears down
  if (warp == 17) tmem_dealloc_2cta(tmem, 512);
}

int launch(const CUtensorMap& tq, const CUtensorMap& tq64, const CUtensorMap& tk, const CUtensorMap& tv128, const PrefillParams& p,
           int cap_qtiles, cudaStream_t st) {
  static bool configured[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (!configured[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(sdag_prefill_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev] = true;
  }
  const int n_sm = device_sm_count();
  const long long n_items = static_cast<long long>(cap_qtiles) * (p.H / 2) * p.B;
  if (n_items <= 0) return SDAG_ERR_INVALID;
  const long long clusters = n_items < n_sm / 2 ? n_items : n_sm / 2;
  sdag_prefill_pair_kernel<<<static_cast<unsigned>(2 * clusters), THREADS, SMEM_TOTAL, st>>>(tq, tq64, tk, tv128, p, n_items);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pairk

}  // namespace

extern "C" {

#ifdef SDAG_DEBUG_TIMING
int sdag_debug_set_timing(unsigned long long* buf) {
  return static_cast<int>(cudaMemcpyToSymbol(g_timing, &buf, sizeof(buf)));
}
int sdag_debug_set_trace(unsigned long long* buf) {  // TRACE_CTAS * 2 * TRACE_EV entries
  return static_cast<int>(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
}
#endif

size_t sdag_prefill_workspace_bytes(int B, int Hkv, int D) {
  if (B <= 0 || Hkv <= 0 || D <= 0) return 0;
  return sizeof(float) * static_cast<size_t>(B) * Hkv * (VSUM_CHUNKS + 1) * D + 16;  // chunk partials + the mean + scheduler counter
}

int sdag_prefill_fwd(const void* q, int64_t sq_b, int64_t sq_h, int64_t sq_l, const void* k, int64_t sk_b,
                     int64_t sk_h, int64_t sk_l, const void* v, int64_t sv_b, int64_t sv_h, int64_t sv_l, void* out,
                     int64_t so_b, int64_t so_l, int64_t so_h, int B, int H, int Hkv, int Lq, int Skv, int D,
                     float scaling, const int32_t* blockmaps_dev, size_t blockmap_stride_bytes, int cap_qtiles,
                     void* workspace, size_t workspace_bytes, int flags, void* stream) {
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

  CUtensorMap tq, tq64, tq32, tk, tv;
  // batch stride of a size-1 dim is irrelevant to addressing but must still be a legal stride
  auto bs = [](int64_t s, int64_t fallback) { return s > 0 ? s : fallback; };
  if (make_map_4d(&tq, q, D, Lq, H, B, sq_l, sq_h, bs(sq_b, sq_l * Lq), BM)) return SDAG_ERR_DRIVER;
  if (make_map_4d(&tq64, q, D, Lq, H, B, sq_l, sq_h, bs(sq_b, sq_l * Lq), 64)) return SDAG_ERR_DRIVER;
  if (make_map_4d(&tq32, q, D, Lq, H, B, sq_l, sq_h, bs(sq_b, sq_l * Lq), 32)) return SDAG_ERR_DRIVER;
  if (make_map_4d(&tk, k, D, Skv, Hkv, B, sk_l, sk_h, bs(sk_b, sk_l * Skv), 64)) return SDAG_ERR_DRIVER;  // 64-row halves
  if (make_map_4d(&tv, v, D, Skv, Hkv, B, sv_l, sv_h, bs(sv_b, sv_l * Skv), 64)) return SDAG_ERR_DRIVER;

  // CTA-pair kernel: head_dim 128 and query heads in pairs per kv head; SDAG_PREFILL_FORCE_SINGLE / _FORCE_PAIR pick a
  // kernel explicitly (tests, A/B), otherwise SDAG_ATTN_PAIR=0 / 1 (read once) or the default below decides
  static int pair_env = -2;
  if (pair_env == -2) {
    const char* e = getenv("SDAG_ATTN_PAIR");
    pair_env = e ? (atoi(e) ? 1 : 0) : -1;
  }
  const bool pair_ok = D == 128 && ((H / Hkv) % 2) == 0;
  if ((flags & SDAG_PREFILL_FORCE_PAIR) && !pair_ok) return SDAG_ERR_UNSUPPORTED;
  const bool use_pair = pair_ok && !(flags & SDAG_PREFILL_FORCE_SINGLE) &&
                        ((flags & SDAG_PREFILL_FORCE_PAIR) || (pair_env >= 0 ? pair_env == 1 : SDAG_PREFILL_PAIR_DEFAULT != 0));
  CUtensorMap tv128;
  if (use_pair && make_map_4d(&tv128, v, D, Skv, Hkv, B, sv_l, sv_h, bs(sv_b, sv_l * Skv), 128)) return SDAG_ERR_DRIVER;

  PrefillParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.so_b = so_b, p.so_l = so_l, p.so_h = so_h;
  p.blockmaps = blockmaps_dev;
  p.bm_stride_ints = static_cast<long long>(blockmap_stride_bytes / 4);
  p.vsum = static_cast<const float*>(workspace);
  p.next_item = reinterpret_cast<int*>(static_cast<float*>(workspace) + static_cast<size_t>(B) * Hkv * (VSUM_CHUNKS + 1) * D);
  p.B = B, p.H = H, p.Hkv = Hkv, p.Lq = Lq, p.Skv = Skv;
  p.scale_log2 = scaling * 1.4426950408889634f;

  const dim3 vgrid(VSUM_CHUNKS, Hkv, B);
  if (flags & SDAG_PREFILL_NO_UNIFORM_ROWS) {
    // caller guarantees every query row has an allowed key (causal / bidirectional maps): no mean-of-V pass, only the
    // work-stealing counter has to be reset
    SDAG_CUDA_TRY(cudaMemsetAsync(p.next_item, 0, sizeof(int), st));
    if (use_pair) return pairk::launch(tq, tq64, tk, tv128, p, cap_qtiles, st);
    if (D == 128) return launch_prefill<128>(tq, tq64, tq32, tk, tv, p, cap_qtiles, st);
    return launch_prefill<64>(tq, tq64, tq32, tk, tv, p, cap_qtiles, st);
  }
  if (D == 128) {
    v_colsum_kernel<128><<<vgrid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(v), sv_b, sv_h, sv_l, Skv,
                                                static_cast<float*>(workspace), blockmaps_dev, p.bm_stride_ints);
    v_mean_kernel<128><<<dim3(Hkv, B), 128, 0, st>>>(static_cast<float*>(workspace), B, Hkv, Skv, blockmaps_dev, p.bm_stride_ints,
                                                     p.out, p.so_b, p.so_l, p.so_h, H, Lq);
    SDAG_CUDA_TRY(cudaGetLastError());
    if (use_pair) return pairk::launch(tq, tq64, tk, tv128, p, cap_qtiles, st);
    return launch_prefill<128>(tq, tq64, tq32, tk, tv, p, cap_qtiles, st);
  }
  v_colsum_kernel<64><<<vgrid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(v), sv_b, sv_h, sv_l, Skv,
                                             static_cast<float*>(workspace), blockmaps_dev, p.bm_stride_ints);
  v_mean_kernel<64><<<dim3(Hkv, B), 64, 0, st>>>(static_cast<float*>(workspace), B, Hkv, Skv, blockmaps_dev, p.bm_stride_ints, p.out,
                                                 p.so_b, p.so_l, p.so_h, H, Lq);
  SDAG_CUDA_TRY(cudaGetLastError());
  return launch_prefill<64>(tq, tq64, tq32, tk, tv, p, cap_qtiles, st);
}

}  // extern "C"
