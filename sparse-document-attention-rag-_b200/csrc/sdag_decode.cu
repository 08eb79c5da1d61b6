// SDAG decode attention for sm_100a (path A5): one new token attends to every cached position — the
// reference passes no mask after the prefill (SDAG.py:193-197), so decode is dense over the KV cache and
// HBM-bound.  Split-KV over a TMA ring: each CTA streams a contiguous range of cache rows of one kv head
// for all G = H/Hkv query heads sharing it (K/V bytes are read exactly once) and writes an (m, l, acc)
// partial; a second kernel merges the partials.  The arithmetic rides on warp-level MMAs only to keep the
// instruction count out of the way of the copy engine: 2*G FLOP per cache byte is far below the tensor ridge.
#include <math.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace {

using namespace sm100;

constexpr int TILE_ROWS = 128;   // cache rows per ring stage (16 per consumer warp)
constexpr int MAX_G = 8;         // query heads per kv head (the M rows of the warp MMA that are live)
constexpr int CONSUMERS = 8;     // consumer warps
constexpr int THREADS = (CONSUMERS + 1) * 32;  // + one TMA producer warp

template <int D>
struct DecodeSmem {
  static constexpr int kTileBytes = TILE_ROWS * D * 2;           // one K (or V) tile
  static constexpr int kStageBytes = 2 * kTileBytes;             // K then V
  static constexpr int kStages = D == 128 ? 3 : 6;               // 192 KiB of ring either way
  static constexpr int kTotal = kStages * kStageBytes + 1024;    // + alignment slack (128B-swizzle atoms)
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col).  Only rows 0..7 of A are live here (a1 = a3 = 0).
__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Split-KV decode: CTA (split, hkv, b) streams its share of the cache rows of one kv head through a TMA ring
// (K and V tiles of 128 rows, 128-byte swizzle) exactly once for all G query heads of the group.  The ranges are cut
// inside the kernel from the device-side cache length, so every CTA of a replayed CUDA graph keeps an equal share as
// the cache grows.  Each consumer warp owns 16 rows of every tile and runs its own online softmax on warp-level
// m16n8k16 MMAs (S = Q K^T with the G heads as the M rows, O += P V) — no block-wide synchronisation in the loop;
// the eight warp states are merged through shared memory at the end and written as one (m, l, acc[D]) partial.
template <int D>
__global__ void __launch_bounds__(THREADS, 1)
decode_split_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __nv_bfloat16* __restrict__ q, long long sq_b, long long sq_h, float* __restrict__ partial,
                    int H, int Hkv, int kv_len_host, float scale_log2, const int32_t* __restrict__ kv_len_dev,
                    int kv_len_dev_stride, __nv_bfloat16* __restrict__ merged_out, unsigned int* __restrict__ tickets,
                    int early_tiles) {
  using S = DecodeSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[S::kStages], empty[S::kStages];

  const int split = blockIdx.x, nsplit = gridDim.x, hkv = blockIdx.y, b = blockIdx.z;
  // Launched as a programmatic dependent.  What the kernel just before this one may still be writing is q and the cache row
  // of the NEWEST token (row kv_len - 1: the q/k/v projection of this step); every other cache row and the length were
  // written at least one step (several kernels) ago, and each kernel of the chain waits for its own predecessor before it
  // lets its dependents start.  So the ring is filled with tiles that lie wholly below row kv_len - 1 BEFORE
  // griddepcontrol.wait (early_tiles; the length is read past L1): the stream starts under the previous kernel's tail.
  if (!early_tiles) pdl_wait();
  // device-side length: one value for the batch (stride 0) or one per prompt of a ragged batch (stride 1)
  const int kv_len = kv_len_dev ? max(0, min(__ldcg(kv_len_dev + b * kv_len_dev_stride), kv_len_host)) : kv_len_host;
  const int G = H / Hkv;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_total = (kv_len + TILE_ROWS - 1) / TILE_ROWS;
  const int tiles_per_split = (tiles_total + nsplit - 1) / nsplit;
  const int t0 = min(tiles_total, split * tiles_per_split);
  const int t1 = min(tiles_total, t0 + tiles_per_split);
  const int n_tiles = t1 - t0;

  if (tid == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMERS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  __syncthreads();

  if (warp == CONSUMERS) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      auto issue = [&](int it) {
        const int s = it % S::kStages;
        if (it >= S::kStages) mbar_wait(&empty[s], ((it / S::kStages) - 1) & 1);
        uint8_t* stage = smem + s * S::kStageBytes;
        const int row0 = (t0 + it) * TILE_ROWS;
        mbar_expect_tx(&full[s], S::kStageBytes);
        for (int dh = 0; dh < D / 64; ++dh) tma_load_4d(stage + dh * TILE_ROWS * 128, &tmK, &full[s], dh * 64, row0, hkv, b);
        for (int dh = 0; dh < D / 64; ++dh)
          tma_load_4d(stage + S::kTileBytes + dh * TILE_ROWS * 128, &tmV, &full[s], dh * 64, row0, hkv, b);
      };
      int it = 0;
      if (early_tiles) {
        for (; it < n_tiles && it < S::kStages && (t0 + it + 1) * TILE_ROWS <= kv_len - 1; ++it) issue(it);
        pdl_wait();
      }
      pdl_trigger();  // the merge kernel may be scheduled; it waits for this grid before reading the partial results
      for (; it < n_tiles; ++it) issue(it);
    }
    return;
  }

  // ======================= consumers =======================
  if (early_tiles) pdl_wait();  // q comes from the kernel just before this one
  pdl_trigger();
  const int g = lane >> 2, qd = lane & 3;  // MMA row (query head within the group) / quad lane
  // A fragments of Q for every 16-wide slice of d: a0 = Q[g][16ks + 2qd .. +1], a2 = Q[g][16ks + 8 + 2qd .. +1]
  uint32_t qa[D / 16][2];
  {
    const __nv_bfloat16* qrow = q + b * sq_b + static_cast<long long>(hkv * G + g) * sq_h;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      qa[ks][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + qd * 2) : 0u;
      qa[ks][1] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + qd * 2) : 0u;
    }
  }
  float o[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m = -INFINITY, l = 0.f;  // running maximum of row g (quad-uniform), this thread's share of the row sum

  // ldmatrix source rows inside the warp's 16-row slice of a tile (see the fragment tables of mma.m16n8k16)
  const int wrow = warp * 16;
  const int krow = wrow + ((lane >> 4) << 3) + (lane & 7);  // K: matrices (n-tile, d-half) = lane / 8 -> (lane/16, (lane/8)&1)
  const int kch = (lane >> 3) & 1;
  const int vrow = wrow + (((lane >> 3) & 1) << 3) + (lane & 7);  // V: matrices (kv-half, d-tile) = ((lane/8)&1, lane/16)
  const int vch = lane >> 4;

  for (int it = 0; it < n_tiles; ++it) {
    const int s = it % S::kStages;
    mbar_wait(&full[s], (it / S::kStages) & 1);
    const uint32_t sK = smem_u32(smem + s * S::kStageBytes);
    const uint32_t sV = sK + S::kTileBytes;
    const int row0 = (t0 + it) * TILE_ROWS;
    if (row0 + TILE_ROWS > kv_len) {
      // tail tile: rows at or beyond the cache length may hold anything (a preallocated cache): zero this warp's V rows
      // there so that P = 0 really contributes nothing; their scores are masked below
      uint8_t* vt = smem + s * S::kStageBytes + S::kTileBytes;
      for (int e = lane; e < 16 * (D / 8); e += 32) {
        const int r = wrow + e / (D / 8), c = e % (D / 8);
        if (row0 + r >= kv_len)
          *reinterpret_cast<uint4*>(vt + (c >> 3) * TILE_ROWS * 128 + r * 128 + ((c & 7) << 4)) = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
    }

    // ---- S[g][kv] for the warp's 16 rows: two n-tiles of 8 ----
    float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      const int c = ks * 2 + kch;  // 16-byte chunk of the row
      uint32_t b00, b01, b10, b11;
      ldmatrix_x4(sK + (c >> 3) * TILE_ROWS * 128 + sw128_offset(krow, c & 7), b00, b01, b10, b11);
      mma_16816(sc[0], qa[ks][0], qa[ks][1], b00, b01);
      mma_16816(sc[1], qa[ks][0], qa[ks][1], b10, b11);
    }
    // this thread: row g, columns kv = row0 + wrow + 8j + 2qd + {0,1}
    float p[2][2];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = row0 + wrow + j * 8 + qd * 2 + e < kv_len;
        p[j][e] = ok ? sc[j][e] * scale_log2 : -INFINITY;
        tmax = fmaxf(tmax, p[j][e]);
      }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const bool grow = tmax > m;
    if (__any_sync(0xffffffffu, grow)) {
      const float alpha = grow ? exp2f(m - tmax) : 1.f;  // m = -inf -> 0 (nothing accumulated yet)
      if (grow) m = tmax;
      l *= alpha;
#pragma unroll
      for (int n = 0; n < D / 8; ++n) o[n][0] *= alpha, o[n][1] *= alpha;
    }
    const float mu = m == -INFINITY ? 0.f : m;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        p[j][e] = exp2f(p[j][e] - mu);
        l += p[j][e];
      }
    const uint32_t pa0 = pack_bf16x2(p[0][0], p[0][1]);  // A[g][kv 2qd..]   of the 16 x 16 P block
    const uint32_t pa2 = pack_bf16x2(p[1][0], p[1][1]);  // A[g][kv 8+2qd..]

    // ---- O[g][d] += P[g][kv] V[kv][d]: V^T fragments straight from the [kv][d] rows (ldmatrix.trans) ----
#pragma unroll
    for (int n2 = 0; n2 < D / 16; ++n2) {
      const int c = n2 * 2 + vch;
      uint32_t b00, b01, b10, b11;
      ldmatrix_x4_trans(sV + (c >> 3) * TILE_ROWS * 128 + sw128_offset(vrow, c & 7), b00, b01, b10, b11);
      mma_16816(o[n2 * 2], pa0, pa2, b00, b01);
      mma_16816(o[n2 * 2 + 1], pa0, pa2, b10, b11);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ---- merge the eight warp states: red[w][g] = {m, l}, acc[w][g][d] (the ring is drained: reuse it) ----
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS * 32) : "memory");
  float* red_ml = reinterpret_cast<float*>(smem);                       // [CONSUMERS][MAX_G][2]
  float* red_o = red_ml + CONSUMERS * MAX_G * 2;                        // [CONSUMERS][MAX_G][D]
  if (qd == 0) {
    red_ml[(warp * MAX_G + g) * 2] = m;
    red_ml[(warp * MAX_G + g) * 2 + 1] = l;
  }
#pragma unroll
  for (int n = 0; n < D / 8; ++n)
    *reinterpret_cast<float2*>(red_o + (warp * MAX_G + g) * D + n * 8 + qd * 2) = make_float2(o[n][0], o[n][1]);
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS * 32) : "memory");
  // partial record per (b, hkv, split, head): [m, l, acc[D]]
  float* out = partial + ((static_cast<long long>(b) * Hkv + hkv) * nsplit + split) * G * (D + 2);
  for (int i = tid; i < G * D; i += CONSUMERS * 32) {
    const int gg = i / D, d = i % D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < CONSUMERS; ++w) M = fmaxf(M, red_ml[(w * MAX_G + gg) * 2]);
    float acc = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < CONSUMERS; ++w) {
      const float mw = red_ml[(w * MAX_G + gg) * 2];
      const float sc = mw == -INFINITY ? 0.f : exp2f(mw - M);
      acc = fmaf(sc, red_o[(w * MAX_G + gg) * D + d], acc);
      den = fmaf(sc, red_ml[(w * MAX_G + gg) * 2 + 1], den);
    }
    out[gg * (D + 2) + 2 + d] = acc;
    if (d == 0) {
      out[gg * (D + 2)] = M;
      out[gg * (D + 2) + 1] = den;
    }
  }
  if (tickets == nullptr) return;  // merged by decode_merge_kernel

  // ---- the split CTA that finishes last merges the records of its (b, hkv): no separate merge launch ----
  __shared__ int is_last_s;
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS * 32) : "memory");
  if (tid == 0) {
    unsigned int* t = tickets + b * Hkv + hkv;
    const unsigned int ticket = atomicAdd(t, 1u);
    is_last_s = ticket == static_cast<unsigned int>(nsplit - 1);
    if (is_last_s) *t = 0u;  // left clean for the next call
  }
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS * 32) : "memory");
  if (!is_last_s) return;
  __threadfence();
  const float* rec = partial + (static_cast<long long>(b) * Hkv + hkv) * nsplit * G * (D + 2);
  const int stride = G * (D + 2);
  for (int i = tid; i < G * D; i += CONSUMERS * 32) {
    const int gg = i / D, d = i % D;
    const float* base = rec + gg * (D + 2);  // other CTAs wrote these: read through L2
    float M = -INFINITY;
#pragma unroll 4
    for (int sp = 0; sp < nsplit; ++sp) M = fmaxf(M, __ldcg(base + sp * stride));
    float num = 0.f, den = 0.f;
#pragma unroll 4
    for (int sp = 0; sp < nsplit; ++sp) {
      const float mw = __ldcg(base + sp * stride);
      const float w = mw == -INFINITY ? 0.f : exp2f(mw - M);
      num = fmaf(w, __ldcg(base + sp * stride + 2 + d), num);
      den = fmaf(w, __ldcg(base + sp * stride + 1), den);
    }
    merged_out[(static_cast<long long>(b) * H + hkv * G + gg) * D + d] = __float2bfloat16(num / den);
  }
}

constexpr int MAX_TICKETS = 4096;
__device__ unsigned int g_decode_tickets[MAX_TICKETS];  // zero at load; every call leaves them zero

constexpr int MAX_SPLITS = 256;  // >= SM count: split_count never exceeds it

template <int D>
__global__ void __launch_bounds__(D)
decode_merge_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out, int H, int Hkv, int nsplit) {
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  pdl_wait();     // the split kernel's partial results
  pdl_trigger();  // the o_proj GEMV that follows may start streaming its weights; it waits for this grid before reading `out`
  const int G = H / Hkv, hkv = h / G, g = h % G;
  const float* base = partial + (static_cast<long long>(b) * Hkv + hkv) * nsplit * G * (D + 2) + g * (D + 2);
  const long long stride = static_cast<long long>(G) * (D + 2);
  // the split weights once per head (threads = splits), then D independent column sums with the loads batched
  __shared__ float w_s[MAX_SPLITS];
  __shared__ float den_s;
  if (d < 32) {
    float M = -INFINITY;
    for (int s = d; s < nsplit; s += 32) M = fmaxf(M, base[s * stride]);
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    float den = 0.f;
    for (int s = d; s < nsplit; s += 32) {
      const float m = base[s * stride];
      const float w = m == -INFINITY ? 0.f : exp2f(m - M);
      w_s[s] = w;
      den = fmaf(w, base[s * stride + 1], den);
    }
    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    if (d == 0) den_s = den;
  }
  __syncthreads();
  float num = 0.f;
#pragma unroll 8
  for (int s = 0; s < nsplit; ++s) num = fmaf(w_s[s], base[s * stride + 2 + d], num);
  const float den = den_s;
  out[(static_cast<long long>(b) * H + h) * D + d] = __float2bfloat16(num / den);
}

int sm_count() {
  return device_sm_count();
}

// One CTA per SM (the ring takes the whole shared memory): as many splits as fill the SMs once, never more than tiles.
inline int split_count(int B, int Hkv, int kv_len) {
  int n = sm_count() / (B * Hkv);
  const int tiles = (kv_len + TILE_ROWS - 1) / TILE_ROWS;
  if (n > tiles) n = tiles;
  if (n > MAX_SPLITS) n = MAX_SPLITS;
  return n < 1 ? 1 : n;
}

int make_kv_map(CUtensorMap* m, const void* base, int D, int rows, int heads, int B, int64_t s_l, int64_t s_h, int64_t s_b) {
  uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(rows), static_cast<uint64_t>(heads),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(s_l) * 2, static_cast<uint64_t>(s_h) * 2,
                         static_cast<uint64_t>(s_b > 0 ? s_b : s_l * rows) * 2};
  uint32_t box[4] = {64, TILE_ROWS, 1, 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box);
}

template <int D>
int launch_decode(const CUtensorMap& tk, const CUtensorMap& tv, const __nv_bfloat16* q, int64_t sq_b, int64_t sq_h,
                  float* part, __nv_bfloat16* out, int B, int H, int Hkv, int kv_len, int nsplit, float sl2,
                  const int32_t* kv_len_dev, int kv_len_dev_stride, cudaStream_t st) {
  static bool cfg[SDAG_MAX_DEVICES] = {};
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (!cfg[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(decode_split_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       DecodeSmem<D>::kTotal));
    cfg[dev] = true;
  }
  // both kernels are programmatic dependents of their predecessor: launch latency hides behind the previous kernel's tail
  unsigned int* tickets = nullptr;
  static int early_tiles = -1;
  if (early_tiles < 0) {
    const char* e = getenv("SDAG_DECODE_EARLY_TILES");  // 0: wait for the previous kernel before the first cache tile is requested
    early_tiles = (e && !atoi(e)) ? 0 : 1;
  }
  static int fused_merge = -1;
  if (fused_merge < 0) {
    const char* e = getenv("SDAG_DECODE_SPLIT_MERGE");  // 1: the last split CTA merges instead of decode_merge_kernel
    fused_merge = (e && atoi(e)) ? 1 : 0;  // measured: the serial tail in 8 CTAs costs 3 us per layer more than the dependent launch
  }
  if (fused_merge && B * Hkv <= MAX_TICKETS)
    SDAG_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&tickets), g_decode_tickets));
  SDAG_CUDA_TRY(launch_pdl(decode_split_kernel<D>, dim3(nsplit, Hkv, B), dim3(THREADS), DecodeSmem<D>::kTotal, st, tk, tv, q,
                           sq_b, sq_h, part, H, Hkv, kv_len, sl2, kv_len_dev, kv_len_dev_stride, out, tickets, early_tiles));
  if (tickets != nullptr) return SDAG_OK;  // the last split CTA of every (b, hkv) merged in place
  return static_cast<int>(launch_pdl(decode_merge_kernel<D>, dim3(H, B), dim3(D), 0, st, static_cast<const float*>(part), out, H,
                                     Hkv, nsplit));
}

}  // namespace

extern "C" {

size_t sdag_decode_workspace_bytes(int B, int H, int D, int kv_len) {
  if (B <= 0 || H <= 0 || D <= 0 || kv_len <= 0) return 0;
  // worst case over Hkv (Hkv = 1 gives the most splits per head): one (m, l, acc[D]) record per head and split
  const size_t max_split = static_cast<size_t>(split_count(1, 1, kv_len));
  return sizeof(float) * static_cast<size_t>(B) * H * max_split * (D + 2);
}

int sdag_decode_fwd(const void* q, int64_t sq_b, int64_t sq_h, const void* k, int64_t sk_b, int64_t sk_h, int64_t sk_l,
                    const void* v, int64_t sv_b, int64_t sv_h, int64_t sv_l, void* out, int B, int H, int Hkv, int kv_len,
                    int D, float scaling, void* workspace, size_t workspace_bytes, const int32_t* kv_len_dev, int kv_len_dev_stride,
                    void* stream) {
  if (!q || !k || !v || !out || !workspace) return SDAG_ERR_INVALID;
  if (kv_len_dev_stride != 0 && kv_len_dev_stride != 1) return SDAG_ERR_INVALID;
  if (B <= 0 || H <= 0 || Hkv <= 0 || kv_len <= 0 || H % Hkv) return SDAG_ERR_INVALID;
  if (D != 64 && D != 128) return SDAG_ERR_UNSUPPORTED;
  if (H / Hkv > MAX_G) return SDAG_ERR_UNSUPPORTED;
  if ((sk_l % 8) || (sv_l % 8) || (sk_h % 8) || (sv_h % 8) || (sk_b % 8) || (sv_b % 8)) return SDAG_ERR_INVALID;
  if ((sq_b % 2) || (sq_h % 2) || (reinterpret_cast<uintptr_t>(q) & 3)) return SDAG_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) return SDAG_ERR_INVALID;
  const int nsplit = split_count(B, Hkv, kv_len);
  const size_t need = sizeof(float) * static_cast<size_t>(B) * H * nsplit * (D + 2);
  if (workspace_bytes < need) return SDAG_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sl2 = scaling * 1.4426950408889634f;
  CUtensorMap tk, tv;
  if (make_kv_map(&tk, k, D, kv_len, Hkv, B, sk_l, sk_h, sk_b)) return SDAG_ERR_DRIVER;
  if (make_kv_map(&tv, v, D, kv_len, Hkv, B, sv_l, sv_h, sv_b)) return SDAG_ERR_DRIVER;
  auto* qq = static_cast<const __nv_bfloat16*>(q);
  float* part = static_cast<float*>(workspace);
  auto* oo = static_cast<__nv_bfloat16*>(out);
  if (D == 128)
    return launch_decode<128>(tk, tv, qq, sq_b, sq_h, part, oo, B, H, Hkv, kv_len, nsplit, sl2, kv_len_dev, kv_len_dev_stride, st);
  return launch_decode<64>(tk, tv, qq, sq_b, sq_h, part, oo, B, H, Hkv, kv_len, nsplit, sl2, kv_len_dev, kv_len_dev_stride, st);
}

}  // extern "C"
