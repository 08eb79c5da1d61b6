// SDAG block map and dense-mask kernels (path A).
//
// The reference builds a dense bool [L,L] allow-mask with one Python slice assignment per row
// (src/pipeline/sparse_attention_RAG/SDAG.py:68-127) and turns it into an additive [1,1,L,L] bias
// (SDAG.py:164-173).  Here the same predicate is kept as a list of 128x128 tiles per query-row tile,
// computed on the device from the document token offsets; tiles the predicate forbids entirely are
// never listed, so the attention kernel skips them instead of masking them.
#include "common.cuh"

namespace {

constexpr int FULL_DOFF = 1 << 20;
constexpr int KV_FLAG_BITMASK = 1;
constexpr int MAX_SEG = 2 * SDAG_MAX_DOCS + 3;
enum SegKind { SEG_SYS = 0, SEG_DOC = 1, SEG_GAP = 2, SEG_QA = 3 };

struct Layout {
  int L, S, qa, n_docs, use_nbrs;
  const int32_t* spans;
  const int32_t* nbrs;
  int nbr_k;
};

__device__ __forceinline__ Layout read_layout(const int32_t* lay, int max_docs, int nbr_k) {
  Layout o;
  o.L = lay[0];
  o.S = lay[1];
  o.qa = lay[2];
  o.n_docs = lay[3];
  o.use_nbrs = lay[4];
  o.spans = lay + SDAG_LAYOUT_HEADER;
  o.nbrs = lay + SDAG_LAYOUT_HEADER + 2 * max_docs;
  o.nbr_k = nbr_k;
  return o;
}

// Python slice-bound normalisation on an axis of length n (negative wraps once, then clips).
__device__ __forceinline__ int py_clip(int a, int n) {
  if (a < 0) {
    a += n;
    if (a < 0) a = 0;
  }
  return a < n ? a : n;
}

// ------------------------------------------------------------------------------------------------
// Dense mask (the build_blocked_causal_mask_full seam).  One block per row; the row's allowed
// column intervals are collected in shared memory, then written as bytes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_mask_kernel(const int32_t* __restrict__ lay_raw, int max_docs,
                                                         int nbr_k, uint8_t* __restrict__ mask, int L) {
  const Layout lay = read_layout(lay_raw, max_docs, nbr_k);
  const int i = blockIdx.x;
  // documents whose span holds row i (one for layouts the prompt builder produces; overlapping spans may give several —
  // bounded by n_docs <= SDAG_MAX_DOCS, so nothing is ever dropped), then every element evaluates the predicate itself
  __shared__ int docs_of_row[SDAG_MAX_DOCS];
  __shared__ int n_docs_of_row;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int d = 0; d < lay.n_docs && d < SDAG_MAX_DOCS; ++d) {  // SDAG.py:113-116 (row range of the slice assignment)
      const int ds = lay.spans[2 * d], de = lay.spans[2 * d + 1];
      if (i >= max(ds, 0) && i < min(de, L)) docs_of_row[n++] = d;
    }
    n_docs_of_row = n;
  }
  __syncthreads();
  const int n = n_docs_of_row;
  const bool causal_row = i < min(max(lay.S, 0), L) || i >= max(lay.qa, 0);  // SDAG.py:107-108, 124-125
  const int s_cols = py_clip(lay.S, L);
  uint8_t* row = mask + static_cast<size_t>(i) * L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    bool a = causal_row && j <= i;
    for (int t = 0; t < n && !a; ++t) {
      const int d = docs_of_row[t];
      a = j < s_cols || (j >= py_clip(lay.spans[2 * d], L) && j <= i);  // SDAG.py:117-118
      if (!a && lay.use_nbrs) {
        for (int u = 0; u < lay.nbr_k && !a; ++u) {  // SDAG.py:119-122
          const int nb = lay.nbrs[d * lay.nbr_k + u];
          if (nb < 0 || nb >= lay.n_docs) continue;
          a = j >= py_clip(lay.spans[2 * nb], L) && j < py_clip(lay.spans[2 * nb + 1], L);
        }
      }
    }
    row[j] = a ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Block map from document offsets
// ------------------------------------------------------------------------------------------------
struct Interval {
  int lo, hi, causal;
};

// Column intervals (disjoint, ascending) allowed for query rows [q0, q1) of a segment.
__device__ int row_intervals(const Layout& lay, int kind, int doc, int q0, int q1, Interval* iv) {
  int n = 0;
  if (kind == SEG_GAP) return 0;
  if (kind == SEG_SYS || kind == SEG_QA) {
    iv[0] = {0, q1, 1};
    return 1;
  }
  // document rows: system block, own span (causal), neighbour spans (full, non-causal — Appendix A3)
  if (lay.S > 0) iv[n++] = {0, lay.S, 0};
  bool self_full = false;
  if (lay.use_nbrs) {
    for (int t = 0; t < lay.nbr_k; ++t) {
      const int nb = lay.nbrs[doc * lay.nbr_k + t];
      if (nb < 0 || nb >= lay.n_docs) continue;
      if (nb == doc) {
        self_full = true;
        continue;
      }
      bool dup = false;
      for (int u = 0; u < t; ++u) dup |= (lay.nbrs[doc * lay.nbr_k + u] == nb);
      if (dup) continue;
      const int lo = lay.spans[2 * nb], hi = lay.spans[2 * nb + 1];
      if (hi > lo) iv[n++] = {lo, hi, 0};
    }
  }
  const int ds = lay.spans[2 * doc], de = lay.spans[2 * doc + 1];
  if (self_full)
    iv[n++] = {ds, de, 0};
  else
    iv[n++] = {ds, q1, 1};
  // insertion sort by lo
  for (int a = 1; a < n; ++a) {
    Interval x = iv[a];
    int b = a - 1;
    while (b >= 0 && iv[b].lo > x.lo) {
      iv[b + 1] = iv[b];
      --b;
    }
    iv[b + 1] = x;
  }
  // merge adjacent intervals that one tile rule can express: full+full, or full (entirely at or left of
  // the first row, hence also causal-true) followed by causal
  int m = 0;
  for (int a = 0; a < n; ++a) {
    if (m > 0 && iv[m - 1].hi == iv[a].lo) {
      Interval& p = iv[m - 1];
      if (!p.causal && !iv[a].causal) {
        p.hi = iv[a].hi;
        continue;
      }
      if (!p.causal && iv[a].causal && p.hi <= q0 + 1) {
        p.hi = iv[a].hi;
        p.causal = 1;
        continue;
      }
    }
    iv[m++] = iv[a];
  }
  return m;
}

// Emit (or just count when out == nullptr) the kv tiles of a query tile; returns count, adds allowed pairs.
__device__ int emit_kv_tiles(const Interval* iv, int n_iv, int q0, int nrows, int4* out, unsigned long long* pairs) {
  int cnt = 0;
  unsigned long long pr = 0;
  for (int a = 0; a < n_iv; ++a) {
    int hi = iv[a].hi;
    if (iv[a].causal) hi = min(hi, q0 + nrows);
    for (int k0 = iv[a].lo; k0 < hi; k0 += SDAG_TILE) {
      const int ncols = min(SDAG_TILE, hi - k0);
      const int doff = iv[a].causal ? (q0 - k0) : FULL_DOFF;
      if (out) {
        out[cnt] = make_int4(k0, ncols, doff, 0);
        if (iv[a].causal) {
          for (int r = 0; r < nrows; ++r) pr += static_cast<unsigned>(max(0, min(ncols, r + doff + 1)));
        } else {
          pr += static_cast<unsigned long long>(nrows) * ncols;
        }
      }
      ++cnt;
    }
  }
  if (out && pairs) atomicAdd(pairs, pr);
  return cnt;
}

__global__ void __launch_bounds__(256) build_blockmap_kernel(const int32_t* __restrict__ layouts, int layout_stride,
                                                             int max_docs, int nbr_k, int32_t* __restrict__ maps,
                                                             long long map_stride_ints, int cap_q, int cap_kv) {
  const Layout lay = read_layout(layouts + static_cast<long long>(blockIdx.x) * layout_stride, max_docs, nbr_k);
  int32_t* map = maps + static_cast<long long>(blockIdx.x) * map_stride_ints;
  int4* qtiles = reinterpret_cast<int4*>(map + SDAG_BLOCKMAP_HEADER);
  int4* kvtiles = qtiles + cap_q;
  int4* scratch = kvtiles + cap_kv;  // unsorted qtiles

  __shared__ int seg_kind[MAX_SEG], seg_lo[MAX_SEG], seg_hi[MAX_SEG], seg_doc[MAX_SEG], seg_toff[MAX_SEG + 1];
  __shared__ int n_seg_s, n_tiles_s, status_s, gap_rows_s;
  __shared__ int part[256];
  __shared__ unsigned long long pairs_s;
  const int tid = threadIdx.x;

  if (tid == 0) {
    int status = 0, n = 0, gaps = 0;
    const int L = lay.L, S = lay.S;
    int qa = lay.qa;
    if (qa > L) qa = L;
    if (L <= 0 || S < 0 || S > L || lay.n_docs < 0 || lay.n_docs > SDAG_MAX_DOCS || nbr_k > SDAG_MAX_NBR_K || qa < S)
      status = 1;
    int pos = S;
    if (!status) {
      if (S > 0) {
        seg_kind[n] = SEG_SYS, seg_lo[n] = 0, seg_hi[n] = S, seg_doc[n] = -1;
        ++n;
      }
      for (int d = 0; d < lay.n_docs && !status; ++d) {
        const int ds = lay.spans[2 * d], de = lay.spans[2 * d + 1];
        if (ds < pos || de < ds || de > qa) {
          status = 1;
          break;
        }
        if (ds > pos) {
          seg_kind[n] = SEG_GAP, seg_lo[n] = pos, seg_hi[n] = ds, seg_doc[n] = -1;
          gaps += ds - pos;
          ++n;
        }
        if (de > ds) {
          seg_kind[n] = SEG_DOC, seg_lo[n] = ds, seg_hi[n] = de, seg_doc[n] = d;
          ++n;
        }
        pos = de;
      }
    }
    if (!status) {
      if (qa > pos) {
        seg_kind[n] = SEG_GAP, seg_lo[n] = pos, seg_hi[n] = qa, seg_doc[n] = -1;
        gaps += qa - pos;
        ++n;
      }
      if (L > qa) {
        seg_kind[n] = SEG_QA, seg_lo[n] = qa, seg_hi[n] = L, seg_doc[n] = -1;
        ++n;
      }
    }
    int t = 0;
    for (int s = 0; s < n; ++s) {
      seg_toff[s] = t;
      t += (seg_hi[s] - seg_lo[s] + SDAG_TILE - 1) / SDAG_TILE;
    }
    seg_toff[n] = t;
    if (!status && t > cap_q) status = 2;
    n_seg_s = n;
    n_tiles_s = status ? 0 : t;
    status_s = status;
    gap_rows_s = gaps;
    pairs_s = 0;
  }
  __syncthreads();
  const int T = n_tiles_s;
  const int n_seg = n_seg_s;

  // pass 1: kv-tile count per query tile
  Interval iv[SDAG_MAX_NBR_K + 2];
  int local = 0;
  for (int t = tid; t < T; t += blockDim.x) {
    int s = 0;
    while (s + 1 < n_seg && seg_toff[s + 1] <= t) ++s;
    const int q0 = seg_lo[s] + (t - seg_toff[s]) * SDAG_TILE;
    const int nrows = min(SDAG_TILE, seg_hi[s] - q0);
    const int n_iv = row_intervals(lay, seg_kind[s], seg_doc[s], q0, q0 + nrows, iv);
    const int cnt = emit_kv_tiles(iv, n_iv, q0, nrows, nullptr, nullptr);
    scratch[t] = make_int4(q0, nrows, 0, cnt);
    local += cnt;
  }
  // exclusive scan of counts in tile order: thread-strided layout makes a direct scan awkward, so
  // thread 0 walks the (at most a few thousand) tiles.
  __syncthreads();
  if (tid == 0) {
    int off = 0;
    for (int t = 0; t < T; ++t) {
      int4 q = scratch[t];
      q.z = off;
      off += q.w;
      scratch[t] = q;
    }
    part[0] = off;
    if (off > cap_kv) {
      status_s = 2;
    }
  }
  __syncthreads();
  const int total_kv = part[0];
  const bool ok = status_s == 0;
  (void)local;

  // pass 2: fill kv tiles
  if (ok) {
    for (int t = tid; t < T; t += blockDim.x) {
      int s = 0;
      while (s + 1 < n_seg && seg_toff[s + 1] <= t) ++s;
      const int4 q = scratch[t];
      const int n_iv = row_intervals(lay, seg_kind[s], seg_doc[s], q.x, q.x + q.y, iv);
      emit_kv_tiles(iv, n_iv, q.x, q.y, kvtiles + q.z, &pairs_s);
    }
  }
  __syncthreads();
  // pass 3: order query tiles by descending kv count (longest first) — rank sort
  if (ok) {
    for (int t = tid; t < T; t += blockDim.x) {
      const int4 q = scratch[t];
      int rank = 0;
      for (int u = 0; u < T; ++u) {
        const int c = scratch[u].w;
        rank += (c > q.w) || (c == q.w && u < t);
      }
      qtiles[rank] = q;
    }
  }
  if (tid == 0) {
    map[0] = ok ? T : 0;
    map[1] = ok ? total_kv : 0;
    map[2] = status_s;
    map[3] = gap_rows_s;
    map[4] = lay.L;
    map[5] = cap_q;
    map[6] = cap_kv;
    map[7] = 0;
    map[8] = static_cast<int>(pairs_s & 0xffffffffull);
    map[9] = static_cast<int>(pairs_s >> 32);
    map[10] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Block map from an arbitrary dense mask: bit-pack + per-tile occupancy, then list non-empty tiles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mask_pack_kernel(const uint8_t* __restrict__ mask, long long row_stride, int L,
                                                        uint32_t* __restrict__ bits, int words_per_row,
                                                        int32_t* __restrict__ tile_any, int32_t* __restrict__ row_any) {
  const int tk = blockIdx.x, tq = blockIdx.y;
  const int row = tq * SDAG_TILE + threadIdx.x;
  uint32_t any = 0;
  if (row < L) {
    const uint8_t* src = mask + static_cast<long long>(row) * row_stride;
    for (int w = 0; w < 4; ++w) {
      uint32_t word = 0;
      const int c0 = tk * SDAG_TILE + w * 32;
      for (int b = 0; b < 32; ++b) {
        const int c = c0 + b;
        if (c < L && src[c]) word |= (1u << b);
      }
      bits[static_cast<long long>(row) * words_per_row + tk * 4 + w] = word;
      any |= word;
    }
    if (any) atomicOr(&row_any[row], 1);
  }
  const int block_any = __syncthreads_or(any != 0);
  if (threadIdx.x == 0) tile_any[tq * gridDim.x + tk] = block_any;
}

__global__ void __launch_bounds__(256) maskmap_build_kernel(int32_t* __restrict__ map, int L, int cap_q, int cap_kv,
                                                            int words_per_row, int bits_off, const int32_t* tile_any,
                                                            const int32_t* row_any) {
  int4* qtiles = reinterpret_cast<int4*>(map + SDAG_BLOCKMAP_HEADER);
  int4* kvtiles = qtiles + cap_q;
  int4* scratch = kvtiles + cap_kv;
  const int T = (L + SDAG_TILE - 1) / SDAG_TILE;
  __shared__ int total_s, gaps_s;
  const int tid = threadIdx.x;
  if (tid == 0) gaps_s = 0;
  __syncthreads();
  for (int t = tid; t < T; t += blockDim.x) {
    int cnt = 0;
    for (int u = 0; u < T; ++u) cnt += tile_any[t * T + u] != 0;
    scratch[t] = make_int4(t * SDAG_TILE, min(SDAG_TILE, L - t * SDAG_TILE), 0, cnt);
  }
  int g = 0;
  for (int r = tid; r < L; r += blockDim.x) g += row_any[r] == 0;
  if (g) atomicAdd(&gaps_s, g);
  __syncthreads();
  if (tid == 0) {
    int off = 0;
    for (int t = 0; t < T; ++t) {
      int4 q = scratch[t];
      q.z = off;
      off += q.w;
      scratch[t] = q;
    }
    total_s = off;
  }
  __syncthreads();
  for (int t = tid; t < T; t += blockDim.x) {
    const int4 q = scratch[t];
    int c = 0;
    for (int u = 0; u < T; ++u)
      if (tile_any[t * T + u]) kvtiles[q.z + c++] = make_int4(u * SDAG_TILE, min(SDAG_TILE, L - u * SDAG_TILE), 0, KV_FLAG_BITMASK);
    int rank = 0;
    for (int u = 0; u < T; ++u) {
      const int cu = scratch[u].w;
      rank += (cu > q.w) || (cu == q.w && u < t);
    }
    qtiles[rank] = q;
  }
  if (tid == 0) {
    map[0] = T;
    map[1] = total_s;
    map[2] = 0;
    map[3] = gaps_s;
    map[4] = L;
    map[5] = cap_q;
    map[6] = cap_kv;
    map[7] = words_per_row;
    map[8] = 0;
    map[9] = 0;
    map[10] = bits_off;
  }
}

__global__ void popcount_pairs_kernel(const uint32_t* bits, long long n_words, int32_t* map) {
  unsigned long long acc = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc += __popc(bits[i]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(reinterpret_cast<unsigned long long*>(map + 8), acc);
}

inline int cap_qtiles(int L, int n_docs) { return (L + SDAG_TILE - 1) / SDAG_TILE + 2 * n_docs + 3; }
inline int cap_kvtiles(int L, int n_docs, int nbr_k) {
  return cap_qtiles(L, n_docs) * ((L + SDAG_TILE - 1) / SDAG_TILE + nbr_k + 3);
}

}  // namespace

extern "C" {

int sdag_layout_ints(int max_docs, int nbr_k) { return SDAG_LAYOUT_HEADER + 2 * max_docs + max_docs * nbr_k; }

size_t sdag_blockmap_bytes(int L, int n_docs, int nbr_k) {
  if (L < 0 || n_docs < 0 || nbr_k < 0) return 0;
  const size_t cq = cap_qtiles(L, n_docs), ckv = cap_kvtiles(L, n_docs, nbr_k);
  return sizeof(int32_t) * (SDAG_BLOCKMAP_HEADER + 4 * (2 * cq + ckv));
}

int sdag_build_mask(const int32_t* layout_dev, int max_docs, int nbr_k, uint8_t* mask_dev, int L, void* stream) {
  if (!layout_dev || !mask_dev || L < 0 || max_docs < 0 || max_docs > SDAG_MAX_DOCS || nbr_k < 0) return SDAG_ERR_INVALID;
  if (L == 0) return SDAG_OK;
  dense_mask_kernel<<<L, 256, 0, static_cast<cudaStream_t>(stream)>>>(layout_dev, max_docs, nbr_k, mask_dev, L);
  return static_cast<int>(cudaGetLastError());
}

int sdag_build_blockmap(const int32_t* layouts_dev, int n_prompts, int max_docs, int nbr_k, int max_L,
                        int32_t* blockmaps_dev, size_t blockmap_stride_bytes, void* stream) {
  if (!layouts_dev || !blockmaps_dev || n_prompts <= 0 || max_docs < 0 || nbr_k < 0 || max_L <= 0) return SDAG_ERR_INVALID;
  if (max_docs > SDAG_MAX_DOCS || nbr_k > SDAG_MAX_NBR_K) return SDAG_ERR_UNSUPPORTED;
  if (blockmap_stride_bytes < sdag_blockmap_bytes(max_L, max_docs, nbr_k) || blockmap_stride_bytes % 16) return SDAG_ERR_WORKSPACE;
  build_blockmap_kernel<<<n_prompts, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      layouts_dev, sdag_layout_ints(max_docs, nbr_k), max_docs, nbr_k, blockmaps_dev,
      static_cast<long long>(blockmap_stride_bytes / 4), cap_qtiles(max_L, max_docs), cap_kvtiles(max_L, max_docs, nbr_k));
  return static_cast<int>(cudaGetLastError());
}

// layout of the mask-map buffer: header | qtiles[T] | kvtiles[T*T] | scratch[T] | tile_any[T*T] | row_any[L] | bits
size_t sdag_maskmap_bytes(int L) {
  if (L <= 0) return 0;
  const size_t T = (L + SDAG_TILE - 1) / SDAG_TILE;
  size_t ints = SDAG_BLOCKMAP_HEADER + 4 * (2 * T + T * T) + T * T + L;
  ints = (ints + 3) & ~size_t(3);
  ints += static_cast<size_t>(L) * T * 4;
  return ints * 4;
}

int sdag_blockmap_from_mask(const uint8_t* mask_dev, int64_t row_stride, int L, int32_t* blockmap_dev,
                            size_t blockmap_bytes, void* stream) {
  if (!mask_dev || !blockmap_dev || L <= 0 || row_stride < L) return SDAG_ERR_INVALID;
  if (blockmap_bytes < sdag_maskmap_bytes(L)) return SDAG_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = (L + SDAG_TILE - 1) / SDAG_TILE;
  const size_t tile_any_off = SDAG_BLOCKMAP_HEADER + 4 * (2 * static_cast<size_t>(T) + static_cast<size_t>(T) * T);
  const size_t row_any_off = tile_any_off + static_cast<size_t>(T) * T;
  const size_t bits_off = (row_any_off + L + 3) & ~size_t(3);
  int32_t* tile_any = blockmap_dev + tile_any_off;
  int32_t* row_any = blockmap_dev + row_any_off;
  uint32_t* bits = reinterpret_cast<uint32_t*>(blockmap_dev + bits_off);
  cudaError_t e = cudaMemsetAsync(row_any, 0, sizeof(int32_t) * L, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  mask_pack_kernel<<<dim3(T, T), 128, 0, st>>>(mask_dev, row_stride, L, bits, T * 4, tile_any, row_any);
  maskmap_build_kernel<<<1, 256, 0, st>>>(blockmap_dev, L, T, T * T, T * 4, static_cast<int>(bits_off), tile_any, row_any);
  const long long n_words = static_cast<long long>(L) * T * 4;
  popcount_pairs_kernel<<<148, 256, 0, st>>>(bits, n_words, blockmap_dev);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
