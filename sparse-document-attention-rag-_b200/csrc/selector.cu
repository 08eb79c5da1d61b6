// Path B selector and path A2 doc-kNN: small fp32 kernels on embeddings that already sit in HBM.
//
// centroid_select replaces the numeric core of _select_malicious_docs_for_query
// (src/pipeline/attack/malicious_selection.py:81-128): centroid = mean of the retrieved embeddings (not
// re-normalised), sims = cand . centroid, order = argsort(-sims) ("closest") or argsort(sims) ("furthest"),
// equal scores keep the lower index first (np.argmax / np.argmin first-occurrence rule).
// doc_knn replaces compute_doc_knn_for_docs' similarity + argsort (SDAG.py:50-63).
// One CTA per query / per document row; everything is HBM-latency bound and tiny.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int MAX_D = 4096;
constexpr int MAX_ITEMS = 1024;

__device__ __forceinline__ float warp_sum(float x) {
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__global__ void __launch_bounds__(256)
centroid_select_kernel(const float* __restrict__ retrieved, const int32_t* __restrict__ r_count, int R_max,
                       const float* __restrict__ cand, const int32_t* __restrict__ c_count, int C_max, int d, int furthest,
                       float* __restrict__ sims_out, int32_t* __restrict__ order_out) {
  __shared__ float centroid[MAX_D];
  __shared__ float sims[MAX_ITEMS];
  const int qi = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = min(r_count[qi], R_max), Cn = min(c_count[qi], C_max);
  const float* rbase = retrieved + static_cast<long long>(qi) * R_max * d;
  const float* cbase = cand + static_cast<long long>(qi) * C_max * d;
  for (int j = tid; j < d; j += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += rbase[static_cast<long long>(r) * d + j];  // row order, as numpy's axis-0 mean
    centroid[j] = R > 0 ? s / static_cast<float>(R) : 0.f;
  }
  __syncthreads();
  for (int c = warp; c < Cn; c += blockDim.x / 32) {
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s = fmaf(cbase[static_cast<long long>(c) * d + j], centroid[j], s);
    s = warp_sum(s);
    if (lane == 0) sims[c] = s;
  }
  __syncthreads();
  for (int c = tid; c < C_max; c += blockDim.x) {
    sims_out[static_cast<long long>(qi) * C_max + c] = c < Cn ? sims[c] : 0.f;
    order_out[static_cast<long long>(qi) * C_max + c] = -1;
  }
  __syncthreads();
  for (int c = tid; c < Cn; c += blockDim.x) {
    const float s = sims[c];
    int rank = 0;
    for (int j = 0; j < Cn; ++j) {
      const float t = sims[j];
      const bool better = furthest ? (t < s) : (t > s);
      rank += better || (t == s && j < c);
    }
    order_out[static_cast<long long>(qi) * C_max + rank] = c;
  }
}

__global__ void __launch_bounds__(256)
doc_knn_kernel(const float* __restrict__ emb, int n, int d, int k_neighbors, int32_t* __restrict__ nbrs_out) {
  __shared__ float sims[MAX_ITEMS];
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* er = emb + static_cast<long long>(r) * d;
  for (int j = warp; j < n; j += blockDim.x / 32) {
    float s = 0.f;
    for (int t = lane; t < d; t += 32) s = fmaf(er[t], emb[static_cast<long long>(j) * d + t], s);
    s = warp_sum(s);
    if (lane == 0) sims[j] = s;
  }
  for (int t = tid; t < k_neighbors; t += blockDim.x) nbrs_out[static_cast<long long>(r) * k_neighbors + t] = -1;
  __syncthreads();
  for (int j = tid; j < n; j += blockDim.x) {
    if (j == r) continue;  // self removed by index (SDAG.py:57-58)
    const float s = sims[j];
    int rank = 0;
    for (int u = 0; u < n; ++u) {
      if (u == r) continue;
      const float t = sims[u];
      rank += (t > s) || (t == s && u < j);
    }
    if (rank < k_neighbors) nbrs_out[static_cast<long long>(r) * k_neighbors + rank] = j;
  }
}

}  // namespace

extern "C" {

int centroid_select(const float* retrieved, const int32_t* r_count, int R_max, const float* cand, const int32_t* c_count,
                    int C_max, int nq, int d, int furthest, float* sims_out, int32_t* order_out, void* stream) {
  if (!retrieved || !r_count || !cand || !c_count || !sims_out || !order_out) return SDAG_ERR_INVALID;
  if (nq <= 0 || d <= 0 || R_max <= 0 || C_max <= 0) return SDAG_ERR_INVALID;
  if (d > MAX_D || C_max > MAX_ITEMS) return SDAG_ERR_UNSUPPORTED;
  centroid_select_kernel<<<nq, 256, 0, static_cast<cudaStream_t>(stream)>>>(retrieved, r_count, R_max, cand, c_count, C_max,
                                                                            d, furthest ? 1 : 0, sims_out, order_out);
  return static_cast<int>(cudaGetLastError());
}

int doc_knn(const float* emb, int n, int d, int k_neighbors, int32_t* nbrs_out, void* stream) {
  if (!emb || !nbrs_out || n <= 0 || d <= 0 || k_neighbors <= 0) return SDAG_ERR_INVALID;
  if (n > MAX_ITEMS) return SDAG_ERR_UNSUPPORTED;
  doc_knn_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(emb, n, d, k_neighbors, nbrs_out);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
