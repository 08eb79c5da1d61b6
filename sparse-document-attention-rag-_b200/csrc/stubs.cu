// TEMPORARY: entry points not implemented yet (each returns SDAG_ERR_UNSUPPORTED).
#include "common.cuh"
extern "C" {
size_t sdag_decode_workspace_bytes(int, int, int, int) { return 0; }
int sdag_decode_fwd(const void*, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t,
                    int64_t, void*, int, int, int, int, int, float, void*, size_t, void*) { return SDAG_ERR_UNSUPPORTED; }
int ip_cast_rows_bf16(const float*, void*, float*, int64_t, int, void*) { return SDAG_ERR_UNSUPPORTED; }
size_t ip_search_workspace_bytes(int, int, int, int) { return 0; }
int ip_search(const float*, int, const float*, const void*, int64_t, int, int64_t, const float*, int, float*, int64_t*,
              int32_t*, void*, size_t, int, void*) { return SDAG_ERR_UNSUPPORTED; }
int topk_merge(const float*, const int64_t*, int, int, int, float*, int64_t*, void*) { return SDAG_ERR_UNSUPPORTED; }
int centroid_select(const float*, const int32_t*, int, const float*, const int32_t*, int, int, int, int, float*, int32_t*,
                    void*) { return SDAG_ERR_UNSUPPORTED; }
int doc_knn(const float*, int, int, int, int32_t*, void*) { return SDAG_ERR_UNSUPPORTED; }
}
