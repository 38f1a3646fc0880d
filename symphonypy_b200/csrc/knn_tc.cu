// tcgen05 (5th-gen tensor core) candidate scan for the kNN — placeholder until the kernel lands.
#include "knn_common.cuh"

namespace sym {
size_t knn_tc_packed_bytes(int64_t, int) { return 0; }
bool knn_tc_supported(int64_t, int64_t, int, int) { return false; }
size_t knn_tc_workspace(int64_t, int64_t, int, int) { return 0; }
int knn_tc_fit(const float *, int, int64_t, int, KnnPacked, cudaStream_t) { return SYM_OK; }
int knn_tc_scan(const float *, int, int64_t, KnnPacked, int64_t, int, int, void *, size_t, cudaStream_t,
                const int32_t **, const float **, int *, int *, double *) {
    set_error("tcgen05 scan not built");
    return SYM_ERR_UNSUPPORTED_SHAPE;
}
}  // namespace sym
