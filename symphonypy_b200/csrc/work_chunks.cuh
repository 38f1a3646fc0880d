// Work list over cells grouped by an integer code (output of sym_batch_order): block -> (level, first position,
// count).  Shared by the MoE statistics / apply kernels (moe.cu) and the per-group moments (confidence.cu).
#pragma once
#include <cstdint>

namespace sym {

constexpr int MOE_CH = 256;       // cells per work chunk (must match the header: "ceil(group/256)")

struct Chunk {
    int level;
    int64_t pos;  // first position in perm (or first cell id when perm == nullptr)
    int n;        // cells in this chunk (0 = nothing to do)
};

__device__ __forceinline__ Chunk find_chunk(int64_t blk, int n_levels, const int64_t *__restrict__ seg,
                                            const int32_t *__restrict__ chunk_start, int64_t N) {
    Chunk c{0, 0, 0};
    if (seg == nullptr) {  // single level, identity order
        const int64_t p = blk * MOE_CH;
        if (p < N) {
            c.pos = p;
            c.n = (int)(N - p < MOE_CH ? N - p : MOE_CH);
        }
        return c;
    }
    if (blk >= chunk_start[n_levels]) return c;
    int lo = 0, hi = n_levels;  // last level with chunk_start[level] <= blk
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_start[mid] <= blk) lo = mid; else hi = mid;
    }
    const int64_t p = seg[lo] + (blk - chunk_start[lo]) * MOE_CH;
    const int64_t e = seg[lo + 1];
    c.level = lo;
    c.pos = p;
    c.n = (int)(e - p < MOE_CH ? e - p : MOE_CH);
    return c;
}

}  // namespace sym
