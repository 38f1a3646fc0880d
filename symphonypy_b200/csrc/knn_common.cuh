// Shared between the FFMA scan (knn.cu) and the tcgen05 scan (knn_tc.cu).
#pragma once
#include "common.cuh"

namespace sym {

struct KnnHeader {
    float rmax2;       // max ||ref||^2 (written by fit with an integer atomicMax: values are >= 0)
    int32_t has_order;  // 1 when the reference was packed in a caller-given (locality) order
    // single-FP16 operand section (method 3): rows are scaled by the power of two `scale16` (largest norm in
    // (8, 16], far from FP16 overflow and underflow); dr16 = max over rows of ||s r - fp16(s r)||, the exact
    // rounding residual the re-rank's error bound is built from
    float scale16;
    float dr16;
    int32_t pad[60];
};

// Layout of the opaque "packed reference" buffer produced by sym_knn_fit:
//   [KnnHeader | norm2 f32[Mpad] | refT f32[d4][Mpad] | BF16x3 tiles | FP16 tiles | order i32[Mpad] | tile_c | tile_r]
struct KnnPacked {
    KnnHeader *hdr;
    float *norm2;
    float *refT;
    unsigned char *tc;    // tensor-core operand section, BF16x3 (layout private to knn_tc.cu)
    unsigned char *tc16;  // tensor-core operand section, single FP16
    int32_t *order;     // [Mpad] packed position -> original reference row (identity when no order was given)
    // bounding ball of every 128-position tile of the packed reference (opt-in exact tile pruning of the tcgen05 scans)
    float *tile_c;      // [Mpad / 128][d4] centre (mean of the tile's rows)
    float *tile_r;      // [Mpad / 128] radius (largest distance of a row to the centre, rounded up)
    int64_t Mpad;
    int d4;
    size_t total;
};

__host__ __device__ inline int64_t knn_mpad(int64_t M) { return (M + 255) / 256 * 256; }

size_t knn_tc_packed_bytes(int64_t M, int d, int prec);   // prec 0 = BF16x3, 1 = single FP16

inline KnnPacked knn_packed_layout(void *base, int64_t M, int d) {
    KnnPacked p;
    p.Mpad = knn_mpad(M);
    p.d4 = (d + 3) & ~3;
    unsigned char *b = (unsigned char *)base;
    size_t off = 0;
    p.hdr = (KnnHeader *)(b + off);
    off += sizeof(KnnHeader);
    p.norm2 = (float *)(b + off);
    off += (size_t)p.Mpad * sizeof(float);
    p.refT = (float *)(b + off);
    off += (size_t)p.d4 * p.Mpad * sizeof(float);
    off = (off + 1023) & ~(size_t)1023;
    p.tc = b + off;
    off += knn_tc_packed_bytes(M, d, 0);
    off = (off + 1023) & ~(size_t)1023;
    p.tc16 = b + off;
    off += knn_tc_packed_bytes(M, d, 1);
    off = (off + 255) & ~(size_t)255;
    p.order = (int32_t *)(b + off);
    off += (size_t)p.Mpad * sizeof(int32_t);
    off = (off + 255) & ~(size_t)255;
    p.tile_c = (float *)(b + off);
    off += (size_t)(p.Mpad / 128) * p.d4 * sizeof(float);
    p.tile_r = (float *)(b + off);
    off += (size_t)(p.Mpad / 128) * sizeof(float);
    off = (off + 255) & ~(size_t)255;
    p.total = off;
    return p;
}

// Order-preserving map float -> uint32, packed with the reference index: comparing keys compares
// (score, index) lexicographically, which makes selection independent of arrival order and sends
// exact distance ties to the lower reference index.
__host__ __device__ __forceinline__ unsigned long long knn_make_key(float s, uint32_t idx) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(s);
#else
    uint32_t u;
    memcpy(&u, &s, 4);
#endif
    u ^= (u & 0x80000000u) ? 0xffffffffu : 0x80000000u;
    return ((unsigned long long)u << 32) | idx;
}
__device__ __forceinline__ float knn_key_score(unsigned long long key) {
    if (key == ~0ull) return INFINITY;  // empty slot
    uint32_t u = (uint32_t)(key >> 32);
    u ^= (u & 0x80000000u) ? 0x80000000u : 0xffffffffu;
    return __uint_as_float(u);
}

// tcgen05 scan entry points (knn_tc.cu)
// `prec`: 0 = BF16x3 split operands (FP32-class scores), 1 = single FP16 operands (three times fewer MMAs;
// the scan reports the exact rounding residual of every query row in *dq16 for the re-rank's error bound)
bool knn_tc_supported(int64_t N, int64_t M, int d, int k, int prec);
size_t knn_tc_workspace(int64_t N, int64_t M, int d, int k, int prec, int prune = 1);
int knn_tc_fit(const float *Ref, int ldr, int64_t M, int d, const int32_t *order, KnnPacked pk, cudaStream_t st);
// `prune`: skip reference tiles whose bounding ball cannot hold a candidate of any row of the CTA (exact; opt-in).
// The first 16 bytes of the workspace then receive {tiles scanned, tiles a full scan visits} (two uint64).
int knn_tc_scan(const float *Q, int ldq, int64_t N, const int32_t *q_perm, const int32_t *q_code,
                const int32_t *cl_tile, KnnPacked pk, int64_t M, int d, int k, int prec, int prune, void *workspace,
                size_t workspace_bytes, cudaStream_t st, const unsigned long long **cand, const float **thr, int *ncand,
                int *nsplit, int *group, double *eps_rel, const float **dq16, const unsigned char **qrows, int *kp);

}  // namespace sym
