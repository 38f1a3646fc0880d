// (4) brute-force Euclidean kNN + majority vote                  — symphonypy/tl.py:425-433
// (sklearn KNeighborsClassifier: EuclideanArgKmin + uniform vote)
//
// Pipeline:   fit (pack reference once)  ->  scan (over-select kc = k + pad candidates per query
// row, distance matrix never materialised)  ->  rerank (float64, sklearn's formula, (dist, index)
// order, certifies the over-selection margin)  ->  vote.
//
// This file holds the FP32 FFMA scan (method 1) plus fit / rerank / vote, which every method
// shares.  The tcgen05 scan (method 2) is in knn_tc.cu.
//
// FFMA scan: a CTA owns 128 query rows and walks the reference in 128-column tiles staged by
// cp.async (3 stages of [32 dims x 128 refs]); each thread keeps an 8x8 register tile of
// s = ||r||^2 - 2 q.r.  A candidate is examined further only if s <= the row's current kc-th best
// (one compare per pair in the steady state); survivors go through a small per-row queue in
// shared memory into the row's top-kc list (replace-the-maximum, warp cooperative).
#include <cuda_fp16.h>
#include <float.h>

#include "common.cuh"
#include "knn_common.cuh"

namespace sym {

// ------------------------------------------------------------------------------------------ fit
// RefT[dd][m] = Ref[m][dd] (zero padded to d4 x Mpad), norm2[m] = ||Ref[m]||^2 (NaN for padding),
// header.rmax2 = max_m norm2[m].
__global__ void __launch_bounds__(256)
knn_fit_kernel(const float *__restrict__ Ref, int ldr, int64_t M, int d, const int32_t *__restrict__ order,
               KnnPacked pk) {
    __shared__ float tile[128][33];
    const int64_t m0 = (int64_t)blockIdx.x * 128;
    const int tid = threadIdx.x;
    float nrm = 0.f;  // threads 0..127 accumulate the norm of row m0 + tid
    for (int d0 = 0; d0 < pk.d4; d0 += 32) {
        __syncthreads();
        for (int i = tid; i < 128 * 32; i += 256) {
            const int r = i / 32, c = i % 32;
            const int64_t m = m0 + r;
            const int64_t src = (m < M && order) ? order[m] : m;  // packed position m holds reference row order[m]
            tile[r][c] = (m < M && d0 + c < d) ? Ref[src * ldr + d0 + c] : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < 32 * 128; i += 256) {
            const int c = i / 128, r = i % 128;
            if (d0 + c < pk.d4) pk.refT[(int64_t)(d0 + c) * pk.Mpad + m0 + r] = tile[r][c];
        }
        if (tid < 128)
            for (int c = 0; c < 32; ++c) nrm = fmaf(tile[tid][c], tile[tid][c], nrm);
    }
    if (tid < 128) {
        const bool real = m0 + tid < M;
        pk.norm2[m0 + tid] = real ? nrm : __int_as_float(0x7fc00000);  // NaN: padding never passes a compare
        pk.order[m0 + tid] = real ? (order ? order[m0 + tid] : (int32_t)(m0 + tid)) : -1;
        // max over the block -> header (float atomic max through the int ordering of non-negatives)
        float v = real ? nrm : 0.f;
        v = warp_max(v);
        if ((tid & 31) == 0) atomicMax(reinterpret_cast<int *>(&pk.hdr->rmax2), __float_as_int(v));
    }
}

// ------------------------------------------------------------------------------------------ scan
constexpr int SC_THREADS = 256;
constexpr int SC_BQ = 128;
constexpr int SC_BR = 128;
constexpr int SC_DC = 32;     // dims per pipeline stage
constexpr int SC_STAGES = 3;
constexpr int SC_QCAP = 8;    // per-row candidate queue

struct ScanSmem {
    float *Qs;                 // [d4][SC_BQ]     (-2 q, transposed)
    float *Rs;                 // [SC_STAGES][SC_DC][SC_BR]
    float *rn;                 // [SC_STAGES][SC_BR]
    float *thr_f;              // [SC_BQ]
    unsigned long long *thr_key;   // [SC_BQ]
    int *thr_pos;              // [SC_BQ]
    int *cnt;                  // [SC_BQ]
    unsigned long long *cand;  // [SC_BQ][SC_QCAP]
    unsigned long long *top;   // [SC_BQ][kcap]  (shared memory when it fits, else global workspace)
};

__host__ __device__ inline size_t scan_smem_bytes(int d4, int kcap, bool top_in_smem) {
    size_t b = (size_t)d4 * SC_BQ * 4 + (size_t)SC_STAGES * SC_DC * SC_BR * 4 + (size_t)SC_STAGES * SC_BR * 4;
    b += SC_BQ * 4 + SC_BQ * 8 + SC_BQ * 4 + SC_BQ * 4 + (size_t)SC_BQ * SC_QCAP * 8;
    if (top_in_smem) b += (size_t)SC_BQ * kcap * 8;
    return b + 16;
}

// warp-cooperative: insert `key` into the row's list if it beats the current maximum
__device__ __forceinline__ void topk_insert(unsigned long long *top_row, int kc, unsigned long long key,
                                            unsigned long long &mx, int &mxpos, int lane) {
    if (key >= mx) return;
    if (lane == 0) top_row[mxpos] = key;
    __syncwarp();
    unsigned long long best = 0ull;
    int bpos = 0;
    for (int j = lane; j < kc; j += 32) {
        const unsigned long long v = top_row[j];
        if (v >= best) { best = v; bpos = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        if (ov > best || (ov == best && op > bpos)) { best = ov; bpos = op; }
    }
    mx = best;
    mxpos = bpos;
}

__global__ void __launch_bounds__(SC_THREADS)
knn_scan_ffma_kernel(const float *__restrict__ Q, int ldq, int64_t N, const int32_t *__restrict__ q_perm, int64_t M,
                     KnnPacked pk, int d, int kc, int kcap,
                     int tiles_per_split, int top_in_smem, unsigned long long *__restrict__ top_ws,
                     unsigned long long *__restrict__ cand_out, float *__restrict__ thr_out) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int d4 = pk.d4;
    ScanSmem S;
    {
        unsigned char *p = smraw;
        S.cand = (unsigned long long *)p; p += (size_t)SC_BQ * SC_QCAP * 8;
        S.thr_key = (unsigned long long *)p; p += SC_BQ * 8;
        if (top_in_smem) { S.top = (unsigned long long *)p; p += (size_t)SC_BQ * kcap * 8; }
        S.Qs = (float *)p; p += (size_t)d4 * SC_BQ * 4;
        S.Rs = (float *)p; p += (size_t)SC_STAGES * SC_DC * SC_BR * 4;
        S.rn = (float *)p; p += SC_STAGES * SC_BR * 4;
        S.thr_f = (float *)p; p += SC_BQ * 4;
        S.thr_pos = (int *)p; p += SC_BQ * 4;
        S.cnt = (int *)p;
        if (!top_in_smem)
            S.top = top_ws + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * SC_BQ * kcap;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t q0 = (int64_t)blockIdx.x * SC_BQ;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int ntiles_all = (int)(pk.Mpad / SC_BR);
    const int t_beg = split * tiles_per_split;
    const int t_end = min(ntiles_all, t_beg + tiles_per_split);
    const int ntiles = max(0, t_end - t_beg);
    const int nchunks = (d4 + SC_DC - 1) / SC_DC;
    const int nstages = ntiles * nchunks;

    // ---- init: Q tile (scaled by -2, transposed), top lists, thresholds
    for (int i = tid; i < SC_BQ * d4; i += SC_THREADS) {
        const int r = i / d4, c = i % d4;
        const int64_t n = q0 + r;
        const int64_t src = (n < N && q_perm) ? q_perm[n] : n;
        S.Qs[c * SC_BQ + r] = (n < N && c < d) ? -2.f * Q[src * ldq + c] : 0.f;
    }
    for (int i = tid; i < SC_BQ * kcap; i += SC_THREADS) S.top[i] = ~0ull;
    for (int i = tid; i < SC_BQ; i += SC_THREADS) {
        S.thr_f[i] = INFINITY;
        S.thr_key[i] = ~0ull;
        S.thr_pos[i] = 0;
        S.cnt[i] = 0;
    }

    auto issue_stage = [&](int s) {
        if (s < nstages) {
            const int t = t_beg + s / nchunks, ck = s % nchunks;
            const int buf = s % SC_STAGES;
            const int dc0 = ck * SC_DC;
            const int len = min(SC_DC, d4 - dc0);
            float *dst = S.Rs + (size_t)buf * SC_DC * SC_BR;
            const float *src = pk.refT + (int64_t)dc0 * pk.Mpad + (int64_t)t * SC_BR;
            for (int i = tid; i < len * (SC_BR / 4); i += SC_THREADS) {
                const int r = i / (SC_BR / 4), c4 = (i % (SC_BR / 4)) * 4;
                cp_async16(dst + r * SC_BR + c4, src + (int64_t)r * pk.Mpad + c4);
            }
            if (ck == 0 && tid < SC_BR / 4)
                cp_async16(S.rn + (t % SC_STAGES) * SC_BR + tid * 4, pk.norm2 + (int64_t)t * SC_BR + tid * 4);
        }
        cp_async_commit();
    };

    issue_stage(0);
    issue_stage(1);

    float acc[8][8];
    for (int s = 0; s < nstages; ++s) {
        const int t = t_beg + s / nchunks, ck = s % nchunks;
        cp_async_wait<SC_STAGES - 2>();
        __syncthreads();
        issue_stage(s + SC_STAGES - 1);
        if (ck == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        }
        const float *rs = S.Rs + (size_t)(s % SC_STAGES) * SC_DC * SC_BR;
        const float *qs = S.Qs + (size_t)ck * SC_DC * SC_BQ;
        const int len = min(SC_DC, d4 - ck * SC_DC);
#pragma unroll 4
        for (int kk = 0; kk < len; ++kk) {
            const float4 qa = *reinterpret_cast<const float4 *>(qs + kk * SC_BQ + ty * 4);
            const float4 qb = *reinterpret_cast<const float4 *>(qs + kk * SC_BQ + 64 + ty * 4);
            const float4 ra = *reinterpret_cast<const float4 *>(rs + kk * SC_BR + tx * 4);
            const float4 rb = *reinterpret_cast<const float4 *>(rs + kk * SC_BR + 64 + tx * 4);
            const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const float r[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(q[i], r[j], acc[i][j]);
        }
        if (ck != nchunks - 1) continue;

        // ---- selection epilogue for reference tile t
        const float *rn = S.rn + (t % SC_STAGES) * SC_BR;
        float rnv[8];
        {
            const float4 a = *reinterpret_cast<const float4 *>(rn + tx * 4);
            const float4 b = *reinterpret_cast<const float4 *>(rn + 64 + tx * 4);
            rnv[0] = a.x; rnv[1] = a.y; rnv[2] = a.z; rnv[3] = a.w;
            rnv[4] = b.x; rnv[5] = b.y; rnv[6] = b.z; rnv[7] = b.w;
        }
        unsigned long long pend = 0ull;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
            const float th = S.thr_f[row];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[i][j] += rnv[j];
                if (acc[i][j] <= th) pend |= 1ull << (i * 8 + j);
            }
        }
        while (__syncthreads_or(pend != 0ull)) {
            unsigned long long todo = pend;
            while (todo) {
                const int b = __ffsll((long long)todo) - 1;
                todo &= todo - 1;
                const int i = b >> 3, j = b & 7;
                const int row = (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
                // select acc[i][j] without dynamic register indexing
                float sv = 0.f;
#pragma unroll
                for (int ii = 0; ii < 8; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        if (ii == i && jj == j) sv = acc[ii][jj];
                const int col = (j < 4 ? 0 : 64) + tx * 4 + (j & 3);
                const unsigned long long key = knn_make_key(sv, (uint32_t)((int64_t)t * SC_BR + col));
                if (key >= S.thr_key[row]) {
                    pend &= ~(1ull << b);
                    continue;
                }
                const int slot = atomicAdd(&S.cnt[row], 1);
                if (slot < SC_QCAP) {
                    S.cand[row * SC_QCAP + slot] = key;
                    pend &= ~(1ull << b);
                }
            }
            __syncthreads();
            for (int row = warp; row < SC_BQ; row += SC_THREADS / 32) {
                const int c = min(S.cnt[row], SC_QCAP);
                if (c == 0) continue;
                unsigned long long mx = S.thr_key[row];
                int mxpos = S.thr_pos[row];
                unsigned long long *top_row = S.top + (size_t)row * kcap;
                for (int e = 0; e < c; ++e) topk_insert(top_row, kc, S.cand[row * SC_QCAP + e], mx, mxpos, lane);
                __syncwarp();
                if (lane == 0) {
                    S.thr_key[row] = mx;
                    S.thr_pos[row] = mxpos;
                    S.thr_f[row] = knn_key_score(mx);
                    S.cnt[row] = 0;
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- write the candidate sets (unordered) and the row thresholds
    for (int i = tid; i < SC_BQ * kc; i += SC_THREADS) {
        const int r = i / kc, e = i % kc;
        const int64_t n = q0 + r;
        if (n < N) cand_out[(n * nsplit + split) * kc + e] = S.top[(size_t)r * kcap + e];
    }
    // a split with no more than kc reference rows hands all of them over: nothing was left out
    const int64_t split_refs = min(M, (int64_t)t_end * SC_BR) - (int64_t)t_beg * SC_BR;
    for (int r = tid; r < SC_BQ; r += SC_THREADS)
        if (q0 + r < N) thr_out[(q0 + r) * nsplit + split] = split_refs <= kc ? INFINITY : S.thr_f[r];
}

// ------------------------------------------------------------------------------------------ rerank
// One warp per query row.  The scan hands over `ncand` candidates per row; a candidate is a GROUP of `gsz`
// consecutive positions of the packed reference, keyed by the smallest scan score in the group (gsz = 8 for the
// tcgen05 scans, which select groups; gsz = 1 for the FP32 scan).  Every member gets its exact float64 distance
// (sklearn's formula); members whose exact score lies below the row's threshold (plus the error margin) are
// ranked by (distance, index); the certificate is checked.
//   gsz = 8: the 8 members of a group sit next to each other in refT[c][pos .. pos+7] — eight lanes read one
//            32-byte sector per dimension;  gsz = 1: the member's row Ref[order[pos]] is read.
constexpr int RR_THREADS = 256;
constexpr int RR_MAXC = 512;   // candidates per row the scan may hand over, and members per row that can be ranked
constexpr int RR_DMAX = 128;
constexpr size_t RR_SMEM = (size_t)(RR_THREADS / 32) * (RR_MAXC * (sizeof(double) + sizeof(int32_t)) + RR_DMAX * sizeof(float));
// groups kernel: 256 survivors per row are plenty (~kc survive pass A; a row with more — hundreds of references tied at
// its threshold — is flagged and goes to the next tier, whose last one ranks up to RR_MAXC single references); the
// smaller lists let 5 instead of 4 blocks share an SM, which is what this latency-bound kernel needs
constexpr int RG_MAXC = 256;
constexpr size_t RR_SMEM_GROUPS = (size_t)(RR_THREADS / 32) * (RG_MAXC * (sizeof(double) + sizeof(int32_t)) + RR_DMAX * sizeof(double) + 256 * 2);

__global__ void __launch_bounds__(RR_THREADS)
knn_rerank_kernel(const float *__restrict__ Q, int ldq, int64_t N, const float *__restrict__ Ref, int ldr,
                  const float *__restrict__ refT, int64_t Mpad, int64_t M, int d, int k,
                  const unsigned long long *__restrict__ cand, const float *__restrict__ thr, int ncand, int nsplit,
                  int gsz, double eps_rel, const KnnHeader *__restrict__ hdr, const int32_t *__restrict__ order,
                  const int32_t *__restrict__ q_perm, const float *__restrict__ dq16, int32_t *__restrict__ idx_out,
                  float *__restrict__ dist_out, uint8_t *__restrict__ unsafe) {
    extern __shared__ __align__(16) unsigned char rr_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wd = reinterpret_cast<double *>(rr_smem) + (size_t)warp * RR_MAXC;
    int32_t *wi = reinterpret_cast<int32_t *>(rr_smem + (size_t)(RR_THREADS / 32) * RR_MAXC * sizeof(double)) + (size_t)warp * RR_MAXC;
    float *wq = reinterpret_cast<float *>(rr_smem + (size_t)(RR_THREADS / 32) * RR_MAXC * (sizeof(double) + sizeof(int32_t))) + (size_t)warp * RR_DMAX;
    const int64_t row = (int64_t)blockIdx.x * (RR_THREADS / 32) + warp;  // row of the scan (locality order)
    if (row >= N) return;
    const int64_t n = q_perm ? q_perm[row] : row;                        // the query it stands for
    const float *q = Q + n * ldq;
    for (int c = lane; c < d; c += 32) wq[c] = q[c];
    __syncwarp();
    double qn = 0.0;
    for (int c = 0; c < d; ++c) {
        const double v = (double)wq[c];
        qn = fma(v, v, qn);
    }
    const double rmax2 = (double)hdr->rmax2;
    double eps = eps_rel * (2.0 * sqrt(qn) * sqrt(rmax2) + rmax2);
    double unscale = 1.0;   // scan scores and thresholds are in units of scale16^2 when the operands were single FP16
    if (dq16 != nullptr) {
        // single-FP16 scan, in scaled units Q' = 2 s q, R' = s r:  score' = ||R'||^2 - Q'.R',  computed from
        // fp16(Q') = Q' - dQ, fp16(R') = R' - dR.  |error| <= ||Q'|| ||dR|| + ||dQ|| (||R'|| + ||dR||), with ||dQ|| of
        // this row measured by the pack kernel and ||dR|| <= header.dr16 (max over the reference).
        const double s = (double)hdr->scale16, dr = (double)hdr->dr16, dq = (double)dq16[row];
        const double qs = 2.0 * s * sqrt(qn), rs = s * sqrt(rmax2);
        unscale = 1.0 / (s * s);
        eps += (qs * dr + dq * (rs + dr)) * unscale;
    }
    double tmin = INFINITY;   // the row's threshold: every reference outside the candidates has scan score >= tmin
    for (int s = 0; s < nsplit; ++s) tmin = fmin(tmin, (double)thr[row * nsplit + s]);
    tmin *= unscale;
    for (int e = lane; e < k; e += 32) {  // slots that no member claims (fewer than k ranked members) stay empty
        idx_out[n * k + e] = -1;
        dist_out[n * k + e] = INFINITY;
    }
    int bad = 0;      // a recorded scan score further than eps from the exact score voids the certificate
    int nsurv = 0;    // members kept for ranking (warp-uniform)
    const int total = ncand * gsz;
    const int gshift = gsz == 8 ? 3 : 0;
    for (int base = 0; base < total; base += 32) {
        const int e = base + lane;
        const bool in_range = e < total;
        const unsigned long long key = in_range ? cand[row * ncand + (e >> gshift)] : ~0ull;
        const bool gvalid = key != ~0ull;
        const int64_t pos = (int64_t)(uint32_t)(key & 0xffffffffull) + (e & (gsz - 1));   // position in the packed reference
        const bool valid = gvalid && pos < M;
        double s_exact = INFINITY;
        int32_t id = 0x7fffffff;
        if (valid) {
            id = order[pos];   // original reference row
            double rn = 0.0, dot = 0.0;
            if (gsz == 8) {
                const float *r = refT + pos;
                for (int c = 0; c < d; ++c) {
                    const double rv = (double)r[(int64_t)c * Mpad];
                    rn = fma(rv, rv, rn);
                    dot = fma((double)wq[c], rv, dot);
                }
            } else {
                const float *r = Ref + (int64_t)id * ldr;
                for (int c = 0; c < d; ++c) {
                    const double rv = (double)r[c];
                    rn = fma(rv, rv, rn);
                    dot = fma((double)wq[c], rv, dot);
                }
            }
            s_exact = rn - 2.0 * dot;
        }
        // the recorded score is the minimum over the group: compare it with the exact minimum over the members
        double gmin = s_exact;
        if (gsz == 8) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) gmin = fmin(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        }
        if (gvalid && (e & (gsz - 1)) == 0 && gmin != INFINITY)
            bad |= !(fabs((double)knn_key_score(key) * unscale - gmin) <= eps);
        // members that can still be among the k nearest IF the row certifies: exact score below the threshold
        // (with the error margin, so that exact ties at the threshold stay in).  The FP32 scan's candidates are
        // single references and all of them are ranked (it is the last tier: its answer stands).
        const bool keep = valid && (gsz == 1 || s_exact <= tmin + 2.0 * eps);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int off = nsurv + __popc(m & ((1u << lane) - 1u));
        if (keep && off < RR_MAXC) {
            wd[off] = fmax(qn + s_exact, 0.0);  // sklearn: ||x||^2 - 2 x.y + ||y||^2, clamped at 0
            wi[off] = id;
        }
        nsurv += __popc(m);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (nsurv > RR_MAXC) {   // cannot rank them all: rank what fits, flag the row
        nsurv = RR_MAXC;
        bad = 1;
    }
    __syncwarp();
    double kth = 0.0;
    for (int e = lane; e < nsurv; e += 32) {
        const double de = wd[e];
        const int32_t ie = wi[e];
        int rank = 0;
        for (int f = 0; f < nsurv; ++f) {
            const double df = wd[f];
            rank += (df < de) || (df == de && wi[f] < ie);
        }
        if (rank < k) {
            idx_out[n * k + rank] = ie;
            dist_out[n * k + rank] = (float)de;
        }
        if (rank == k - 1) kth = de;
    }
    if (unsafe != nullptr) {
        // certificate: every reference row NOT among the candidates has scan score >= tmin (its split's kc-th
        // best group minimum).  With |scan score - exact| <= eps it cannot beat the exact k-th neighbour if
        // kth_exact - ||q||^2 + 2 eps < tmin.
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kth = fmax(kth, __shfl_xor_sync(0xffffffffu, kth, o));
        bad |= nsurv < k;
        if (lane == 0) unsafe[n] = (!bad && kth - qn + 2.0 * eps < tmin) ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------ rerank of candidate GROUPS
// The tcgen05 scans hand over groups of 8 consecutive packed positions.  Giving all 8 x kc members of a row their
// float64 distance (the kernel above) costs a 32-byte sector per member group and dimension — 17 KB of gathers and
// 4300 float64 FMAs per row at config2, 6.7 ms of a 57 ms step (profiles/r2k).  Here the members are first scored the
// way the scan scored them, from the SAME packed operand tiles (one 128-byte line per group and k-unit: half the
// bytes, fully coalesced) in FP32; only the members that can still be among the k nearest if the row certifies —
// score below the row's threshold plus the error margin, ~kc of them — get the exact float64 distance from their
// reference row.  One warp per query row; lane <-> member in pass A, lane <-> survivor in pass B.
// a0 += lo(q) * lo(r), a1 += hi(q) * hi(r): FP32 accumulation of 16-bit operand pairs without converting them first
// (sm_100 mixed-precision FMA, `FHFMA`; the product of two 11-bit — or 8-bit — significands is exact in FP32, so this is
// bit for bit the FFMA of the converted values)
template <int PREC>
__device__ __forceinline__ void fma_pair16(float &a0, float &a1, uint32_t q, uint32_t r) {
    if (PREC == 1)
        asm("{\n .reg .b16 ql, qh, rl, rh;\n mov.b32 {ql, qh}, %2;\n mov.b32 {rl, rh}, %3;\n"
            " fma.rn.f32.f16 %0, ql, rl, %0;\n fma.rn.f32.f16 %1, qh, rh, %1;\n}\n"
            : "+f"(a0), "+f"(a1) : "r"(q), "r"(r));
    else
        asm("{\n .reg .b16 ql, qh, rl, rh;\n mov.b32 {ql, qh}, %2;\n mov.b32 {rl, rh}, %3;\n"
            " fma.rn.f32.bf16 %0, ql, rl, %0;\n fma.rn.f32.bf16 %1, qh, rh, %1;\n}\n"
            : "+f"(a0), "+f"(a1) : "r"(q), "r"(r));
}

template <int PREC>   // 0: BF16x3 operands, 1: single FP16
__global__ void __launch_bounds__(RR_THREADS)
knn_rerank_groups_kernel(const float *__restrict__ Q, int ldq, int64_t N, const float *__restrict__ Ref, int ldr, int64_t M,
                         int d, int k, const unsigned long long *__restrict__ cand, const float *__restrict__ thr, int ncand,
                         int nsplit, double eps_rel, const KnnHeader *__restrict__ hdr, const int32_t *__restrict__ order,
                         const int32_t *__restrict__ q_perm, const float *__restrict__ dq16,
                         const unsigned char *__restrict__ qrows, const unsigned char *__restrict__ rtiles, int kp,
                         int32_t *__restrict__ idx_out, float *__restrict__ dist_out, uint8_t *__restrict__ unsafe) {
    extern __shared__ __align__(16) unsigned char rr_smem[];
    constexpr int W = RR_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: survivors (double distance | int32 id, RG_MAXC each), the query row, its packed operand row
    double *wd = reinterpret_cast<double *>(rr_smem) + (size_t)warp * RG_MAXC;
    int32_t *wi = reinterpret_cast<int32_t *>(rr_smem + (size_t)W * RG_MAXC * sizeof(double)) + (size_t)warp * RG_MAXC;
    // the query row in float64 (pass B multiplies it with every survivor) and its packed 16-bit operand row (kp <= 256
    // values) as the scan saw it (pass A)
    double *wq = reinterpret_cast<double *>(rr_smem + (size_t)W * RG_MAXC * (sizeof(double) + sizeof(int32_t))) + (size_t)warp * RR_DMAX;
    uint4 *wo = reinterpret_cast<uint4 *>(rr_smem + (size_t)W * (RG_MAXC * (sizeof(double) + sizeof(int32_t)) + RR_DMAX * sizeof(double))) +
                (size_t)warp * 32;
    const int64_t row = (int64_t)blockIdx.x * W + warp;   // row of the scan (locality order)
    if (row >= N) return;
    const int64_t n = q_perm ? q_perm[row] : row;         // the query it stands for
    const float *q = Q + n * ldq;
    const int nunits = kp / 8;
    for (int c = lane; c < d; c += 32) wq[c] = (double)q[c];
    for (int j = lane; j < nunits; j += 32) wo[j] = __ldg(reinterpret_cast<const uint4 *>(qrows + (size_t)row * kp * 2) + j);
    __syncwarp();
    // ||q||^2 summed in the order pass B sums ||r||^2 and q.r: a query that IS a reference row then gets the distance
    // qn - 2 dot + rn = 0 exactly (every lane computes the same bits)
    double qn = 0.0;
    for (int c = 0; c < d; ++c) {
        const double v = wq[c];
        qn = fma(v, v, qn);
    }
    const double rmax2 = (double)hdr->rmax2;
    double eps = eps_rel * (2.0 * sqrt(qn) * sqrt(rmax2) + rmax2);
    double unscale = 1.0;   // scan scores and thresholds are in units of scale16^2 when the operands were single FP16
    if (PREC == 1) {        // (error bound of the single-FP16 scores: see knn_rerank_kernel)
        const double s = (double)hdr->scale16, dr = (double)hdr->dr16, dq = (double)dq16[row];
        const double qs = 2.0 * s * sqrt(qn), rs = s * sqrt(rmax2);
        unscale = 1.0 / (s * s);
        eps += (qs * dr + dq * (rs + dr)) * unscale;
    }
    double tmin = INFINITY;
    for (int s = 0; s < nsplit; ++s) tmin = fmin(tmin, (double)thr[row * nsplit + s]);
    tmin *= unscale;
    for (int e = lane; e < k; e += 32) {
        idx_out[n * k + e] = -1;
        dist_out[n * k + e] = INFINITY;
    }
    // ---- pass A: operand-precision scores of all members; survivors = members that may matter
    const size_t tile_bytes = (size_t)kp * 256;
    const float keep_below = (float)((tmin + 3.0 * eps) / unscale);   // in scan units (3 eps: the pass-A score is as far
                                                                      // from the exact one as the scan's was)
    int bad = 0, nsurv = 0;
    const int total = ncand * 8;
    float *ws = reinterpret_cast<float *>(wd);   // pass A parks (score | position) where pass B writes (distance | id)
    for (int base = 0; base < total; base += 32) {
        const int e = base + lane;
        const unsigned long long key = e < total ? cand[row * ncand + (e >> 3)] : ~0ull;
        const bool gvalid = key != ~0ull;
        const int64_t pos = (int64_t)(uint32_t)(key & 0xffffffffull) + (e & 7);
        const bool valid = gvalid && pos < M;
        float sc = INFINITY;
        // A group whose recorded minimum is already above the bar cannot hold a survivor (members score >= the group's
        // minimum): its tiles are not even read.  With one split every candidate group is below its split's threshold;
        // with many (few query rows, e.g. the repair passes: up to 28 splits x kc groups per row) almost all are skipped.
        const bool wanted = gvalid && knn_key_score(key) <= keep_below;
        if (wanted) {   // (padding positions of the last tile hold +inf norms: their score is +inf)
            const unsigned char *t = rtiles + (size_t)(pos >> 7) * tile_bytes + ((pos & 127) >> 3) * 128 + (pos & 7) * 16;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
            for (int j = 0; j < nunits; ++j) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4 *>(t + (size_t)j * 2048));
                const uint4 qv = wo[j];
                fma_pair16<PREC>(a0, a1, qv.x, rv.x);
                fma_pair16<PREC>(a0, a1, qv.y, rv.y);
                fma_pair16<PREC>(a0, a1, qv.z, rv.z);
                fma_pair16<PREC>(a0, a1, qv.w, rv.w);
            }
            sc = a0 + a1;
        }
        // the recorded group score is the minimum over the group as the tensor core computed it: the same operands
        // and products, another summation order
        float gmin = sc;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        if (wanted && (e & 7) == 0) {
            const float rec = knn_key_score(key);
            bad |= !(fabs((double)rec - (double)gmin) * unscale <= eps);
        }
        const bool keep = valid && sc <= keep_below;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int off = nsurv + __popc(m & ((1u << lane) - 1u));
        if (keep && off < RG_MAXC) {
            ws[2 * off] = sc;
            wi[off] = (int32_t)pos;
        }
        nsurv += __popc(m);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (nsurv > RG_MAXC) {   // cannot rank them all: rank what fits, flag the row
        nsurv = RG_MAXC;
        bad = 1;
    }
    __syncwarp();
    // ---- pass B: exact float64 distance of the survivors (sklearn's formula), from their reference rows
    const bool ref_pairs = (ldr & 1) == 0 && (reinterpret_cast<uintptr_t>(Ref) & 7) == 0;
    int nkeep = 0;
    for (int base = 0; base < nsurv; base += 32) {
        const int e = base + lane;
        double s_exact = INFINITY;
        int32_t id = 0x7fffffff;
        float sA = 0.f;
        if (e < nsurv) {
            sA = ws[2 * e];
            id = order[wi[e]];
            const float *r = Ref + (int64_t)id * ldr;
            double rn = 0.0, dot = 0.0;
            int c = 0;
            if (ref_pairs) {   // rows start on 8-byte boundaries: two components per load
                for (; c + 1 < d; c += 2) {
                    const float2 rv2 = __ldg(reinterpret_cast<const float2 *>(r + c));
                    const double2 qv = *reinterpret_cast<const double2 *>(wq + c);
                    const double r0 = (double)rv2.x, r1 = (double)rv2.y;
                    rn = fma(r0, r0, rn);
                    dot = fma(qv.x, r0, dot);
                    rn = fma(r1, r1, rn);
                    dot = fma(qv.y, r1, dot);
                }
            }
            for (; c < d; ++c) {
                const double rv = (double)__ldg(r + c);
                rn = fma(rv, rv, rn);
                dot = fma(wq[c], rv, dot);
            }
            s_exact = rn - 2.0 * dot;
            bad |= !(fabs((double)sA * unscale - s_exact) <= eps);   // the error bound holds on the members that matter
        }
        __syncwarp();   // all lanes have read their (score | position) slots before they are overwritten
        const bool keep = e < nsurv && s_exact <= tmin + 2.0 * eps;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int off = nkeep + __popc(m & ((1u << lane) - 1u));
        if (keep) {   // off <= e: slots below the read front
            wd[off] = fmax(qn + s_exact, 0.0);  // sklearn: ||x||^2 - 2 x.y + ||y||^2, clamped at 0
            wi[off] = id;
        }
        nkeep += __popc(m);
        __syncwarp();
    }
    bad = __any_sync(0xffffffffu, bad);
    double kth = 0.0;
    for (int e = lane; e < nkeep; e += 32) {
        const double de = wd[e];
        const int32_t ie = wi[e];
        int rank = 0;
        for (int f = 0; f < nkeep; ++f) {
            const double df = wd[f];
            rank += (df < de) || (df == de && wi[f] < ie);
        }
        if (rank < k) {
            idx_out[n * k + rank] = ie;
            dist_out[n * k + rank] = (float)de;
        }
        if (rank == k - 1) kth = de;
    }
    if (unsafe != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kth = fmax(kth, __shfl_xor_sync(0xffffffffu, kth, o));
        bad |= nkeep < k;
        if (lane == 0) unsafe[n] = (!bad && kth - qn + 2.0 * eps < tmin) ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------ exact search of a few rows
// The handful of rows per pass whose first-tier result could not be certified used to go through a second scan and
// re-rank launched for them alone: two small-grid launches that spend ~0.8 ms on 4 rows.  For a few rows brute force
// in float64 is cheaper than any scan — 2 R M d flop at R <= 64 is a reading of the reference — and needs no
// certificate, given an upper bound on each row's k-th distance: the first tier's own k-th result is one (they are exact
// distances of real reference rows).  Kernel 1 lists, per row, every reference row within the bound (normally k plus
// a few); kernel 2 ranks the list by (distance, index) and writes the first k.
constexpr int XR_MAX_ROWS = 64;
constexpr int XR_CAP = 1024;     // listed reference rows per query row; more (a useless bound) -> the row is flagged
constexpr int XR_THREADS = 256;
constexpr int XR_RJ = 4;         // query rows per pass over a reference row

__global__ void __launch_bounds__(XR_THREADS)
knn_exact_collect_kernel(const float *__restrict__ Q, int ldq, const int64_t *__restrict__ rows, int n_rows,
                         const float *__restrict__ Ref, int ldr, int64_t M, int d, const float *__restrict__ ub,
                         int32_t *__restrict__ counts, double *__restrict__ ldist, int32_t *__restrict__ lid) {
    extern __shared__ __align__(16) unsigned char xr_smem[];
    double *qd = reinterpret_cast<double *>(xr_smem);          // [n_rows][d]
    double *qn = qd + (size_t)n_rows * d;                      // [n_rows]
    double *bound = qn + n_rows;                               // [n_rows]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < n_rows; j += XR_THREADS / 32) {
        const float *q = Q + rows[j] * ldq;
        for (int c = lane; c < d; c += 32) qd[(size_t)j * d + c] = (double)q[c];
        __syncwarp();
        if (lane == 0) {   // ||q||^2 in the order ||r||^2 and q.r are summed below (a row's distance to itself is exactly 0)
            double s = 0.0;
            for (int c = 0; c < d; ++c) s = fma(qd[(size_t)j * d + c], qd[(size_t)j * d + c], s);
            qn[j] = s;
            bound[j] = (double)ub[j];
        }
    }
    // 256 reference rows at a time are staged in shared memory with coalesced loads (row stride odd: thread t then
    // reads row t without bank conflicts; straight from global memory every load of a warp touched 32 lines)
    float *rs = reinterpret_cast<float *>(bound + n_rows);     // [XR_THREADS][DS]
    const int DS = d | 1;
    for (int64_t m0 = (int64_t)blockIdx.x * XR_THREADS; m0 < M; m0 += (int64_t)gridDim.x * XR_THREADS) {
        __syncthreads();
        const int nr = M - m0 < XR_THREADS ? (int)(M - m0) : XR_THREADS;
        if (ldr == d) {
            const float *src = Ref + m0 * ldr;
            for (int i = threadIdx.x; i < nr * d; i += XR_THREADS) rs[(i / d) * DS + (i % d)] = __ldg(src + i);
        } else {
            for (int i = threadIdx.x; i < nr * d; i += XR_THREADS) rs[(i / d) * DS + (i % d)] = __ldg(Ref + (m0 + i / d) * ldr + (i % d));
        }
        __syncthreads();
        if ((int)threadIdx.x >= nr) continue;
        const int64_t m = m0 + threadIdx.x;
        const float *r = rs + threadIdx.x * DS;
        double rn = 0.0;
        for (int c = 0; c < d; ++c) {
            const double rv = (double)r[c];
            rn = fma(rv, rv, rn);
        }
        for (int j0 = 0; j0 < n_rows; j0 += XR_RJ) {
            double dot[XR_RJ];
#pragma unroll
            for (int j = 0; j < XR_RJ; ++j) dot[j] = 0.0;
            for (int c = 0; c < d; ++c) {
                const double rv = (double)r[c];
#pragma unroll
                for (int j = 0; j < XR_RJ; ++j)
                    if (j0 + j < n_rows) dot[j] = fma(qd[(size_t)(j0 + j) * d + c], rv, dot[j]);
            }
#pragma unroll
            for (int j = 0; j < XR_RJ; ++j) {
                if (j0 + j >= n_rows) break;
                const double dist = fmax(qn[j0 + j] + (rn - 2.0 * dot[j]), 0.0);   // sklearn's formula, as in the re-rank
                if (dist <= bound[j0 + j]) {
                    const int pos = atomicAdd(&counts[j0 + j], 1);
                    if (pos < XR_CAP) {
                        ldist[(size_t)(j0 + j) * XR_CAP + pos] = dist;
                        lid[(size_t)(j0 + j) * XR_CAP + pos] = (int32_t)m;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(XR_THREADS)
knn_exact_select_kernel(int k, const int32_t *__restrict__ counts, const double *__restrict__ ldist,
                        const int32_t *__restrict__ lid, int32_t *__restrict__ idx_out, float *__restrict__ dist_out,
                        uint8_t *__restrict__ overflow) {
    __shared__ double sd[XR_CAP];
    __shared__ int32_t si[XR_CAP];
    const int j = blockIdx.x, n = counts[j];
    if (n > XR_CAP || n < k) {   // the bound was useless (or not a bound at all): the caller takes the scan tiers
        if (threadIdx.x == 0) overflow[j] = 1;
        for (int e = threadIdx.x; e < k; e += XR_THREADS) {
            idx_out[(size_t)j * k + e] = -1;
            dist_out[(size_t)j * k + e] = INFINITY;
        }
        return;
    }
    if (threadIdx.x == 0) overflow[j] = 0;
    for (int e = threadIdx.x; e < n; e += XR_THREADS) {
        sd[e] = ldist[(size_t)j * XR_CAP + e];
        si[e] = lid[(size_t)j * XR_CAP + e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += XR_THREADS) {
        const double de = sd[e];
        const int32_t ie = si[e];
        int rank = 0;
        for (int f = 0; f < n; ++f) rank += (sd[f] < de) || (sd[f] == de && si[f] < ie);
        if (rank < k) {
            idx_out[(size_t)j * k + rank] = ie;
            dist_out[(size_t)j * k + rank] = (float)de;
        }
    }
}

// ------------------------------------------------------------------------------------------ nearest centroid
// code[n] = argmin_c ||x_n - cent_c||^2 (lowest c on ties).  Used to put queries and reference rows into the
// same coarse spatial order, so that a query tile meets its likely neighbours in the first reference tiles
// it scans and its thresholds are tight for the rest of the scan.  128 rows per block, rows and a chunk of
// centroids in shared memory.
constexpr int NC_ROWS = 128;
constexpr int NC_CHUNK = 64;
// One thread per row with the row in registers (DPAD = d rounded up to 32 / 64 / 128); a chunk of centroids, zero
// padded to DPAD and followed by ||c||^2, is staged in shared memory and read as broadcast float4s: one LDS.128
// per four FMAs (the first version kept the rows in shared memory too: two loads per FMA, 7.7 TF/s).
template <int DPAD>
__global__ void __launch_bounds__(NC_ROWS)
nearest_centroid_kernel(const float *__restrict__ X, int ldx, int64_t N, int d, const float *__restrict__ cent, int C,
                        int32_t *__restrict__ code) {
    constexpr int CS = DPAD + 4;          // row stride of the staged centroids (16-byte aligned rows)
    __shared__ __align__(16) float cs[NC_CHUNK * CS];
    const int tid = threadIdx.x;
    const int64_t n = (int64_t)blockIdx.x * NC_ROWS + tid;
    float x[DPAD];
#pragma unroll
    for (int j = 0; j < DPAD; ++j) x[j] = (n < N && j < d) ? X[n * ldx + j] : 0.f;
    float best = INFINITY;
    int bestc = 0;
    for (int c0 = 0; c0 < C; c0 += NC_CHUNK) {
        __syncthreads();
        const int nc = min(NC_CHUNK, C - c0);
        for (int i = tid; i < nc * DPAD; i += NC_ROWS) {
            const int c = i / DPAD, j = i % DPAD;
            cs[c * CS + j] = j < d ? cent[(int64_t)(c0 + c) * d + j] : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < nc; i += NC_ROWS) {
            float nn = 0.f;
            for (int j = 0; j < d; ++j) nn = fmaf(cs[i * CS + j], cs[i * CS + j], nn);
            cs[i * CS + DPAD] = nn;
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            const float4 *cr = reinterpret_cast<const float4 *>(cs + c * CS);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int j = 0; j < DPAD / 4; ++j) {
                const float4 v = cr[j];
                d0 = fmaf(x[4 * j], v.x, d0);
                d1 = fmaf(x[4 * j + 1], v.y, d1);
                d0 = fmaf(x[4 * j + 2], v.z, d0);
                d1 = fmaf(x[4 * j + 3], v.w, d1);
            }
            const float s = cs[c * CS + DPAD] - 2.f * (d0 + d1);  // ||x||^2 is common to all centroids
            if (s < best) { best = s; bestc = c0 + c; }
        }
    }
    if (n < N) code[n] = bestc;
}

// ------------------------------------------------------------------------------------------ vote
// One thread per query row: label = most frequent code among the k neighbours, ties to the lowest
// code (sklearn: classes_ sorted, argmax takes the first maximum); conf = votes / k.
constexpr int VT_KMAX = 128;
__global__ void __launch_bounds__(128)
knn_vote_kernel(const int32_t *__restrict__ idx, const float *__restrict__ dist2, int64_t N, int k,
                const int32_t *__restrict__ ref_codes, int64_t M, int32_t *__restrict__ label, float *__restrict__ conf) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int32_t code[VT_KMAX];
    for (int j = 0; j < k; ++j) {
        const int32_t id = idx[n * k + j];
        code[j] = (id >= 0 && id < M) ? ref_codes[id] : -1;
    }
    if (dist2 == nullptr) {  // weights="uniform"
        int best = -1, bestc = 0;
        for (int j = 0; j < k; ++j) {
            const int32_t cj = code[j];
            if (cj < 0) continue;
            int cnt = 0;
            for (int l = 0; l < k; ++l) cnt += code[l] == cj;
            if (cnt > bestc || (cnt == bestc && cj < best)) { bestc = cnt; best = cj; }
        }
        label[n] = best;
        if (conf) conf[n] = (float)bestc / (float)k;
        return;
    }
    // weights="distance" (sklearn _get_weights): w = 1/dist; a row with zero distances gives weight 1 to those
    // neighbours and 0 to the others.  Class with the largest total weight, ties to the lowest class code.
    bool any_zero = false;
    for (int j = 0; j < k; ++j) any_zero |= (code[j] >= 0 && dist2[n * k + j] == 0.f);
    double bestw = -1.0, total = 0.0;
    int best = -1;
    for (int j = 0; j < k; ++j) {
        const int32_t cj = code[j];
        if (cj < 0) continue;
        double w = 0.0;
        for (int l = 0; l < k; ++l) {
            if (code[l] != cj) continue;
            const float d2 = dist2[n * k + l];
            w += any_zero ? (d2 == 0.f ? 1.0 : 0.0) : 1.0 / sqrt((double)d2);
        }
        if (w > bestw || (w == bestw && cj < best)) { bestw = w; best = cj; }
    }
    for (int l = 0; l < k; ++l) {
        if (code[l] < 0) continue;
        const float d2 = dist2[n * k + l];
        total += any_zero ? (d2 == 0.f ? 1.0 : 0.0) : 1.0 / sqrt((double)d2);
    }
    label[n] = best;
    if (conf) conf[n] = total > 0.0 ? (float)(bestw / total) : 0.f;
}

}  // namespace sym

// ============================================================================================
using namespace sym;

extern "C" size_t sym_knn_packed_bytes(int64_t M, int32_t d) { return knn_packed_layout(nullptr, M, d).total; }

extern "C" int sym_nearest_centroid(const float *X, int32_t ldx, int64_t N, int32_t d, const float *centroids, int32_t C,
                                    int32_t *code, sym_stream_t stream) {
    SYM_REQUIRE(N >= 0 && d > 0 && d <= 128 && C > 0 && ldx >= d, SYM_ERR_INVALID_ARGUMENT, "sym_nearest_centroid: bad sizes");
    if (N == 0) return SYM_OK;
    const unsigned grid = (unsigned)ceil_div(N, NC_ROWS);
    cudaStream_t st = (cudaStream_t)stream;
    if (d <= 32) nearest_centroid_kernel<32><<<grid, NC_ROWS, 0, st>>>(X, ldx, N, d, centroids, C, code);
    else if (d <= 64) nearest_centroid_kernel<64><<<grid, NC_ROWS, 0, st>>>(X, ldx, N, d, centroids, C, code);
    else nearest_centroid_kernel<128><<<grid, NC_ROWS, 0, st>>>(X, ldx, N, d, centroids, C, code);
    SYM_LAUNCH_CHECK("nearest_centroid_kernel");
    return SYM_OK;
}

extern "C" int sym_knn_fit(const float *Ref, int32_t ldr, int64_t M, int32_t d, const int32_t *order, void *packed,
                           sym_stream_t stream) {
    SYM_REQUIRE(M > 0 && d > 0 && ldr >= d, SYM_ERR_INVALID_ARGUMENT, "sym_knn_fit: bad sizes");
    SYM_REQUIRE(d <= 128, SYM_ERR_UNSUPPORTED_SHAPE, "sym_knn_fit: d=%d > 128", d);
    SYM_REQUIRE(M < (1ll << 31) - 256, SYM_ERR_UNSUPPORTED_SHAPE, "sym_knn_fit: M too large");
    SYM_REQUIRE(((uintptr_t)packed & 255) == 0, SYM_ERR_INVALID_ARGUMENT, "sym_knn_fit: packed must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    KnnPacked pk = knn_packed_layout(packed, M, d);
    SYM_CUDA(cudaMemsetAsync(pk.hdr, 0, sizeof(KnnHeader), st));
    knn_fit_kernel<<<(unsigned)(pk.Mpad / 128), 256, 0, st>>>(Ref, ldr, M, d, order, pk);
    SYM_LAUNCH_CHECK("knn_fit_kernel");
    return knn_tc_fit(Ref, ldr, M, d, order, pk, st);
}

namespace {
struct ScanPlan {
    int kc, kcap, nsplit, tiles_per_split, top_in_smem;
    int64_t qtiles;
    size_t smem, cand_bytes, thr_bytes, top_bytes;
};

int plan_scan(int64_t N, int64_t M, int d, int k, ScanPlan &p) {
    const int64_t Mpad = knn_mpad(M);
    const int d4 = (d + 3) & ~3;
    int kc = k + 8;
    if (kc > M) kc = (int)M;
    p.kc = kc;
    p.kcap = kc <= 16 ? 16 : kc <= 32 ? 32 : kc <= 64 ? 64 : 128;
    p.top_in_smem = p.kcap <= 64;
    p.smem = scan_smem_bytes(d4, p.kcap, p.top_in_smem);
    p.qtiles = ceil_div(N, SC_BQ);
    const int ntiles = (int)(Mpad / SC_BR);
    // split the reference across CTAs only when the query alone cannot fill the GPU
    int nsplit = 1;
    if (p.qtiles < 2 * kNumSMs) {
        nsplit = (int)ceil_div(2 * kNumSMs, p.qtiles);
        const int by_work = ntiles / 8 > 0 ? ntiles / 8 : 1;   // keep >= 8 tiles per split
        const int by_rerank = RR_MAXC / kc > 0 ? RR_MAXC / kc : 1;
        nsplit = nsplit < by_work ? nsplit : by_work;
        nsplit = nsplit < by_rerank ? nsplit : by_rerank;
        if (nsplit < 1) nsplit = 1;
    }
    p.tiles_per_split = (int)ceil_div(ntiles, nsplit);
    p.nsplit = (int)ceil_div(ntiles, p.tiles_per_split);
    p.cand_bytes = ((size_t)N * p.nsplit * kc * sizeof(unsigned long long) + 255) & ~(size_t)255;
    p.thr_bytes = ((size_t)N * p.nsplit * sizeof(float) + 255) & ~(size_t)255;
    p.top_bytes = p.top_in_smem ? 0 : (size_t)p.qtiles * p.nsplit * SC_BQ * p.kcap * 8;
    return SYM_OK;
}
}  // namespace

extern "C" size_t sym_knn_workspace(int64_t N, int64_t M, int32_t d, int32_t k, int32_t method) {
    if (N <= 0 || M <= 0 || d <= 0 || k <= 0) return 256;
    ScanPlan p;
    plan_scan(N, M, d, k, p);
    size_t ffma = p.cand_bytes + p.thr_bytes + p.top_bytes + 256;
    const size_t tc0 = knn_tc_workspace(N, M, d, k, 0), tc1 = knn_tc_workspace(N, M, d, k, 1);
    const size_t tc = tc0 > tc1 ? tc0 : tc1;
    (void)method;
    return ffma > tc ? ffma : tc;
}

extern "C" int sym_knn_method_supported(int64_t N, int64_t M, int32_t d, int32_t k, int32_t method) {
    if (N < 0 || M <= 0 || d <= 0 || k < 1 || k > M || k > 120 || d > 128) return 0;
    if (method == 0 || method == 1) return 1;
    if (method == 2 || method == 3) return sym::knn_tc_supported(N > 0 ? N : 1, M, d, k, method == 3 ? 1 : 0) ? 1 : 0;
    return 0;
}

extern "C" int sym_knn(const float *Q, int32_t ldq, int64_t N, const float *Ref, int32_t ldr, int64_t M,
                       const void *packed, int32_t d, int32_t k, int32_t method, const int32_t *q_perm,
                       const int32_t *q_code, const int32_t *cl_tile, int32_t *idx, float *dist2, uint8_t *unsafe,
                       void *workspace, size_t workspace_bytes, sym_stream_t stream) {
    SYM_REQUIRE(N >= 0 && M > 0 && d > 0 && ldq >= d && ldr >= d, SYM_ERR_INVALID_ARGUMENT, "sym_knn: bad sizes");
    SYM_REQUIRE(k >= 1 && k <= M, SYM_ERR_INVALID_ARGUMENT,
                "sym_knn: Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %lld", k,
                (long long)M);
    SYM_REQUIRE(k <= 120 && d <= 128, SYM_ERR_UNSUPPORTED_SHAPE, "sym_knn: k=%d (<=120) d=%d (<=128) unsupported", k, d);
    const int prune = (method & SYM_KNN_PRUNE) ? 1 : 0;   // opt-in exact tile pruning (tcgen05 scans)
    method &= ~SYM_KNN_PRUNE;
    SYM_REQUIRE(method >= 0 && method <= 3, SYM_ERR_INVALID_ARGUMENT, "sym_knn: unknown method %d", method);
    SYM_REQUIRE(workspace_bytes >= sym_knn_workspace(N, M, d, k, method), SYM_ERR_WORKSPACE_TOO_SMALL,
                "sym_knn: workspace too small");
    SYM_REQUIRE(((uintptr_t)workspace & 255) == 0, SYM_ERR_INVALID_ARGUMENT, "sym_knn: workspace must be 256-byte aligned");
    if (N == 0) return SYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    KnnPacked pk = knn_packed_layout(const_cast<void *>(packed), M, d);
    if (method == 0) method = knn_tc_supported(N, M, d, k, 0) ? 2 : 1;

    int ncand = 0, nsplit = 1, group = 1;   // group: reference positions per candidate (8 for the tcgen05 scans)
    double eps_rel = 0.0;
    const unsigned long long *cand = nullptr;
    const float *thr = nullptr;
    const float *dq16 = nullptr;
    const unsigned char *qrows = nullptr;   // tcgen05 scans: the packed operand rows of the queries (scan order) and their width
    int kp = 0;
    if (method >= 2) {
        const int prec = method == 3 ? 1 : 0;
        SYM_REQUIRE(knn_tc_supported(N, M, d, k, prec), SYM_ERR_UNSUPPORTED_SHAPE,
                    "sym_knn: tcgen05 scan does not support this shape");
        int rc = knn_tc_scan(Q, ldq, N, q_perm, q_code, cl_tile, pk, M, d, k, prec, prune, workspace, workspace_bytes, st, &cand,
                             &thr, &ncand, &nsplit, &group, &eps_rel, &dq16, &qrows, &kp);
        if (rc != SYM_OK) return rc;
    } else {
        ScanPlan p;
        plan_scan(N, M, d, k, p);
        unsigned char *w = (unsigned char *)workspace;
        unsigned long long *cand_w = (unsigned long long *)w; w += p.cand_bytes;
        float *thr_w = (float *)w; w += p.thr_bytes;
        unsigned long long *top_w = (unsigned long long *)w;
        SYM_REQUIRE(p.smem <= 227 * 1024, SYM_ERR_UNSUPPORTED_SHAPE, "sym_knn: d=%d k=%d needs %zu B of shared memory", d, k, p.smem);
        SYM_CUDA(cudaFuncSetAttribute(knn_scan_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        dim3 grid((unsigned)p.qtiles, (unsigned)p.nsplit);
        knn_scan_ffma_kernel<<<grid, SC_THREADS, p.smem, st>>>(Q, ldq, N, q_perm, M, pk, d, p.kc, p.kcap, p.tiles_per_split,
                                                               p.top_in_smem, top_w, cand_w, thr_w);
        SYM_LAUNCH_CHECK("knn_scan_ffma_kernel");
        cand = cand_w; thr = thr_w; ncand = p.nsplit * p.kc; nsplit = p.nsplit;
        // FP32 dot product of d terms, accumulated by fmaf, plus the rounding of ||r||^2 and the sum
        eps_rel = (double)(d + 4) * 5.97e-8;
    }
    SYM_REQUIRE(ncand <= RR_MAXC, SYM_ERR_UNSUPPORTED_SHAPE, "sym_knn: internal: %d candidates per row", ncand);
    if (group == 8) {   // tcgen05 scans: candidate groups, scored from the packed tiles first
        const unsigned grid = (unsigned)ceil_div(N, RR_THREADS / 32);
        if (method == 3) {
            SYM_CUDA(cudaFuncSetAttribute(knn_rerank_groups_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM_GROUPS));
            knn_rerank_groups_kernel<1><<<grid, RR_THREADS, RR_SMEM_GROUPS, st>>>(
                Q, ldq, N, Ref, ldr, M, d, k, cand, thr, ncand, nsplit, eps_rel, pk.hdr, pk.order, q_perm, dq16, qrows, pk.tc16, kp,
                idx, dist2, unsafe);
        } else {
            SYM_CUDA(cudaFuncSetAttribute(knn_rerank_groups_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM_GROUPS));
            knn_rerank_groups_kernel<0><<<grid, RR_THREADS, RR_SMEM_GROUPS, st>>>(
                Q, ldq, N, Ref, ldr, M, d, k, cand, thr, ncand, nsplit, eps_rel, pk.hdr, pk.order, q_perm, dq16, qrows, pk.tc, kp,
                idx, dist2, unsafe);
        }
        SYM_LAUNCH_CHECK("knn_rerank_groups_kernel");
        return SYM_OK;
    }
    SYM_CUDA(cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM));
    knn_rerank_kernel<<<(unsigned)ceil_div(N, RR_THREADS / 32), RR_THREADS, RR_SMEM, st>>>(
        Q, ldq, N, Ref, ldr, pk.refT, pk.Mpad, M, d, k, cand, thr, ncand, nsplit, group, eps_rel, pk.hdr, pk.order, q_perm,
        dq16, idx, dist2, unsafe);
    SYM_LAUNCH_CHECK("knn_rerank_kernel");
    return SYM_OK;
}

extern "C" int sym_vote(const int32_t *idx, const float *dist2, int64_t N, int32_t k, const int32_t *ref_codes, int64_t M,
                        int32_t *label, float *conf, sym_stream_t stream) {
    SYM_REQUIRE(N >= 0 && k >= 1 && k <= VT_KMAX, SYM_ERR_INVALID_ARGUMENT, "sym_vote: bad sizes (k <= %d)", VT_KMAX);
    if (N == 0) return SYM_OK;
    knn_vote_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(idx, dist2, N, k, ref_codes, M, label,
                                                                                    conf);
    SYM_LAUNCH_CHECK("knn_vote_kernel");
    return SYM_OK;
}

// ---- exact search of a few rows (see knn_exact_collect_kernel)
extern "C" int32_t sym_knn_exact_rows_max(void) { return sym::XR_MAX_ROWS; }
extern "C" size_t sym_knn_exact_rows_workspace(int32_t n_rows) {
    const size_t n = n_rows > 0 ? (size_t)n_rows : 0;
    return 256 + ((n * sizeof(int32_t) + 255) & ~(size_t)255) + n * sym::XR_CAP * (sizeof(double) + sizeof(int32_t));
}
extern "C" int sym_knn_exact_rows(const float *Q, int32_t ldq, const int64_t *rows, int32_t n_rows, const float *Ref,
                                  int32_t ldr, int64_t M, int32_t d, int32_t k, const float *ub, int32_t *idx_out,
                                  float *dist2_out, uint8_t *overflow, void *workspace, size_t workspace_bytes,
                                  sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(n_rows >= 0 && n_rows <= XR_MAX_ROWS && M >= 1 && d >= 1 && d <= 128 && k >= 1 && k <= M && ldq >= d && ldr >= d,
                SYM_ERR_INVALID_ARGUMENT, "sym_knn_exact_rows: bad sizes (n_rows <= %d, d <= 128, k <= M)", XR_MAX_ROWS);
    SYM_REQUIRE(k <= XR_CAP, SYM_ERR_UNSUPPORTED_SHAPE, "sym_knn_exact_rows: k=%d (<= %d)", k, XR_CAP);
    if (n_rows == 0) return SYM_OK;
    SYM_REQUIRE(workspace_bytes >= sym_knn_exact_rows_workspace(n_rows), SYM_ERR_WORKSPACE_TOO_SMALL,
                "sym_knn_exact_rows: workspace of %zu B, %zu needed", workspace_bytes, sym_knn_exact_rows_workspace(n_rows));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *w = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    int32_t *counts = reinterpret_cast<int32_t *>(w);
    double *ldist = reinterpret_cast<double *>(w + (((size_t)n_rows * sizeof(int32_t) + 255) & ~(size_t)255));
    int32_t *lid = reinterpret_cast<int32_t *>(ldist + (size_t)n_rows * XR_CAP);
    SYM_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_rows * sizeof(int32_t), st));
    const size_t smem = ((size_t)n_rows * d + 2 * (size_t)n_rows) * sizeof(double) + (size_t)XR_THREADS * (d | 1) * sizeof(float);
    SYM_CUDA(cudaFuncSetAttribute(knn_exact_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = ceil_div(M, XR_THREADS);
    const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
    if (grid > (int64_t)per_sm * kNumSMs) grid = (int64_t)per_sm * kNumSMs;
    knn_exact_collect_kernel<<<(unsigned)grid, XR_THREADS, smem, st>>>(Q, ldq, rows, n_rows, Ref, ldr, M, d, ub, counts, ldist, lid);
    SYM_LAUNCH_CHECK("knn_exact_collect_kernel");
    knn_exact_select_kernel<<<(unsigned)n_rows, XR_THREADS, 0, st>>>(k, counts, ldist, lid, idx_out, dist2_out, overflow);
    SYM_LAUNCH_CHECK("knn_exact_select_kernel");
    return SYM_OK;
}
