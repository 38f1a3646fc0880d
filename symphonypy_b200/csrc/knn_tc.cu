// tcgen05 candidate scan for the brute-force kNN (methods 2 and 3)  — symphonypy/tl.py:425-433
//
// The distance GEMM  S[q,r] = ||r||^2 - 2 q.r  runs on the 5th-generation tensor cores, with the norm carried as
// extra K columns so that ONE chain of tcgen05.mma instructions per tile yields the finished score.  Two operand
// precisions, both packed once per reference by sym_knn_fit:
//
//   method 3, single FP16 (first tier)        K' = d + 2
//     K' index      query operand (A, M=128 rows)      reference operand (B, N=128 rows)
//     [0,d)         fp16(-2 s q)                       fp16(s r)         s = power of two, max ||s r|| in (8, 16]
//     d, d+1        1                                  the two fp16 pieces of ||s r||^2
//   method 2, BF16x3 (second tier)             K' = 3d + 3   (each float is hi + lo bf16; hi.hi + hi.lo + lo.hi)
//     [0,d)         hi(-2q)                            hi(r)
//     [d,2d)        hi(-2q)                            lo(r)
//     [2d,3d)       lo(-2q)                            hi(r)
//     3d..3d+2      1                                  the three bf16 pieces of ||r||^2 (exact)
//
// Reference tiles are stored in the canonical K-major no-swizzle UMMA layout, so a tile is one contiguous block
// moved by a single cp.async.bulk into shared memory; the accumulators (2-4 x 128x128 FP32) never leave tensor
// memory until the selection warps read them.
//
// Warp roles (384 threads for two query tiles):  warp 0 = bulk-copy producer; warps 1-2 = MMA issuers, one per
// query tile (warp 1 owns the TMEM allocation); warps 4-11 = selection, thread <-> query row <-> TMEM lane.  A
// selection thread pulls the 128 scores of its row with four tcgen05.ld, hands the accumulator back and folds the
// scores with 3-input min into the minima of 16 GROUPS of 8 references, compared with its row's threshold in one
// straight-line block.  What is selected are groups (minimum score, first position).  On a hit the thread only PARKS
// the tile (its 16 group minima + first position, five stores) in its row's small FIFO; at the same tiles for every
// warp of the CTA each row then moves one parked tile into its 4-ary max-heap of the kc best groups in shared
// memory (root = threshold), lanes in lock-step — so the four warps that share an accumulator are slow together
// instead of one after the other.
// The scan over-selects (kc = k + 8 groups, k + 16 for k >= 40); the re-rank in knn.cu expands the groups, certifies
// the margin with the scores recorded here (error bound per row: measured FP16 rounding residuals, or 6e-5 relative
// for BF16x3) and flags the rows that have to be re-run with the next tier.  Opt-in: exact tile pruning (bounding
// balls of the reference tiles against the rows' published thresholds; tile numbers through a ring in shared memory).
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "knn_common.cuh"
#include "tc_ptx.cuh"

namespace sym {

constexpr int TC_BM = 128;        // query rows per CTA (UMMA M)
constexpr int TC_BN = 128;        // reference rows per tile (UMMA N)
// query tiles per CTA: QT = 2 (256 rows share every reference tile) when shared memory allows, else 1;
// TMEM accumulator buffers per query tile: 4 / QT  (QT * ACC * 128 = 512 columns)
constexpr int TC_MAX_STAGES = 16;   // reference tiles in flight (a bulk copy takes 1-2 us from L2: ~2000+ cycles of prefetch distance)
constexpr int TC_PAD = 8;         // over-selection: kc = k + TC_PAD

// operand precision: 0 = BF16x3 (three split products, K' = 3d+3), 1 = single FP16 (K' = d+2)
__host__ __device__ inline int tc_kp(int d, int prec) { return prec ? (d + 2 + 15) / 16 * 16 : (3 * d + 3 + 15) / 16 * 16; }
__host__ __device__ inline size_t tc_tile_bytes(int kp) { return (size_t)kp * 256; }  // 128 rows x kp bf16

// ------------------------------------------------------------------------------------------ packing
__device__ __forceinline__ unsigned short bf16_rn(float x) {
    unsigned u = __float_as_uint(x);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
__device__ __forceinline__ float bf16_to_f(unsigned short h) { return __uint_as_float((unsigned)h << 16); }
__device__ __forceinline__ unsigned short f16_rn(float x) { return __half_as_ushort(__float2half_rn(x)); }
__device__ __forceinline__ float f16_to_f(unsigned short h) { return __half2float(__ushort_as_half(h)); }
// |v - fp16(v)| as the error bound of the re-rank counts it.  Values below the smallest normal fp16 are charged in
// full: the bound then holds whether or not the tensor core honours fp16 subnormal inputs (nothing here depends on it).
__device__ __forceinline__ float tc_f16_residual(float v, unsigned short h) {
    return fabsf(v) < 6.103515625e-5f ? v : v - f16_to_f(h);
}

// One thread = one row x one k-unit (8 consecutive K' indices, 16 bytes).
// Element (r, kk) of a tile lives at  (kk/8)*2048 + (r/8)*128 + (r%8)*16 + (kk%8)*2  bytes.
__global__ void __launch_bounds__(256)
tc_pack_kernel(const float *__restrict__ src, int ld, int64_t rows, int64_t rows_pad, int d, int kp, int is_query,
               const int32_t *__restrict__ gather, unsigned char *__restrict__ dst, int prec,
               KnnHeader *__restrict__ hdr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;  // k-unit
    if (r >= rows_pad) return;
    unsigned short out[8];
    const bool real = r < rows;
    const float *x = src + ((real && gather) ? (int64_t)gather[r] : (real ? r : 0)) * ld;
    float nrm = 0.f;
    const int k0 = j * 8;
    const int nk = prec ? d : 3 * d;   // first K' index of the ||r||^2 pieces
    const float sc = (prec && hdr) ? hdr->scale16 : 1.f;   // single FP16: rows scaled by a power of two
    if (real && !is_query && k0 + 8 > nk && k0 < nk + 3) {
        for (int c = 0; c < d; ++c) nrm = fmaf(x[c], x[c], nrm);
        nrm *= sc * sc;
    }
    if (prec && hdr && real && !is_query && j == 0) {   // exact rounding residual of this row, max over rows -> header
        float res = 0.f;
        for (int c = 0; c < d; ++c) {
            const float v = x[c] * sc;
            const float dl = tc_f16_residual(v, f16_rn(v));
            res = fmaf(dl, dl, res);
        }
        atomicMax(reinterpret_cast<int *>(&hdr->dr16), __float_as_int(sqrtf(res) * 1.0000002f));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int kk = k0 + e;
        unsigned short v = 0;
        if (prec) {   // single FP16: [x (d) | two fp16 pieces of ||r||^2  /  1 1]
            if (kk < d) {
                if (real) v = f16_rn((is_query ? -2.f * x[kk] : x[kk]) * sc);
            } else if (kk < d + 2) {
                const int p = kk - d;
                if (is_query) {
                    v = real ? 0x3c00 : 0;  // 1.0
                } else if (real) {
                    const unsigned short n1 = f16_rn(nrm);
                    v = p == 0 ? n1 : f16_rn(nrm - f16_to_f(n1));
                } else {
                    v = p == 0 ? 0x7c00 : 0;  // +inf
                }
            }
        } else if (kk < 3 * d) {
            if (real) {
                const int seg = kk / d, c = kk - seg * d;
                const float val = is_query ? -2.f * x[c] : x[c];
                const unsigned short hi = bf16_rn(val);
                const bool want_lo = is_query ? (seg == 2) : (seg == 1);
                v = want_lo ? bf16_rn(val - bf16_to_f(hi)) : hi;
            }
        } else if (kk < 3 * d + 3) {
            const int p = kk - 3 * d;
            if (is_query) {
                v = real ? 0x3f80 : 0;  // 1.0
            } else if (real) {
                const unsigned short n1 = bf16_rn(nrm);
                const float r1 = nrm - bf16_to_f(n1);
                const unsigned short n2 = bf16_rn(r1);
                const unsigned short n3 = bf16_rn(r1 - bf16_to_f(n2));
                v = p == 0 ? n1 : (p == 1 ? n2 : n3);
            } else {
                v = p == 0 ? 0x7f80 : 0;  // +inf: padding rows can never be selected
            }
        }
        out[e] = v;
    }
    const int64_t tile = r / TC_BM;
    const int rr = (int)(r % TC_BM);
    unsigned char *p = dst + tile * tc_tile_bytes(kp) + (size_t)j * 2048 + (rr / 8) * 128 + (rr % 8) * 16;
    uint4 w;
    w.x = out[0] | ((unsigned)out[1] << 16);
    w.y = out[2] | ((unsigned)out[3] << 16);
    w.z = out[4] | ((unsigned)out[5] << 16);
    w.w = out[6] | ((unsigned)out[7] << 16);
    *reinterpret_cast<uint4 *>(p) = w;
}

// Row-major BF16x3 query rows [rows_pad][kp] (the A operand lives in tensor memory: each selection thread
// loads its own row).  One thread = one row x one k-unit, as above.
__global__ void __launch_bounds__(256)
tc_pack_rows_kernel(const float *__restrict__ src, int ld, int64_t rows, int64_t rows_pad, int d, int kp,
                    const int32_t *__restrict__ gather, unsigned char *__restrict__ dst, int prec,
                    const KnnHeader *__restrict__ hdr, float *__restrict__ dq16) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;  // a warp walks the k-units of a row
    if (r >= rows_pad) return;
    const bool real = r < rows;
    const float *x = src + ((real && gather) ? (int64_t)gather[r] : (real ? r : 0)) * ld;
    const float sc = (prec && hdr) ? hdr->scale16 : 1.f;
    float res = 0.f;   // single FP16: squared rounding residual of this row's operand, summed over the lanes
    for (int j = threadIdx.x % 32; j < kp / 8; j += 32) {
        unsigned short out[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kk = j * 8 + e;
            unsigned short v = 0;
            if (prec) {
                if (real && kk < d) {
                    const float val = -2.f * x[kk] * sc;
                    v = f16_rn(val);
                    const float dl = tc_f16_residual(val, v);   // inf / NaN on overflow: the row then fails its certificate
                    res = fmaf(dl, dl, res);
                } else if (real && kk < d + 2) {
                    v = 0x3c00;
                }
            } else if (real && kk < 3 * d) {
                const int seg = kk / d, c = kk - seg * d;
                const float val = -2.f * x[c];
                const unsigned short hi = bf16_rn(val);
                v = seg == 2 ? bf16_rn(val - bf16_to_f(hi)) : hi;
            } else if (real && kk < 3 * d + 3) {
                v = 0x3f80;  // 1.0 against the three ||r||^2 pieces
            }
            out[e] = v;
        }
        uint4 w;
        w.x = out[0] | ((unsigned)out[1] << 16);
        w.y = out[2] | ((unsigned)out[3] << 16);
        w.z = out[4] | ((unsigned)out[5] << 16);
        w.w = out[6] | ((unsigned)out[7] << 16);
        *reinterpret_cast<uint4 *>(dst + (size_t)r * kp * 2 + (size_t)j * 16) = w;
    }
    if (dq16 != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
        if ((threadIdx.x & 31) == 0) dq16[r] = sqrtf(res) * 1.0000002f;
    }
}

// scale16 = the power of two that brings the largest reference norm into (8, 16]
__global__ void tc_scale_kernel(KnnHeader *hdr) {
    const float rmax = sqrtf(hdr->rmax2);
    float s = 1.f;
    if (rmax > 0.f && isfinite(rmax)) {
        int e;
        frexpf(rmax, &e);            // rmax = m * 2^e, m in [0.5, 1)
        s = ldexpf(1.f, 4 - e);      // rmax * s = m * 16 in [8, 16)
    }
    hdr->scale16 = s;
    hdr->dr16 = 0.f;
}

// kind::f16 instruction descriptor: D=F32 (bit 4), A=B=BF16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);   // A = B = F16 (format 0)

#ifdef TC_PROFILE
// Cycle accounting of the scan (build with -DTC_PROFILE; read back with sym_debug_tc_profile):
//  [0] selection: wait for a finished accumulator   [1] tcgen05.ld + wait   [2] examine (mins, park)   [3] drain
//  [4] selection warp-tiles   [5] warp-tiles in which some lane had a hit   [6] lanes with a hit   [7] mins + vote
//  [14] periodic drains   [15] early drains (full FIFO)
//  [8] issuer: wait reference tile   [9] wait free accumulator   [10] wait turn   [11] issue   [12] issuer tiles
__device__ unsigned long long tc_prof[16];
#define TC_CLK(x) const long long x = clock64()
#define TC_ACC(slot, val) prof_##slot += (val)
#else
#define TC_CLK(x)
#define TC_ACC(slot, val)
#endif

// ------------------------------------------------------------------------------------------ scan kernel
// A CTA owns TC_QT query tiles (256 rows) and streams the reference tiles once for both.
//   warp 0       : producer (cp.async.bulk of one packed tile per stage)
//   warps 1..QT  : MMA issuers, one per query tile (warp 1 also owns the TMEM allocation)
//   warps 4..    : selection; warp w serves query tile (w-4)/4, TMEM lane quadrant w%4, thread <-> query row
// TMEM: the accumulators occupy the last nbuf*128 columns; in front of them the query operands when they live in
// tensor memory (TS form).  With the query operand in shared memory (SS form) there are four accumulators, two
// per query tile; otherwise 2-3, rotated through both query tiles in global issue order.
//
// Selection per tile: four tcgen05.ld (128 scores per row), the accumulator handed back, 16 independent
// 3-input-min chains over groups of 8, one compare against the row's threshold; on a hit the tile is parked (see
// below) and drained into the row's heap later.

__device__ __forceinline__ float tc_min8(const uint32_t *v) {
    float m = fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    m = fmin3(m, __uint_as_float(v[3]), __uint_as_float(v[4]));
    m = fmin3(m, __uint_as_float(v[5]), __uint_as_float(v[6]));
    return fminf(m, __uint_as_float(v[7]));
}

// (Padding needs no bounds check: padded reference rows carry ||r||^2 = +inf and padded query rows are all-zero
// operands, so their scores are +inf / NaN and never pass `s < thr`.)
// The common case (no hit anywhere in the tile) is one straight-line block: the group minima of ALL chunks of
// the tile (16 independent 3-input-min chains for 128 scores), one overall minimum, one compare, one vote.
struct TcMins {
    float g0, g1, g2, g3;
};
__device__ __forceinline__ TcMins tc_mins32(const uint32_t (&v)[32]) {
    return TcMins{tc_min8(v), tc_min8(v + 8), tc_min8(v + 16), tc_min8(v + 24)};
}
__device__ __forceinline__ float tc_min4(const TcMins &m) { return fminf(fmin3(m.g0, m.g1, m.g2), m.g3); }

// The row's kc best candidates form a 4-ary MAX-HEAP of 64-bit keys (score bits | position) in shared memory;
// the root is the kc-th best = the row's threshold.  A hit replaces the root and sifts down: ~log4(kc) levels of
// four independent 8-byte loads, three compares and one 8-byte store, instead of a pass over the list.
// Every lane works on ITS OWN row (lanes run in lock-step when several rows hit at once); the row stride `ks`
// is odd, so equal heap levels of the 32 rows of a warp fall into different banks.  Returns the new threshold.
// (A warp-cooperative variant — one row at a time, REDUX for the maximum — was measured 2x slower end to end:
// its votes and shuffles serialise on the critical path.  Deferring the updates into the wait for the next
// accumulator was slower too: the selection warps have no such slack once updates are frequent.)
__host__ __device__ inline int tc_heap_stride(int kc) { return 1 + 4 * ((kc - 1 + 3) / 4); }   // odd, >= kc

__device__ __forceinline__ float tc_insert(unsigned long long *heap, int kc, float s, uint32_t id) {
    // entries are (raw float score bits << 32 | position); only the scores are compared (one FSETP per level) —
    // equal scores may settle in either order, the float64 re-rank orders the survivors by (distance, index).
    // 4-ary: node i has children 4i+1..4i+4, so kc = 18 is two levels below the root instead of four; the four
    // child loads of a level are independent.  Slots kc..stride-1 hold -inf sentinels (never the largest child).
    const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | id;
    int i = 0;
    float root = s;   // score at the root after the update (no read-back of what was just stored)
#pragma unroll 1
    while (true) {
        const int c = 4 * i + 1;
        if (c >= kc) break;
        const unsigned long long v0 = heap[c], v1 = heap[c + 1], v2 = heap[c + 2], v3 = heap[c + 3];
        const float f0 = __uint_as_float((uint32_t)(v0 >> 32)), f1 = __uint_as_float((uint32_t)(v1 >> 32));
        const float f2 = __uint_as_float((uint32_t)(v2 >> 32)), f3 = __uint_as_float((uint32_t)(v3 >> 32));
        const bool b01 = f1 > f0, b23 = f3 > f2;
        const float fa = b01 ? f1 : f0, fb = b23 ? f3 : f2;
        const unsigned long long va = b01 ? v1 : v0, vb = b23 ? v3 : v2;
        const bool bab = fb > fa;
        const float fc = bab ? fb : fa;
        if (!(fc > s)) break;
        heap[i] = bab ? vb : va;
        if (i == 0) root = fc;
        i = c + (bab ? 2 + (int)b23 : (int)b01);
    }
    heap[i] = key;
    return root;
}


// ---- what is selected: GROUPS of 8 consecutive reference positions, keyed by the group's minimum score ----
// The fast path already has the minimum of every group of 8 scores (the leaves of its min tree).  Selecting groups
// instead of single references means nothing ever has to look INSIDE a group (registers cannot be indexed: round 1
// walked the 8 scores of a hit group through a 16-way switch, ~250 instructions and 4.5 KB of code per event).
// The float64 re-rank expands each of a row's kc candidate groups into its 8 members and ranks those whose exact
// score is below the row's threshold.  The certificate is unchanged: a reference outside the candidate groups has
// score >= its group's minimum >= the row's final threshold.
//
// ---- a hit is parked raw; all rows of the CTA are brought up to date at the SAME tiles ----
// Measured (profiles/r2b, ncu source view): every instruction of a selection warp costs 5-10 cycles (two warps per
// scheduler, dependent chains, ~15 cycles per taken branch), and the four warps that share an accumulator advance
// at the pace of the slowest.  A warp has a hit in 26 % of its tiles, so in 70 % of the tiles one of the four is on
// its slow path — whatever that path costs sits on everybody's critical path (a located + appended hit: ~85
// instructions, ~800 cycles; 36 % of a selection warp's time was spent waiting for an accumulator).
// So the hit path does next to nothing: a lane whose row minimum is below its threshold stores the 16 group
// minima of the tile (four 16-byte stores) and the tile's first position into its row's small FIFO.  Every
// `period` tiles — the same tiles for every warp of the CTA — all rows are drained, lanes in lock-step: the parked
// minima are compared with the threshold and the hit groups go into the row's heap.  Between drains the threshold
// is stale (never too small: rejected scores are >= the threshold in force >= the final one, which is all the
// certificate needs; a parked score is re-checked against the root when it is drained — tc_insert must never be
// given a score that is not below the current root).  A full FIFO drains the whole warp early.
constexpr int TC_GROUP = 8;        // reference positions per candidate group
// one parked tile of a row: GN group minima, the first position, padding up to an ODD number of 16-byte units
// (GN = 16: 20 words; GN = 8: 12 words), so that the 128-bit accesses of 8 consecutive rows fall into 8 bank groups
__host__ __device__ inline int tc_fifo_words(int gn) { return gn + 4; }
__host__ __device__ inline int tc_fifo_stride(int depth, int gn) { return 4 * ((depth * (tc_fifo_words(gn) / 4)) | 1); }

// ---- two warps per (query tile, lane quadrant) ----
// With 128 scores per thread only 8 selection warps fit in the register file (two per scheduler), and a lone warp
// issues an instruction every 5-10 cycles (dependent chains, branches).  With HV = 2 a row is served by TWO threads
// in different warps (same TMEM lanes): each pulls 64 of the tile's 128 columns, has its own FIFO of parked tiles
// (8 group minima each) and its own copy of the threshold, and needs ~100 registers — four selection warps per
// scheduler.  The row still has ONE heap: a warp drains its FIFO into it under the pair's lock (a shared-memory
// word, taken by lane 0 for the whole warp; drains are rare), starting from the heap's current root, which its
// partner may have lowered in the meantime.  A warp's threshold copy is then as fresh as its own last drain — stale
// the safe way round.
__device__ __forceinline__ void tc_pair_lock(uint32_t *lock, int lane) {
    if (lane == 0) {
        while (atomicCAS(lock, 0u, 1u) != 0u) {}
    }
    __syncwarp();
    __threadfence_block();
}
__device__ __forceinline__ void tc_pair_unlock(uint32_t *lock, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(lock, 0u);
}

// One drain step: every row of the warp that has parked tiles puts its NEWEST one into its heap (the order does not
// matter) and keeps the others for the next step.  One tile per row and step makes all steps — and so the drains of
// the four warps that share an accumulator, which happen at the same tiles — take about the same time: a step used
// to take as long as the fullest FIFO of the warp (a second tile in one of 32 rows doubled it; the other three warps
// then waited for this one at the next accumulator).  Only ~3 % of the rows with a parked tile have a second one.
template <int GN>
__device__ __noinline__ float tc_drain(unsigned long long *heap, int kc, const float *fifo, int cnt, float thr,
                                       uint32_t *lock, int lane) {
    constexpr int FW = GN + 4;
    if (lock != nullptr) {
        tc_pair_lock(lock, lane);
        thr = __uint_as_float((uint32_t)(heap[0] >> 32));   // the root as the partner left it
    }
    if (cnt > 0) {
        const float *f = fifo + (cnt - 1) * FW;
        const uint32_t ref0 = __float_as_uint(f[GN]);
        unsigned hm = 0;
#define TC_BIT(GV, B) hm |= (GV < thr) ? (1u << (B)) : 0u;
        {
            const float4 a = *reinterpret_cast<const float4 *>(f), b = *reinterpret_cast<const float4 *>(f + 4);
            TC_BIT(a.x, 0) TC_BIT(a.y, 1) TC_BIT(a.z, 2) TC_BIT(a.w, 3) TC_BIT(b.x, 4) TC_BIT(b.y, 5) TC_BIT(b.z, 6) TC_BIT(b.w, 7)
        }
        if (GN == 16) {
            const float4 c = *reinterpret_cast<const float4 *>(f + 8), d = *reinterpret_cast<const float4 *>(f + 12);
            TC_BIT(c.x, 8) TC_BIT(c.y, 9) TC_BIT(c.z, 10) TC_BIT(c.w, 11) TC_BIT(d.x, 12) TC_BIT(d.y, 13) TC_BIT(d.z, 14) TC_BIT(d.w, 15)
        }
#undef TC_BIT
#pragma unroll 1
        while (hm) {
            const int gi = __ffs(hm) - 1;
            hm &= hm - 1;
            const float gm = f[gi];
            if (gm < thr) thr = tc_insert(heap, kc, gm, ref0 + (uint32_t)(gi * TC_GROUP));
        }
    }
    if (lock != nullptr) tc_pair_unlock(lock, lane);
    return thr;
}

// ------------------------------------------------------------------------------------------ exact tile pruning (opt-in)
// Every 128-position tile of the packed reference (positions are in coarse-cluster order, so a tile is a compact
// piece of the data) has a bounding ball (centre, radius), computed once at fit time; every CTA of a scan has one for
// its 256 query rows.  No row of the CTA can have a neighbour closer than  LB = max(0, |c_q - c_t| - r_q - r_t)^2  in
// the tile, so the tile is skipped when LB >= max over the CTA's rows of (row threshold + ||q||^2): every reference of
// a skipped tile has EXACT score >= that row's threshold at the time >= its final threshold — the certificate of the
// re-rank holds as it stands (for skipped references even without the scan's rounding error).
// The producer warp decides, 32 tiles at a time (one per lane), from the thresholds the selection threads publish after
// every drain (stale = larger = safe); the tiles that remain flow to the issuers and the selection warps through a ring
// of tile numbers in shared memory, closed by a -1.
struct TcPrune {
    const float *tile_c, *tile_r;   // reference tiles: centre [ntiles][d4], radius [ntiles]
    const float *qb_c, *qb_r;       // groups of 32 query rows: centre [groups][d4], radius [groups] (-1: empty group)
    const float *qn2;               // ||q||^2 per row, scan order
    const KnnHeader *hdr;
    unsigned long long *stats;      // {tiles scanned, tiles of a full scan}
    int d, d4, prec;
};
__device__ __forceinline__ float tc_round_up(float x) { return x + fabsf(x) * 1e-4f; }
constexpr int TC_RING = 32;         // tile numbers in flight between producer and selection (stages + accumulators < 32)

// one warp per reference tile: centre = mean of its rows, radius = largest distance to it (rounded up)
__global__ void __launch_bounds__(256)
tc_tile_bounds_kernel(const float *__restrict__ refT, int64_t Mpad, int64_t M, int d, int d4, float *__restrict__ tile_c,
                      float *__restrict__ tile_r) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= Mpad / TC_BN) return;
    const int64_t p0 = tile * TC_BN + lane * 4;
    const int cnt = (int)max((int64_t)0, min((int64_t)TC_BN, M - tile * TC_BN));
    float *c = tile_c + tile * d4;
    float r2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < d4; ++j) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < d) v = *reinterpret_cast<const float4 *>(refT + (int64_t)j * Mpad + p0);   // padding positions hold 0
        float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = cnt > 0 ? sum / (float)cnt : 0.f;
        if (lane == 0) c[j] = mean;
        r2[0] = fmaf(v.x - mean, v.x - mean, r2[0]);
        r2[1] = fmaf(v.y - mean, v.y - mean, r2[1]);
        r2[2] = fmaf(v.z - mean, v.z - mean, r2[2]);
        r2[3] = fmaf(v.w - mean, v.w - mean, r2[3]);
    }
    float m = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (p0 + e < M) m = fmaxf(m, r2[e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) tile_r[tile] = sqrtf(m) * 1.0001f + 1e-30f;
}

// one warp per group of 32 query rows (= the rows of one selection warp), `rows`/32 groups per scan CTA: centre and
// radius of the group (radius -1: no real row in the group), ||q||^2 of every row (scan order).
// (One ball per CTA is not enough: a CTA that straddles two clusters of the sorted queries has a ball as large as the
// distance between them and prunes nothing.)
__global__ void __launch_bounds__(256)
tc_qbounds_kernel(const float *__restrict__ Q, int ldq, int64_t N, const int32_t *__restrict__ q_perm, int d, int d4,
                  float *__restrict__ qb_c, float *__restrict__ qb_r, float *__restrict__ qn2) {
    __shared__ float c_s[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t group = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t row = group * 32 + lane;
    const bool real = row < N;
    const float *q = Q + (real ? (q_perm ? (int64_t)q_perm[row] : row) : 0) * ldq;
    const int cnt = __popc(__ballot_sync(0xffffffffu, real));
    float n2 = 0.f;
    for (int j = 0; j < d4; ++j) {
        const float v = (real && j < d) ? q[j] : 0.f;
        n2 = fmaf(v, v, n2);
        float sum = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = cnt > 0 ? sum / (float)cnt : 0.f;
        if (lane == 0) {
            c_s[warp][j] = mean;
            qb_c[group * d4 + j] = mean;
        }
    }
    __syncwarp();
    float r2 = 0.f;
    if (real)
        for (int j = 0; j < d; ++j) {
            const float df = q[j] - c_s[warp][j];
            r2 = fmaf(df, df, r2);
        }
    qn2[row] = n2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if (lane == 0) qb_r[group] = cnt > 0 ? sqrtf(r2) * 1.0001f + 1e-30f : -1.f;
}

// One chain of MMAs: D[tmem] = A (shared-memory descriptor or tensor memory) x B (shared-memory descriptor), KS
// slices of K = 16 (KS = 0: `ksteps` at run time).  Everything the chain needs is computed BEFORE the issuer waits for
// the accumulator: what follows the wait is on the critical path of every tile (accumulator released -> refilled).
template <bool SS, int KS>
__device__ __forceinline__ void tc_issue_chain(uint32_t d_tmem, uint64_t q_desc, uint32_t a_tmem, uint64_t r_desc,
                                               uint32_t idesc, int ksteps) {
    if (KS > 0) {
#pragma unroll
        for (int k = 0; k < KS; ++k) {   // one K=16 slice: 8 TMEM columns of A, 4096 B (256 units) of B
            if (SS) tc_mma_bf16(d_tmem, q_desc + (uint64_t)(k * 256), r_desc + (uint64_t)(k * 256), idesc, k > 0);
            else tc_mma_bf16_ts(d_tmem, a_tmem + (uint32_t)(k * 8), r_desc + (uint64_t)(k * 256), idesc, k > 0);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
            if (SS) tc_mma_bf16(d_tmem, q_desc + (uint64_t)(k * 256), r_desc + (uint64_t)(k * 256), idesc, k > 0);
            else tc_mma_bf16_ts(d_tmem, a_tmem + (uint32_t)(k * 8), r_desc + (uint64_t)(k * 256), idesc, k > 0);
        }
    }
}

template <int TC_QT, int HV, int SS_MODE, int PRUNE>
__global__ void __launch_bounds__(128 + 128 * TC_QT * HV, 1)
knn_scan_tc_kernel(const unsigned char *__restrict__ qrows, const unsigned char *__restrict__ rpack, int64_t N,
                   const int32_t *__restrict__ q_perm, const int32_t *__restrict__ q_code,
                   const int32_t *__restrict__ cl_tile, int64_t M, int kp, int kc, int stages, int nbuf,
                   int tiles_per_split, int ntiles_all, uint32_t idesc, int fifo_depth, int period_mask,
                   unsigned long long *__restrict__ cand_out, float *__restrict__ thr_out, const TcPrune pr) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const uint32_t tile_bytes = (uint32_t)tc_tile_bytes(kp);
    unsigned char *r_s = smraw;                                        // [stages][tile_bytes]
    // HV selection warps per (query tile, TMEM lane quadrant), each examining 128 / HV columns of every tile
    constexpr int GN = 16 / HV;              // candidate groups per warp and tile
    const int ks = tc_heap_stride(kc);       // odd row stride: the 32 rows of a warp hit 32 different banks
    const int fs = tc_fifo_stride(fifo_depth, GN);
    float *fifo_s = (float *)(r_s + (size_t)stages * tile_bytes);                                  // [HV][TC_QT][TC_BM][fs]
    unsigned long long *heap_s = (unsigned long long *)(fifo_s + (size_t)HV * TC_QT * TC_BM * fs);   // [TC_QT][TC_BM][ks]
    uint64_t *bars = (uint64_t *)(heap_s + (size_t)TC_QT * TC_BM * ks);
    uint64_t *r_full = bars;                      // [TC_MAX_STAGES]
    uint64_t *r_empty = r_full + TC_MAX_STAGES;   // [TC_MAX_STAGES]
    uint64_t *t_full = r_empty + TC_MAX_STAGES;   // [4]  accumulator buffers, used in rotation by all query tiles
    uint64_t *t_empty = t_full + 4;               // [4]
    uint32_t *tmem_slot = (uint32_t *)(t_empty + 4);
    volatile uint32_t *turn = tmem_slot + 1;      // sequence number of the next chain allowed to issue
    uint32_t *pair_lock = tmem_slot + 4;          // [TC_QT * 4] one per (query tile, quadrant): serialises heap drains (HV = 2)
    // pruning: tile numbers on their way from the producer to the selection warps, the rows' published thresholds,
    // the CTA's centre (after the 1 KB that holds the barriers; unused otherwise)
    int *ring_s = (int *)(tmem_slot + 4 + 4 * TC_QT);           // [TC_RING]
    float *rowT_s = (float *)(ring_s + TC_RING);                 // [TC_QT * TC_BM]
    float *cq_s = rowT_s + TC_QT * TC_BM;                        // [TC_QT * 4][d4 <= 84]: centres of the CTA's row groups
    // SS experiment: the query operand stays in shared memory (UMMA tile layout) instead of tensor memory
    unsigned char *q_s = (unsigned char *)(((uintptr_t)(cq_s + TC_QT * 4 * 84) + 1023) & ~(uintptr_t)1023);   // [TC_QT][tile_bytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int t_beg = split * tiles_per_split;
    const int t_end = min(ntiles_all, t_beg + tiles_per_split);
    const int ntiles = max(0, t_end - t_beg);
    // Scan order: cyclic, starting at the reference tile where the cluster of this CTA's first query row begins
    // (queries and reference are both sorted by coarse cluster): the likely neighbours come first, the row
    // thresholds are tight after a few tiles and the rest of the scan is almost free of list updates.
    int t_home = t_beg;
    if (cl_tile != nullptr && ntiles > 0) {
        const int64_t r0 = min((int64_t)blockIdx.x * TC_QT * TC_BM, N - 1);
        const int t = cl_tile[q_code[q_perm ? q_perm[r0] : r0]];
        t_home = min(max(t, t_beg), t_end - 1);
    }
    // tensor memory map: A operand of query tile q in columns [q*kp/2, (q+1)*kp/2); accumulator b at acc0 + b*128
    constexpr bool ss_mode = SS_MODE != 0;
    const uint32_t a_cols = ss_mode ? 0u : (uint32_t)(kp / 2);
    const uint32_t acc0 = 512u - (uint32_t)nbuf * TC_BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAX_STAGES; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], TC_QT); }
        for (int i = 0; i < 4; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4 * HV); }
        *turn = 0u;
        for (int i = 0; i < 4 * TC_QT; ++i) pair_lock[i] = 0u;
    }
    if (PRUNE) {
        for (int i = threadIdx.x; i < TC_QT * TC_BM; i += blockDim.x) rowT_s[i] = INFINITY;
        for (int i = threadIdx.x; i < TC_QT * 4 * pr.d4; i += blockDim.x) cq_s[i] = pr.qb_c[(size_t)blockIdx.x * TC_QT * 4 * pr.d4 + i];
    }
    if (threadIdx.x == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns, owned by this warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 4 && warp < 4 + 4 * TC_QT) {  // one thread per row initialises its heap (all slots empty = the largest key)
        const int list = (warp - 4) >> 2, row = (warp & 3) * 32 + lane;   // list = query tile
        for (int e = 0; e < ks; ++e)
            heap_s[((size_t)list * TC_BM + row) * ks + e] = ((e < kc ? 0x7f800000ull : 0xff800000ull) << 32) | 0xffffffffull;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 4 && warp < 4 + 4 * TC_QT) {
        // park this thread's query row (BF16x3, kp values) in tensor memory: it is the A operand of every MMA
        const int qt = (warp - 4) >> 2, quad = warp & 3;
        const int64_t n = ((int64_t)blockIdx.x * TC_QT + qt) * TC_BM + quad * 32 + lane;
        const uint4 *src = reinterpret_cast<const uint4 *>(qrows + (size_t)n * kp * 2);
        const uint32_t dst = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)qt * a_cols;
        if (ss_mode) {
            const int r = quad * 32 + lane;
            unsigned char *qt_s = q_s + (size_t)qt * tile_bytes + (r / 8) * 128 + (r % 8) * 16;
            for (int j = 0; j < kp / 8; ++j) *reinterpret_cast<uint4 *>(qt_s + (size_t)j * 2048) = __ldg(src + j);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
            for (int k = 0; k < kp / 16; ++k) tc_st8(dst + k * 8, __ldg(src + 2 * k), __ldg(src + 2 * k + 1));
            tc_wait_st();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // Register reallocation between the warp groups (the launch bound of 168 registers per thread is what 384
    // threads can share evenly): the producer / issuer group needs few, the selection warps hold a whole
    // accumulator row (128 scores) plus the group minima and spill without the extra room.
    // (setmaxnreg moves registers only within what the CTA was launched with: what the four service warps give up
    // must cover what the selection warps ask for, or the last of them spins in the allocation forever.)
    if (TC_QT * HV == 4) {
        // 640 threads x 96 registers at launch; a selection thread holds 64 scores and needs < 96: nothing to move
    } else {
        if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    }
    if (warp == 0) {
        // ===== producer: one bulk copy per reference tile =====
        if (PRUNE) {
            // the whole warp walks the cyclic tile order 32 tiles at a time: lane <-> tile, bound against the CTA's
            // largest published (threshold + ||q||^2); lane 0 issues the copies of the tiles that remain
            constexpr int NG = TC_QT * 4;   // groups of 32 rows (one per selection warp), each with its own ball
            float rq[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) rq[g] = pr.qb_r[(size_t)blockIdx.x * NG + g];
            int s = 0, issued = 0;
            uint32_t ph = 0;
            for (int base = 0; base < ntiles; base += 32) {
                float tc[NG];   // per group: the largest published (threshold + ||q||^2) of its rows
                bool all_finite = true;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float v = ((volatile float *)rowT_s)[g * 32 + lane];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    tc[g] = v;
                    all_finite &= v < INFINITY;
                }
                const int i = base + lane;
                int t = t_home + i;
                if (t >= t_end) t -= ntiles;
                bool keep = i < ntiles;
                if (keep && all_finite) {
                    const float *c = pr.tile_c + (size_t)t * pr.d4;
                    float d2[NG];
#pragma unroll
                    for (int g = 0; g < NG; ++g) d2[g] = 0.f;
                    for (int j = 0; j < pr.d; ++j) {
                        const float cj = __ldg(c + j);
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            const float df = cq_s[g * pr.d4 + j] - cj;
                            d2[g] = fmaf(df, df, d2[g]);
                        }
                    }
                    const float rt = __ldg(pr.tile_r + t);
                    keep = false;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if (rq[g] < 0.f) continue;   // no real row in this group
                        const float gap = sqrtf(d2[g]) * 0.9999f - rq[g] - rt;   // (rounded towards "keep")
                        keep |= !(gap > 0.f && gap * gap * 0.9999f >= tc[g]);
                    }
                }
                unsigned mask = __ballot_sync(0xffffffffu, keep);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int tb = __shfl_sync(0xffffffffu, t, b);
                    if (lane == 0) {
                        mbar_wait_sleepy(&r_empty[s], ph ^ 1);
                        ring_s[issued & (TC_RING - 1)] = tb;
                        mbar_arrive_expect_tx(&r_full[s], tile_bytes);
                        bulk_g2s(r_s + (size_t)s * tile_bytes, rpack + (size_t)tb * tile_bytes, tile_bytes, &r_full[s]);
                    }
                    ++issued;
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
            if (lane == 0) {   // end of the tile stream
                mbar_wait_sleepy(&r_empty[s], ph ^ 1);
                ring_s[issued & (TC_RING - 1)] = -1;
                mbar_arrive(&r_full[s]);
                atomicAdd(pr.stats, (unsigned long long)issued);
                atomicAdd(pr.stats + 1, (unsigned long long)ntiles);
            }
        } else if (lane == 0) {
            int s = 0, t = t_home;
            uint32_t ph = 0;  // ring position and phase kept incrementally: no integer division on any per-tile path
            for (int i = 0; i < ntiles; ++i) {
                mbar_wait_sleepy(&r_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&r_full[s], tile_bytes);
                bulk_g2s(r_s + (size_t)s * tile_bytes, rpack + (size_t)t * tile_bytes, tile_bytes, &r_full[s]);
                if (++s == stages) { s = 0; ph ^= 1; }
                if (++t == t_end) t = t_beg;
            }
        }
    } else if (warp <= TC_QT) {
        // ===== MMA issuers: warp 1+q issues the chains of query tile q.  One warp's serial instruction stream
        // (barrier poll, descriptor moves, 6 UTCHMMA, commit) costs about as much as the 6 MMAs take to
        // execute, so each query tile gets its own issuer and the two streams overlap.  Chains are still
        // ISSUED in sequence order (tile i: q0, q1; tile i+1: q0, ...) via the `turn` counter: the tensor
        // pipe completes them in that order, which is what lets all query tiles rotate through one set of
        // accumulators (a parity wait on a buffer's barrier is only unambiguous if its uses complete in order). =====
        const int q = warp - 1;
        const int ksteps = kp / 16;
        const bool leader = elect_one();
        // when every query tile owns its accumulators outright (nbuf multiple of QT) the issuers are independent
        const bool ordered = TC_QT > 1 && (nbuf % TC_QT) != 0;
        const uint64_t q_desc = umma_desc(smem_u32(q_s + (size_t)q * tile_bytes));
        int buf = q % nbuf, use = q / nbuf;  // chain (tile i, query tile q) = sequence number i*QT + q
        int s = 0;
        uint32_t ph = 0;
        const uint64_t r_desc0 = umma_desc(smem_u32(r_s));
        const uint32_t r_step = tile_bytes >> 4;
        const uint32_t a_tmem = tmem_base + (uint32_t)q * a_cols;
#ifdef TC_PROFILE
        long long prof_8 = 0, prof_9 = 0, prof_10 = 0, prof_11 = 0;
#endif
        // (the loop is instantiated per chain length — 2 and 4 slices cover the single-FP16 operands up to d = 62 —
        // so that no dispatch sits between the accumulator hand-off and the first MMA)
        auto issue_loop = [&](auto ks_tag) {
        constexpr int KS = decltype(ks_tag)::value;
        uint64_t r_desc = r_desc0;
        for (int i = 0; PRUNE || i < ntiles; ++i) {
            // operands of this chain, ready before the waits
            const uint32_t d_tmem = tmem_base + acc0 + (uint32_t)(buf * TC_BN);
            const uint32_t full_bar = smem_u32(&t_full[buf]), empty_bar = smem_u32(&r_empty[s]);
            asm volatile("" ::"l"(r_desc), "r"(d_tmem), "r"(full_bar), "r"(empty_bar));
            TC_CLK(c0);
            mbar_wait_handoff(&r_full[s], ph);
            TC_CLK(c1);
            mbar_wait_handoff(&t_empty[buf], (use & 1) ^ 1);
            TC_CLK(c2);
            if (ordered) {
                const uint32_t seq = (uint32_t)(i * TC_QT + q);
                while (*turn != seq) {}
            }
            TC_CLK(c3);
            tc_fence_after();
            if (PRUNE && ring_s[i & (TC_RING - 1)] < 0) {   // end of the tile stream: wake the selection warps, stop
                if (leader) {
                    mbar_arrive(&t_full[buf]);
                    if (ordered) *turn = (uint32_t)(i * TC_QT + q) + 1u;
                }
                __syncwarp();
                break;
            }
            if (leader) {
#ifndef TC_EXPERIMENT_NO_MMA
                tc_issue_chain<ss_mode, KS>(d_tmem, q_desc, a_tmem, r_desc, idesc, ksteps);
#endif
                tc_commit_addr(full_bar);    // this chain's accumulator is complete
                tc_commit_addr(empty_bar);   // and this issuer is done with the shared-memory stage
                if (ordered) *turn = (uint32_t)(i * TC_QT + q) + 1u;
            }
            __syncwarp();
            TC_CLK(c4);
            TC_ACC(8, c1 - c0); TC_ACC(9, c2 - c1); TC_ACC(10, c3 - c2); TC_ACC(11, c4 - c3);
            r_desc += r_step;
            if (++s == stages) { s = 0; ph ^= 1; r_desc = r_desc0; }
            buf += TC_QT;
            if (buf >= nbuf) { buf -= nbuf; ++use; }
        }
        };
        if (ksteps == 2) issue_loop(std::integral_constant<int, 2>{});
        else if (ksteps == 4) issue_loop(std::integral_constant<int, 4>{});
        else issue_loop(std::integral_constant<int, 0>{});
#ifdef TC_PROFILE
        if (lane == 0) {
            atomicAdd(&tc_prof[8], (unsigned long long)prof_8); atomicAdd(&tc_prof[9], (unsigned long long)prof_9);
            atomicAdd(&tc_prof[10], (unsigned long long)prof_10); atomicAdd(&tc_prof[11], (unsigned long long)prof_11);
            atomicAdd(&tc_prof[12], (unsigned long long)ntiles);
        }
#endif
    } else if (warp < 4) {
        // spare warp(s): nothing to do until the final barrier
    } else {
        // ===== selection =====
        const int sel = warp - 4;
        const int half = sel / (4 * TC_QT);                  // which 128 / HV columns of every tile this warp examines
        const int qt = (sel % (4 * TC_QT)) >> 2, quad = warp & 3;
        const int row = quad * 32 + lane;
        unsigned long long *heap = heap_s + ((size_t)qt * TC_BM + row) * ks;   // the row's kc best groups
        float *fifo = fifo_s + ((size_t)(half * TC_QT + qt) * TC_BM + row) * fs;   // this thread's parked tiles
        uint32_t *lock = HV == 2 ? pair_lock + qt * 4 + quad : nullptr;
        float thr = INFINITY;
        int qn = 0;                                                            // parked tiles
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + acc0 + (uint32_t)(half * (TC_BN / HV));
        const uint32_t col0 = (uint32_t)(half * (TC_BN / HV));
        int buf = qt % nbuf, use = qt / nbuf;   // chain (i, qt) = sequence number i*QT + qt
        // first position of the current tile's columns; wraps once, from the end of the split back to its start
        // (the wrap target is recomputed from the launch parameters: nothing extra stays live across the loop)
        uint32_t ref_cur = (uint32_t)(t_home * TC_BN) + col0;
        int until_wrap = t_end - t_home;
#ifdef TC_PROFILE
        long long prof_0 = 0, prof_1 = 0, prof_2 = 0, prof_3 = 0, prof_5 = 0, prof_6 = 0, prof_7 = 0, prof_14 = 0, prof_15 = 0;
#endif
        // pruning: what this row publishes after a drain is (threshold + ||q||^2) in exact units
        float pub_scale = 1.f, pub_add = -INFINITY;
        if (PRUNE) {
            const int64_t n_row = ((int64_t)blockIdx.x * TC_QT + qt) * TC_BM + row;
            if (pr.prec) {
                const float s16 = pr.hdr->scale16;
                pub_scale = 1.f / (s16 * s16);
            }
            if (n_row < N) pub_add = pr.qn2[n_row];   // (rows past the end never hold a tile back)
        }
        for (int i = 0; PRUNE || i < ntiles; ++i) {
            TC_CLK(c0);
#ifdef TC_SLEEPY_SELECT
            mbar_wait_sleepy(&t_full[buf], use & 1);
#else
            mbar_wait_handoff(&t_full[buf], use & 1);
#endif
            tc_fence_after();
            uint32_t ref0;
            if (PRUNE) {
                const int t = ring_s[i & (TC_RING - 1)];
                if (t < 0) break;   // end of the tile stream
                ref0 = (uint32_t)(t * TC_BN) + col0;
            } else {
                ref0 = ref_cur;
                ref_cur += TC_BN;
                if (--until_wrap == 0) ref_cur = (uint32_t)(blockIdx.y * tiles_per_split * TC_BN) + col0;
            }
            const uint32_t acc = lane_base + (uint32_t)(buf * TC_BN);
#ifdef TC_EXPERIMENT_NO_SELECT
            tc_fence_before();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
            __syncwarp();
#else
            // pull the accumulator row into registers and hand the TMEM buffer straight back: the next chain can
            // then run while this tile is examined (the buffer is held for one load latency).  (Splitting the load
            // in two with a wait in between, to fold the first half while the second is in flight, was slower:
            // 211 vs 131 cycles — the scoreboard already lets the min chains start as each LDTM lands.)
            uint32_t v0[32], v1[32];
            TcMins m0, m1, m2, m3;
            TC_CLK(c1);
            tc_ld32(acc, v0);
            tc_ld32(acc + 32, v1);
            if (HV == 1) {
                uint32_t v2[32], v3[32];
                tc_ld32(acc + 64, v2);
                tc_ld32(acc + 96, v3);
                tc_wait_ld();
                tc_fence_before();
                if (lane == 0) mbar_arrive(&t_empty[buf]);
                m2 = tc_mins32(v2);
                m3 = tc_mins32(v3);
            } else {
                tc_wait_ld();
                tc_fence_before();
                if (lane == 0) mbar_arrive(&t_empty[buf]);
            }
            TC_CLK(c2);
            m0 = tc_mins32(v0);
            m1 = tc_mins32(v1);
            const float mall = HV == 1 ? fminf(fmin3(tc_min4(m0), tc_min4(m1), tc_min4(m2)), tc_min4(m3))
                                       : fminf(tc_min4(m0), tc_min4(m1));
#ifndef TC_EXPERIMENT_NO_WALK
            const bool hit = mall < thr;
            const bool any_hit = __any_sync(0xffffffffu, hit);
            TC_CLK(c2a);
            TC_ACC(7, c2a - c2);
            if (any_hit) {
                // a hit row with a full FIFO: the whole warp drains first (rare; lanes in lock-step)
                if (__any_sync(0xffffffffu, hit && qn == fifo_depth)) {
                    thr = tc_drain<GN>(heap, kc, fifo, qn, thr, lock, lane);
                    qn = max(qn - 1, 0);
                    TC_ACC(15, 1);
                    if (PRUNE) rowT_s[qt * TC_BM + row] = tc_round_up(fmaf(thr, pub_scale, pub_add));
                }
                if (hit) {   // park the tile's 16 group minima and its first position; nothing else happens here
                    float *f = fifo + qn * (GN + 4);
                    *reinterpret_cast<float4 *>(f) = make_float4(m0.g0, m0.g1, m0.g2, m0.g3);
                    *reinterpret_cast<float4 *>(f + 4) = make_float4(m1.g0, m1.g1, m1.g2, m1.g3);
                    if (HV == 1) {
                        *reinterpret_cast<float4 *>(f + 8) = make_float4(m2.g0, m2.g1, m2.g2, m2.g3);
                        *reinterpret_cast<float4 *>(f + 12) = make_float4(m3.g0, m3.g1, m3.g2, m3.g3);
                    }
                    f[GN] = __uint_as_float(ref0);
                    ++qn;
                }
                __syncwarp();
#ifdef TC_PROFILE
                prof_5 += 1;
                prof_6 += __popc(__ballot_sync(0xffffffffu, hit));
#endif
            }
#else
            if (mall < -1e30f) thr = 0.f;
#endif
            TC_CLK(c3);
            // at the same tiles for all warps of the CTA — every 4th tile at the start of a scan, when thresholds still
            // move fast, every `period`-th later — a parked tile per row goes into the heaps
            if ((i & (i < 128 ? 3 : period_mask)) == (i < 128 ? 3 : period_mask) && __any_sync(0xffffffffu, qn != 0)) {
                thr = tc_drain<GN>(heap, kc, fifo, qn, thr, lock, lane);
                qn = max(qn - 1, 0);
                TC_ACC(14, 1);
                if (PRUNE) rowT_s[qt * TC_BM + row] = tc_round_up(fmaf(thr, pub_scale, pub_add));
            }
            __syncwarp();
#ifdef TC_PROFILE
            prof_0 += c1 - c0; prof_1 += c2 - c1; prof_2 += c3 - c2; prof_3 += clock64() - c3;
#endif
#endif
            // advance the rotation by TC_QT chains (TC_QT <= nbuf)
            buf += TC_QT;
            if (buf >= nbuf) { buf -= nbuf; ++use; }
        }
        while (__any_sync(0xffffffffu, qn != 0)) {
            thr = tc_drain<GN>(heap, kc, fifo, qn, thr, lock, lane);
            qn = max(qn - 1, 0);
        }
        __syncwarp();
        if (HV == 2) {   // both warps of every pair are done with the heaps before they are read out
            asm volatile("bar.sync 1, %0;" ::"r"(128 * TC_QT * HV) : "memory");
            thr = __uint_as_float((uint32_t)(heap[0] >> 32));
        }
#ifdef TC_PROFILE
        if (lane == 0) {
            atomicAdd(&tc_prof[0], (unsigned long long)prof_0); atomicAdd(&tc_prof[1], (unsigned long long)prof_1);
            atomicAdd(&tc_prof[2], (unsigned long long)prof_2); atomicAdd(&tc_prof[3], (unsigned long long)prof_3);
            atomicAdd(&tc_prof[4], (unsigned long long)ntiles); atomicAdd(&tc_prof[5], (unsigned long long)prof_5);
            atomicAdd(&tc_prof[6], (unsigned long long)prof_6); atomicAdd(&tc_prof[7], (unsigned long long)prof_7);
            atomicAdd(&tc_prof[14], (unsigned long long)prof_14); atomicAdd(&tc_prof[15], (unsigned long long)prof_15);
        }
#endif
        // results: the row's candidates as (score, position) keys (unordered) and its final threshold
        const int64_t n = ((int64_t)blockIdx.x * TC_QT + qt) * TC_BM + row;
        if (n < N && half == 0) {
            const int64_t slot = n * nsplit + split;
            for (int e = 0; e < kc; ++e) {
                const unsigned long long h = heap[e];
                const uint32_t pos = (uint32_t)h;
                cand_out[slot * kc + e] = pos == 0xffffffffu ? ~0ull : knn_make_key(__uint_as_float((uint32_t)(h >> 32)), pos);
            }
            const int64_t split_refs = min(M, (int64_t)t_end * TC_BN) - (int64_t)t_beg * TC_BN;
            thr_out[slot] = ceil_div(split_refs, TC_GROUP) <= kc ? INFINITY : thr;  // every group of such a split was handed over
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ debug
// One query tile x one reference tile: dumps the 128x128 FP32 accumulator (test-only entry point).
__global__ void __launch_bounds__(128, 1)
tc_debug_kernel(const unsigned char *__restrict__ qpack, const unsigned char *__restrict__ rpack, int kp, uint32_t idesc,
                float *__restrict__ out, const unsigned char *__restrict__ qrows) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const uint32_t tile_bytes = (uint32_t)tc_tile_bytes(kp);
    unsigned char *q_s = smraw, *r_s = smraw + tile_bytes;
    __shared__ uint64_t full, done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&full, 1);
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (qrows != nullptr) {  // TS form: every thread parks its own query row in tensor memory columns [256, 256 + kp/2)
        const uint4 *src = reinterpret_cast<const uint4 *>(qrows + (size_t)threadIdx.x * kp * 2);
        for (int k = 0; k < kp / 16; ++k)
            tc_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + k * 8, src[2 * k], src[2 * k + 1]);
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&full, 2 * tile_bytes);
        bulk_g2s(q_s, qpack, tile_bytes, &full);
        bulk_g2s(r_s, rpack, tile_bytes, &full);
        mbar_wait(&full, 0);
        tc_fence_after();
        for (int k = 0; k < kp / 16; ++k) {
            if (qrows != nullptr)
                tc_mma_bf16_ts(tmem_base, tmem_base + 256 + k * 8, umma_desc(smem_u32(r_s) + k * 4096), idesc, k > 0);
            else
                tc_mma_bf16(tmem_base, umma_desc(smem_u32(q_s) + k * 4096), umma_desc(smem_u32(r_s) + k * 4096), idesc,
                            k > 0);
        }
        tc_commit(&done);
    }
    __syncwarp();
    mbar_wait(&done, 0);
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
namespace {
struct TcPlan {
    int kp, kc, fifo, period_mask, hv, qt, stages, nbuf, nsplit, tiles_per_split, ntiles, ss;
    int64_t qtiles;
    size_t smem, qpack_bytes, cand_bytes, thr_bytes, dq_bytes, prune_bytes;
    bool ok;
};

TcPlan tc_plan(int64_t N, int64_t M, int d, int k, int prec) {
    TcPlan p{};
    p.kp = tc_kp(d, prec);
    const char *force_pad = getenv("SYM_TC_PAD");   // tuning / experiments only
    // over-selection: 8 groups, 16 for k >= 40 — the gap between the k-th and the (k+8)-th neighbour shrinks relative to
    // the FP16 score error as k grows: at config5 (k = 50) 5 % of the rows failed their certificate with 8 (253k rows
    // re-run with BF16x3: 0.45 s of 3.35 s), 54 rows with 16 (3.23 s in all)
    int kc = k + (force_pad ? atoi(force_pad) : (k >= 40 ? 2 * TC_PAD : TC_PAD));
    if (kc > M) kc = (int)M;
    p.kc = kc;
    const char *force_hv = getenv("SYM_TC_HALVES");      // tuning / experiments only
    // selection warps per (query tile, quadrant): 1.  Two (each examining 64 columns; four selection warps per scheduler)
    // measured 5 % slower at config2: the per-tile overheads double and the pair's drains serialise on its lock.
    p.hv = force_hv ? (atoi(force_hv) == 2 ? 2 : 1) : 1;
    const int gn = 16 / p.hv;
    const char *force_fifo = getenv("SYM_TC_FIFO");
    const char *force_period = getenv("SYM_TC_PERIOD");
    int period = force_period ? atoi(force_period) : 32;   // tiles between two drains (power of two)
    if (period < 1) period = 1;
    while (period & (period - 1)) period &= period - 1;
    p.period_mask = period - 1;
    const size_t tile = tc_tile_bytes(p.kp);
    const size_t budget = 227 * 1024;
    size_t fixed = 0;
    int stages = 0;
    const char *force_qt = getenv("SYM_TC_QT");      // tuning / experiments only
    const char *force_nbuf = getenv("SYM_TC_NBUF");
    const char *force_ss = getenv("SYM_TC_SS");
    // Preference: 256 rows per CTA, then 128; within that, the query operand in SHARED memory when there is room
    // for it next to >= 3 reference stages: tensor memory then holds four accumulators, two per query tile, and the
    // two query tiles stop sharing a rotation (measured 2-5 % faster than the tensor-memory operand with three
    // shared accumulators; the SS form's slower MMA does not matter because selection is the bottleneck).
    const int cfg[4][2] = {{2, 1}, {2, 0}, {1, 1}, {1, 0}};
    for (int c = 0; c < 4; ++c) {
        const int qt = cfg[c][0], ss = cfg[c][1];
        if (force_qt && atoi(force_qt) != qt) continue;
        if (force_ss && (atoi(force_ss) != 0) != (ss != 0)) continue;
        // parked tiles per row: up to 6, as many as fit beside >= 3 reference stages (config2: 4 -> 6 parked tiles
        // = 1.4 % faster: fewer early, unsynchronised drains; 3 stages of reference tiles are as good as 16)
        for (p.fifo = force_fifo ? atoi(force_fifo) : 6; p.fifo >= 1; --p.fifo) {
            fixed = 64 /* pair locks */ + (size_t)qt * TC_BM * ((size_t)tc_heap_stride(kc) * 8 + (size_t)p.hv * tc_fifo_stride(p.fifo, gn) * 4) +
                    512 /* barriers */ + 4096 /* pruning: ring, thresholds, centres */ + 1024 /* alignment slack */ + (ss ? (size_t)qt * tile + 1024 : 0);
            stages = budget > fixed ? (int)((budget - fixed) / tile) : 0;
            if (stages >= 3 || p.fifo == 1 || force_fifo) break;
        }
        if (p.fifo < 1) p.fifo = 1;
        const char *force_stages = getenv("SYM_TC_STAGES");   // tuning / experiments only
        if (force_stages && atoi(force_stages) >= 2 && atoi(force_stages) < stages) stages = atoi(force_stages);
        if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
        p.qt = qt;
        p.ss = ss;
        p.nbuf = (512 - (ss ? 0 : qt * p.kp / 2)) / TC_BN;  // accumulators that fit beside the A operands in tensor memory
        if (p.nbuf > 4) p.nbuf = 4;
        if (force_nbuf && atoi(force_nbuf) >= 1 && atoi(force_nbuf) <= p.nbuf) p.nbuf = atoi(force_nbuf);
        if (stages >= 3 && p.nbuf >= 2) break;
    }
    p.stages = stages;
    p.ok = stages >= 2 && p.nbuf >= 1 && p.kp <= 256 && d <= 84;
    p.smem = fixed + (size_t)(stages > 0 ? stages : 0) * tile;
    if (p.smem < 120 * 1024) p.smem = 120 * 1024;  // one CTA per SM: a CTA owns all 512 TMEM columns
    p.qtiles = ceil_div(N, TC_BM * p.qt);  // CTAs along the query
    p.ntiles = (int)(knn_mpad(M) / TC_BN);
    int nsplit = 1;
    if (p.qtiles < kNumSMs) {
        nsplit = (int)ceil_div(kNumSMs, p.qtiles);
        const int by_work = p.ntiles / 16 > 0 ? p.ntiles / 16 : 1;
        const int by_rerank = 512 / kc > 0 ? 512 / kc : 1;
        nsplit = nsplit < by_work ? nsplit : by_work;
        nsplit = nsplit < by_rerank ? nsplit : by_rerank;
        if (nsplit < 1) nsplit = 1;
    }
    p.tiles_per_split = (int)ceil_div(p.ntiles, nsplit);
    p.nsplit = (int)ceil_div(p.ntiles, p.tiles_per_split);
    p.qpack_bytes = (size_t)p.qtiles * p.qt * TC_BM * p.kp * 2;  // row-major query operand rows
    p.dq_bytes = prec ? (((size_t)p.qtiles * p.qt * TC_BM * sizeof(float) + 255) & ~(size_t)255) : 0;
    // pruning: CTA centres [qtiles][d4], radii [qtiles], ||q||^2 [rows]
    p.prune_bytes = (((size_t)p.qtiles * p.qt * 4 * (((d + 3) & ~3) + 1) + (size_t)p.qtiles * p.qt * TC_BM) * sizeof(float) + 255) & ~(size_t)255;
    p.cand_bytes = ((size_t)N * p.nsplit * kc * 8 + 255) & ~(size_t)255;
    p.thr_bytes = ((size_t)N * p.nsplit * 4 + 255) & ~(size_t)255;
    return p;
}
}  // namespace

size_t knn_tc_packed_bytes(int64_t M, int d, int prec) {
    if (d > 84) return 0;
    return (size_t)(knn_mpad(M) / TC_BN) * tc_tile_bytes(tc_kp(d, prec));
}

bool knn_tc_supported(int64_t N, int64_t M, int d, int k, int prec) {
    if (d > 84 || k > 120) return false;
    return tc_plan(N, M, d, k, prec).ok;
}

size_t knn_tc_workspace(int64_t N, int64_t M, int d, int k, int prec, int) {
    if (!knn_tc_supported(N, M, d, k, prec)) return 0;
    TcPlan p = tc_plan(N, M, d, k, prec);
    return 256 /* scan statistics */ + ((p.qpack_bytes + 255) & ~(size_t)255) + p.cand_bytes + p.thr_bytes + p.dq_bytes +
           p.prune_bytes + 1024;
}

int knn_tc_fit(const float *Ref, int ldr, int64_t M, int d, const int32_t *order, KnnPacked pk, cudaStream_t st) {
    if (d > 84) return SYM_OK;
    const int64_t rows_pad = knn_mpad(M);
    tc_scale_kernel<<<1, 1, 0, st>>>(pk.hdr);   // after knn_fit_kernel (same stream): needs header.rmax2
    SYM_LAUNCH_CHECK("tc_scale_kernel");
    tc_tile_bounds_kernel<<<(unsigned)ceil_div(rows_pad / TC_BN, 8), 256, 0, st>>>(pk.refT, pk.Mpad, M, d, pk.d4, pk.tile_c,
                                                                                    pk.tile_r);   // after knn_fit_kernel: needs refT
    SYM_LAUNCH_CHECK("tc_tile_bounds_kernel");
    for (int prec = 0; prec < 2; ++prec) {
        const int kp = tc_kp(d, prec);
        dim3 grid((unsigned)ceil_div(rows_pad, 256), (unsigned)(kp / 8));
        tc_pack_kernel<<<grid, 256, 0, st>>>(Ref, ldr, M, rows_pad, d, kp, 0, order, prec ? pk.tc16 : pk.tc, prec, pk.hdr);
        SYM_LAUNCH_CHECK("tc_pack_kernel(ref)");
    }
    return SYM_OK;
}

int knn_tc_scan(const float *Q, int ldq, int64_t N, const int32_t *q_perm, const int32_t *q_code,
                const int32_t *cl_tile, KnnPacked pk, int64_t M, int d, int k, int prec, int prune, void *workspace,
                size_t workspace_bytes, cudaStream_t st, const unsigned long long **cand, const float **thr,
                int *ncand, int *nsplit, int *group, double *eps_rel, const float **dq16, const unsigned char **qrows,
                int *kp) {
    TcPlan p = tc_plan(N, M, d, k, prec);
    SYM_REQUIRE(p.ok, SYM_ERR_UNSUPPORTED_SHAPE, "tcgen05 scan: d=%d k=%d does not fit", d, k);
    SYM_REQUIRE(workspace_bytes >= knn_tc_workspace(N, M, d, k, prec) - 1024, SYM_ERR_WORKSPACE_TOO_SMALL,
                "tcgen05 scan: workspace too small");
    unsigned char *w = (unsigned char *)workspace;
    unsigned long long *stats_w = (unsigned long long *)w; w += 256;
    unsigned char *qpack = w; w += (p.qpack_bytes + 255) & ~(size_t)255;
    unsigned long long *cand_w = (unsigned long long *)w; w += p.cand_bytes;
    float *thr_w = (float *)w; w += p.thr_bytes;
    float *dq_w = prec ? (float *)w : nullptr; w += p.dq_bytes;
    const int d4 = (d + 3) & ~3;
    const size_t ngroups = (size_t)p.qtiles * p.qt * 4;   // groups of 32 query rows
    float *qbc_w = (float *)w, *qbr_w = qbc_w + ngroups * d4, *qn2_w = qbr_w + ngroups;
    if (p.hv != 1) prune = 0;   // (the two-warps-per-quadrant experiment has no pruning instantiation)
    TcPrune pr{};
    if (prune) {
        SYM_CUDA(cudaMemsetAsync(stats_w, 0, 16, st));
        tc_qbounds_kernel<<<(unsigned)p.qtiles, p.qt * TC_BM, 0, st>>>(Q, ldq, N, q_perm, d, d4, qbc_w, qbr_w, qn2_w);
        SYM_LAUNCH_CHECK("tc_qbounds_kernel");
        pr = TcPrune{pk.tile_c, pk.tile_r, qbc_w, qbr_w, qn2_w, pk.hdr, stats_w, d, d4, prec};
    }
    {
        const int64_t rows_pad = p.qtiles * p.qt * TC_BM;
        tc_pack_rows_kernel<<<(unsigned)ceil_div(rows_pad, 8), 256, 0, st>>>(Q, ldq, N, rows_pad, d, p.kp, q_perm, qpack,
                                                                                 prec, pk.hdr, dq_w);
        SYM_LAUNCH_CHECK("tc_pack_rows_kernel(query)");
    }
    dim3 grid((unsigned)p.qtiles, (unsigned)p.nsplit);
#define SYM_TC_LAUNCH(QT, HV, SS, PR)                                                                                 \
    do {                                                                                                              \
        SYM_CUDA(cudaFuncSetAttribute(knn_scan_tc_kernel<QT, HV, SS, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)p.smem));                                                                  \
        knn_scan_tc_kernel<QT, HV, SS, PR><<<grid, 128 + 128 * QT * HV, p.smem, st>>>(                                \
            qpack, prec ? pk.tc16 : pk.tc, N, q_perm, q_code, cl_tile, M, p.kp, p.kc, p.stages, p.nbuf,               \
            p.tiles_per_split, p.ntiles, prec ? TC_IDESC_F16 : TC_IDESC_BF16, p.fifo, p.period_mask, cand_w, thr_w,    \
            pr);                                                                                                      \
    } while (0)
#define SYM_TC_LAUNCH_SS(QT, HV, PR)                                                                                  \
    do {                                                                                                              \
        if (p.ss) SYM_TC_LAUNCH(QT, HV, 1, PR);                                                                       \
        else SYM_TC_LAUNCH(QT, HV, 0, PR);                                                                            \
    } while (0)
    if (p.qt == 2 && p.hv == 2) SYM_TC_LAUNCH_SS(2, 2, 0);
    else if (p.qt == 2 && prune) SYM_TC_LAUNCH_SS(2, 1, 1);
    else if (p.qt == 2) SYM_TC_LAUNCH_SS(2, 1, 0);
    else if (p.hv == 2) SYM_TC_LAUNCH_SS(1, 2, 0);
    else if (prune) SYM_TC_LAUNCH_SS(1, 1, 1);
    else SYM_TC_LAUNCH_SS(1, 1, 0);
#undef SYM_TC_LAUNCH_SS
#undef SYM_TC_LAUNCH
    SYM_LAUNCH_CHECK("knn_scan_tc_kernel");
    *cand = cand_w;
    *thr = thr_w;
    *ncand = p.nsplit * p.kc;
    *nsplit = p.nsplit;
    *group = TC_GROUP;   // a candidate is a group of 8 consecutive packed positions, keyed by its minimum score
    // BF16x3: dropped lo.lo / residual terms are <= ~3 * 2^-16 of |q_i r_i| each; FP32 accumulation on top
    // single FP16: the operand rounding is bounded by the re-rank from the exact residuals (dq16, header.dr16);
    // eps_rel then only covers the FP32 accumulation of the kp products and the two-piece ||r||^2
    *eps_rel = prec ? 1.0e-5 : 6.0e-5;
    *dq16 = dq_w;
    *qrows = qpack;
    *kp = p.kp;
    return SYM_OK;
}

}  // namespace sym

// Test-only: scores of the first query tile against the first reference tile, through the same packing,
// descriptors and MMA chain as the scan.  Q [<=128, d], Ref [<=128, d] -> out [128,128]; scratch >= 3 tiles.
extern "C" int sym_debug_tc_scores(const float *Q, int32_t nq, const float *Ref, int32_t nr, int32_t d, float *out,
                                   void *scratch, sym_stream_t stream, int32_t flags) {
    const int a_in_tmem = flags & 1, prec = (flags >> 1) & 1;   // bit 0: query operand in tensor memory; bit 1: single FP16
    using namespace sym;
    SYM_REQUIRE(nq > 0 && nq <= 128 && nr > 0 && nr <= 128 && d > 0 && d <= 84, SYM_ERR_INVALID_ARGUMENT,
                "sym_debug_tc_scores: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = tc_kp(d, prec);
    unsigned char *qp = (unsigned char *)scratch, *rp = qp + tc_tile_bytes(kp);
    dim3 grid(1, (unsigned)(kp / 8));
    tc_pack_kernel<<<grid, 256, 0, st>>>(Q, d, nq, 128, d, kp, 1, nullptr, qp, prec, nullptr);
    SYM_LAUNCH_CHECK("tc_pack_kernel");
    tc_pack_kernel<<<grid, 256, 0, st>>>(Ref, d, nr, 128, d, kp, 0, nullptr, rp, prec, nullptr);
    SYM_LAUNCH_CHECK("tc_pack_kernel");
    unsigned char *qrows = rp + tc_tile_bytes(kp);
    if (a_in_tmem) {
        tc_pack_rows_kernel<<<16, 256, 0, st>>>(Q, d, nq, 128, d, kp, nullptr, qrows, prec, nullptr, nullptr);
        SYM_LAUNCH_CHECK("tc_pack_rows_kernel");
    }
    const size_t smem = 2 * tc_tile_bytes(kp) + 1024;
    SYM_CUDA(cudaFuncSetAttribute(tc_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_debug_kernel<<<1, 128, smem, st>>>(qp, rp, kp, prec ? TC_IDESC_F16 : TC_IDESC_BF16, out,
                                          a_in_tmem ? qrows : nullptr);
    SYM_LAUNCH_CHECK("tc_debug_kernel");
    return SYM_OK;
}


#ifdef TC_PROFILE
// Profiling builds only: read and clear the cycle counters of the scan (16 x u64, host memory).
extern "C" int sym_debug_tc_profile(unsigned long long *out) {
    SYM_CUDA(cudaDeviceSynchronize());
    SYM_CUDA(cudaMemcpyFromSymbol(out, sym::tc_prof, sizeof(unsigned long long) * 16));
    unsigned long long zero[16] = {0};
    SYM_CUDA(cudaMemcpyToSymbol(sym::tc_prof, zero, sizeof(zero)));
    return SYM_OK;
}
#endif
