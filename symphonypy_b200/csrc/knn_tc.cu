// tcgen05 candidate scan for the brute-force kNN (method 2)      — symphonypy/tl.py:425-433
//
// The distance GEMM  S[q,r] = ||r||^2 - 2 q.r  runs on the 5th-generation tensor cores with BF16x3
// split operands (FP32-class accuracy: each float is hi + lo bf16, and hi.hi + hi.lo + lo.hi are
// summed in the FP32 accumulator), concatenated along K so that ONE chain of tcgen05.mma
// instructions per tile yields the finished score, ||r||^2 included:
//
//     K' index      query operand (A, M=128 rows)      reference operand (B, N=128 rows)
//     [0,d)         hi(-2q)                            hi(r)
//     [d,2d)        hi(-2q)                            lo(r)
//     [2d,3d)       lo(-2q)                            hi(r)
//     3d..3d+2      1                                  the three bf16 pieces of ||r||^2 (exact)
//
// Both operands are packed ahead of time (reference once in sym_knn_fit, queries per call) in the
// canonical K-major no-swizzle UMMA layout, so a tile is one contiguous block moved by a single
// cp.async.bulk into shared memory, and the accumulator never leaves the SM: four 128x128 FP32
// accumulators rotate through the 512 TMEM columns.
//
// Warp roles (192 threads):  warp 0 = bulk-copy producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = selection: thread <-> query row <-> TMEM lane.  A selection thread pulls 32 scores
// with one tcgen05.ld, folds them with 3-input min (16 FMNMX3) against its row's kc-th best,
// and only on a hit walks the 32 values and updates the row's top-kc list in shared memory.
// The scan over-selects (kc = k + 8); the float64 re-rank in knn.cu certifies the margin with the
// scores recorded here and flags rows that need the FP32 scan.
#include "knn_common.cuh"

namespace sym {

constexpr int TC_BM = 128;        // query rows per CTA (UMMA M)
constexpr int TC_BN = 128;        // reference rows per tile (UMMA N)
// query tiles per CTA: QT = 2 (256 rows share every reference tile) when shared memory allows, else 1;
// TMEM accumulator buffers per query tile: 4 / QT  (QT * ACC * 128 = 512 columns)
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_PAD = 8;         // over-selection: kc = k + TC_PAD

__host__ __device__ inline int tc_kp(int d) { return (3 * d + 3 + 15) / 16 * 16; }
__host__ __device__ inline size_t tc_tile_bytes(int kp) { return (size_t)kp * 256; }  // 128 rows x kp bf16

// ------------------------------------------------------------------------------------------ packing
__device__ __forceinline__ unsigned short bf16_rn(float x) {
    unsigned u = __float_as_uint(x);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
__device__ __forceinline__ float bf16_to_f(unsigned short h) { return __uint_as_float((unsigned)h << 16); }

// One thread = one row x one k-unit (8 consecutive K' indices, 16 bytes).
// Element (r, kk) of a tile lives at  (kk/8)*2048 + (r/8)*128 + (r%8)*16 + (kk%8)*2  bytes.
__global__ void __launch_bounds__(256)
tc_pack_kernel(const float *__restrict__ src, int ld, int64_t rows, int64_t rows_pad, int d, int kp, int is_query,
               unsigned char *__restrict__ dst) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;  // k-unit
    if (r >= rows_pad) return;
    unsigned short out[8];
    const bool real = r < rows;
    const float *x = src + r * ld;
    float nrm = 0.f;
    const int k0 = j * 8;
    if (real && !is_query && k0 + 8 > 3 * d && k0 < 3 * d + 3) {
        for (int c = 0; c < d; ++c) nrm = fmaf(x[c], x[c], nrm);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int kk = k0 + e;
        unsigned short v = 0;
        if (kk < 3 * d) {
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

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Busy-poll (mbarrier.test_wait): the hand-offs between the MMA issuer and the selection warps sit on the
// critical path of every tile, and the suspending form (try_wait -> NANOSLEEP.SYNCS) was measured to add
// microseconds per hand-off (profiles/r1d).  Only ~10 warps live on the SM, so polling costs nothing.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Suspending wait for the producer, which runs far ahead and is never on the critical path.
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one lane of a converged warp (elect.sync): keeps the surrounding control flow warp-uniform, so that
// tcgen05.mma / commit operands stay in uniform registers instead of being re-uniformised per instruction
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// K-major, no-swizzle shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
//   start >> 4 | LBO >> 4 at bit 16 (distance between the two 8-element k-units of one MMA)
//   | SBO >> 4 at bit 32 (distance between 8-row groups) | version 1 at bit 46 | layout NONE at bit 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    constexpr uint64_t LBO = 2048 >> 4, SBO = 128 >> 4;
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | (LBO << 16) | (SBO << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=F32 (bit 4), A=B=BF16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);

// ------------------------------------------------------------------------------------------ scan kernel
// A CTA owns TC_QT query tiles (256 rows) and streams the reference tiles once for both.
//   warp 0       : producer (cp.async.bulk of one packed tile per stage)
//   warp 1       : TMEM owner + MMA issuer (TC_QT chains of kp/16 tcgen05.mma per reference tile)
//   warps 2..9   : selection; warp w serves query tile (w-2)/4, TMEM lane quadrant w%4, thread <-> query row
// TMEM: accumulator (query tile q, buffer b) lives at columns (q*TC_ACC + b)*128.
//
// Selection, per 32-column chunk: tcgen05.ld (issued one chunk ahead), four independent 3-input-min
// chains over groups of 8, one compare against the row's kc-th best.  Only groups in which some lane
// of the warp has a hit are walked, and a hit calls the (single, out-of-line) list update.

// Replace the slot holding the row's current maximum `thr` by (s, id); returns the new maximum.  Every lane
// updates ITS OWN row (lanes run in lock-step when several rows hit at once).  Row stride `ks` is odd, so the
// 32 rows of a warp never bank-conflict.  (A warp-cooperative variant — one row at a time, REDUX for the
// maximum — was measured 2x slower end to end: its votes/shuffles serialise on the critical path.)
__device__ __noinline__ float tc_insert(float *sc_row, uint32_t *idx_row, int kc, float s, uint32_t id, float thr) {
    int pos = -1;
    float rest = -INFINITY;  // maximum of the slots that stay
#pragma unroll 8
    for (int e = 0; e < kc; ++e) {
        const float v = sc_row[e];
        const bool take = (pos < 0) && (v == thr);
        pos = take ? e : pos;
        rest = take ? rest : fmaxf(rest, v);
    }
    if (pos < 0) return thr;  // cannot happen: thr is the maximum of the list
    sc_row[pos] = s;
    idx_row[pos] = id;
    return fmaxf(rest, s);
}

// walk one group of 8 scores, only if some lane of the warp has a hit in it
#define TC_WALK8(G)                                                                                   \
    if (__any_sync(0xffffffffu, g##G < thr)) {                                                        \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) {                                               \
            const float sv = __uint_as_float(v[(G) * 8 + j]);                                         \
            if (sv < thr) {                                                                           \
                const int64_t id = ref0 + (G) * 8 + j;                                                \
                if (id < M) thr = tc_insert(sc_row, idx_row, kc, sv, (uint32_t)id, thr);              \
            }                                                                                         \
        }                                                                                             \
    }

__device__ __forceinline__ float tc_min8(const uint32_t *v) {
    float m = fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    m = fmin3(m, __uint_as_float(v[3]), __uint_as_float(v[4]));
    m = fmin3(m, __uint_as_float(v[5]), __uint_as_float(v[6]));
    return fminf(m, __uint_as_float(v[7]));
}

// one 32-column chunk of the warp's 32 rows against the rows' thresholds
__device__ __forceinline__ float tc_select32(const uint32_t (&v)[32], float thr, int64_t ref0, int64_t M, float *sc_row,
                                             uint32_t *idx_row, int kc) {
    const float g0 = tc_min8(v), g1 = tc_min8(v + 8), g2 = tc_min8(v + 16), g3 = tc_min8(v + 24);
    const float m = fminf(fmin3(g0, g1, g2), g3);
#ifndef TC_EXPERIMENT_NO_WALK
    if (__any_sync(0xffffffffu, m < thr)) {
        TC_WALK8(0)
        TC_WALK8(1)
        TC_WALK8(2)
        TC_WALK8(3)
    }
#else
    if (m < -1e30f) thr = m;
#endif
    return thr;
}

template <int TC_QT>
__global__ void __launch_bounds__(64 + 128 * TC_QT, 1)
knn_scan_tc_kernel(const unsigned char *__restrict__ qpack, const unsigned char *__restrict__ rpack, int64_t N,
                   int64_t M, int kp, int kc, int stages, int tiles_per_split, int ntiles_all,
                   unsigned long long *__restrict__ cand_out, float *__restrict__ thr_out) {
    constexpr int TC_ACC = 4 / TC_QT;
    constexpr int TC_SEL_WARPS = 4 * TC_QT;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const uint32_t tile_bytes = (uint32_t)tc_tile_bytes(kp);
    unsigned char *q_s = smraw;                                        // [TC_QT][tile_bytes]
    unsigned char *r_s = q_s + (size_t)TC_QT * tile_bytes;             // [stages][tile_bytes]
    const int ks = kc | 1;  // odd row stride: the 32 rows of a warp hit 32 different banks
    float *sc_s = (float *)(r_s + (size_t)stages * tile_bytes);        // [TC_QT][TC_BM][ks]
    uint32_t *idx_s = (uint32_t *)(sc_s + (size_t)TC_QT * ks * TC_BM); // [TC_QT][TC_BM][ks]
    uint64_t *bars = (uint64_t *)(((uintptr_t)(idx_s + (size_t)TC_QT * ks * TC_BM) + 7) & ~(uintptr_t)7);
    uint64_t *q_full = bars;                      // 1
    uint64_t *r_full = bars + 1;                  // [TC_MAX_STAGES]
    uint64_t *r_empty = r_full + TC_MAX_STAGES;   // [TC_MAX_STAGES]
    uint64_t *t_full = r_empty + TC_MAX_STAGES;   // [4]
    uint64_t *t_empty = t_full + 4;               // [4]
    uint32_t *tmem_slot = (uint32_t *)(t_empty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int t_beg = split * tiles_per_split;
    const int t_end = min(ntiles_all, t_beg + tiles_per_split);
    const int ntiles = max(0, t_end - t_beg);

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < TC_MAX_STAGES; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }  // [query tile][buffer]
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns, owned by this warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {  // selection threads initialise their row's list
        const int qt = (warp - 2) >> 2, row = (warp & 3) * 32 + lane;
        for (int e = 0; e < kc; ++e) {
            sc_s[((size_t)qt * TC_BM + row) * ks + e] = INFINITY;
            idx_s[((size_t)qt * TC_BM + row) * ks + e] = 0xffffffffu;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TC_QT * tile_bytes);
            for (int q = 0; q < TC_QT; ++q)
                bulk_g2s(q_s + (size_t)q * tile_bytes, qpack + ((size_t)blockIdx.x * TC_QT + q) * tile_bytes, tile_bytes,
                         q_full);
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % stages;
                mbar_wait_sleepy(&r_empty[s], ((i / stages) & 1) ^ 1);
                mbar_arrive_expect_tx(&r_full[s], tile_bytes);
                bulk_g2s(r_s + (size_t)s * tile_bytes, rpack + (size_t)(t_beg + i) * tile_bytes, tile_bytes, &r_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop, one elected lane issues =====
        mbar_wait(q_full, 0);
        const uint64_t q_desc = umma_desc(smem_u32(q_s));
        const uint32_t q_step = tile_bytes >> 4;          // descriptor address units (16 B) between query tiles
        const int ksteps = kp / 16;
        for (int i = 0; i < ntiles; ++i) {
            const int s = i % stages, b = i % TC_ACC;
            mbar_wait(&r_full[s], (i / stages) & 1);
            const uint64_t r_desc = umma_desc(smem_u32(r_s + (size_t)s * tile_bytes));
#pragma unroll
            for (int q = 0; q < TC_QT; ++q) {
                mbar_wait(&t_empty[q * TC_ACC + b], ((i / TC_ACC) & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
#ifndef TC_EXPERIMENT_NO_MMA
                    const uint32_t d_tmem = tmem_base + (uint32_t)((q * TC_ACC + b) * TC_BN);
                    const uint64_t a0 = q_desc + (uint64_t)(q * q_step);
#pragma unroll 2
                    for (int k = 0; k < ksteps; ++k)   // one K=16 slice = two k-units = 4096 B = 256 address units
                        tc_mma_bf16(d_tmem, a0 + (uint64_t)(k * 256), r_desc + (uint64_t)(k * 256), TC_IDESC, k > 0);
#endif
                    tc_commit(&t_full[q * TC_ACC + b]);   // this query tile's accumulator is complete
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(&r_empty[s]);   // smem stage reusable once all these MMAs have read it
            __syncwarp();
        }
    } else {
        // ===== selection =====
        const int qt = (warp - 2) >> 2, quad = warp & 3;
        const int row = quad * 32 + lane;
        float *sc_row = sc_s + ((size_t)qt * TC_BM + row) * ks;       // this thread's row
        uint32_t *idx_row = idx_s + ((size_t)qt * TC_BM + row) * ks;
        float thr = INFINITY;
        uint64_t *my_full = t_full + qt * TC_ACC, *my_empty = t_empty + qt * TC_ACC;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(qt * TC_ACC * TC_BN);
        for (int i = 0; i < ntiles; ++i) {
            const int b = i % TC_ACC;
            mbar_wait(&my_full[b], (i / TC_ACC) & 1);
            tc_fence_after();
            const int64_t ref0 = (int64_t)(t_beg + i) * TC_BN;
            const uint32_t acc = lane_base + (uint32_t)(b * TC_BN);
#ifdef TC_EXPERIMENT_NO_SELECT
            tc_fence_before();
            if (lane == 0) mbar_arrive(&my_empty[b]);
            __syncwarp();
            continue;
#endif
            uint32_t va[32], vb[32];
            tc_ld32(acc, va);
            tc_wait_ld();
            tc_ld32(acc + 32, vb);                       // in flight while chunk 0 is examined
            thr = tc_select32(va, thr, ref0, M, sc_row, idx_row, kc);
            __syncwarp();
            tc_wait_ld();
            tc_ld32(acc + 64, va);
            thr = tc_select32(vb, thr, ref0 + 32, M, sc_row, idx_row, kc);
            __syncwarp();
            tc_wait_ld();
            tc_ld32(acc + 96, vb);
            thr = tc_select32(va, thr, ref0 + 64, M, sc_row, idx_row, kc);
            __syncwarp();
            tc_wait_ld();
            // every score of this accumulator is in registers: hand the TMEM buffer back early
            tc_fence_before();
            if (lane == 0) mbar_arrive(&my_empty[b]);
            thr = tc_select32(vb, thr, ref0 + 96, M, sc_row, idx_row, kc);
            __syncwarp();
        }
        // results: the row's candidates as (score, index) keys (unordered) and its final threshold
        const int64_t n = ((int64_t)blockIdx.x * TC_QT + qt) * TC_BM + row;
        if (n < N) {
            for (int e = 0; e < kc; ++e) {
                const uint32_t id = idx_row[e];
                cand_out[(n * nsplit + split) * kc + e] = id == 0xffffffffu ? ~0ull : knn_make_key(sc_row[e], id);
            }
            const int64_t split_refs = min(M, (int64_t)t_end * TC_BN) - (int64_t)t_beg * TC_BN;
            thr_out[n * nsplit + split] = split_refs <= kc ? INFINITY : thr;
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
tc_debug_kernel(const unsigned char *__restrict__ qpack, const unsigned char *__restrict__ rpack, int kp,
                float *__restrict__ out) {
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
                     "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&full, 2 * tile_bytes);
        bulk_g2s(q_s, qpack, tile_bytes, &full);
        bulk_g2s(r_s, rpack, tile_bytes, &full);
        mbar_wait(&full, 0);
        tc_fence_after();
        for (int k = 0; k < kp / 16; ++k)
            tc_mma_bf16(tmem_base, umma_desc(smem_u32(q_s) + k * 4096), umma_desc(smem_u32(r_s) + k * 4096), TC_IDESC,
                        k > 0);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
namespace {
struct TcPlan {
    int kp, kc, qt, stages, nsplit, tiles_per_split, ntiles;
    int64_t qtiles;
    size_t smem, qpack_bytes, cand_bytes, thr_bytes;
    bool ok;
};

TcPlan tc_plan(int64_t N, int64_t M, int d, int k) {
    TcPlan p{};
    p.kp = tc_kp(d);
    int kc = k + TC_PAD;
    if (kc > M) kc = (int)M;
    p.kc = kc;
    const size_t tile = tc_tile_bytes(p.kp);
    const size_t budget = 227 * 1024;
    size_t fixed = 0;
    int stages = 0;
    for (int qt = 2; qt >= 1; --qt) {
        fixed = qt * tile + (size_t)qt * (kc | 1) * TC_BM * 8 + 256 /* barriers */ + 1024 /* alignment slack */;
        stages = budget > fixed ? (int)((budget - fixed) / tile) : 0;
        if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
        p.qt = qt;
        if (stages >= 2) break;
    }
    p.stages = stages;
    p.ok = stages >= 2 && p.kp <= 256 && d <= 84;
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
    p.qpack_bytes = (size_t)p.qtiles * p.qt * tile;
    p.cand_bytes = ((size_t)N * p.nsplit * kc * 8 + 255) & ~(size_t)255;
    p.thr_bytes = ((size_t)N * p.nsplit * 4 + 255) & ~(size_t)255;
    return p;
}
}  // namespace

size_t knn_tc_packed_bytes(int64_t M, int d) {
    if (d > 84) return 0;
    return (size_t)(knn_mpad(M) / TC_BN) * tc_tile_bytes(tc_kp(d));
}

bool knn_tc_supported(int64_t N, int64_t M, int d, int k) {
    if (d > 84 || k > 120) return false;
    return tc_plan(N, M, d, k).ok;
}

size_t knn_tc_workspace(int64_t N, int64_t M, int d, int k) {
    if (!knn_tc_supported(N, M, d, k)) return 0;
    TcPlan p = tc_plan(N, M, d, k);
    return p.qpack_bytes + p.cand_bytes + p.thr_bytes + 1024;
}

int knn_tc_fit(const float *Ref, int ldr, int64_t M, int d, KnnPacked pk, cudaStream_t st) {
    if (d > 84) return SYM_OK;
    const int kp = tc_kp(d);
    const int64_t rows_pad = knn_mpad(M);
    dim3 grid((unsigned)ceil_div(rows_pad, 256), (unsigned)(kp / 8));
    tc_pack_kernel<<<grid, 256, 0, st>>>(Ref, ldr, M, rows_pad, d, kp, 0, pk.tc);
    SYM_LAUNCH_CHECK("tc_pack_kernel(ref)");
    return SYM_OK;
}

int knn_tc_scan(const float *Q, int ldq, int64_t N, KnnPacked pk, int64_t M, int d, int k, void *workspace,
                size_t workspace_bytes, cudaStream_t st, const unsigned long long **cand, const float **thr,
                int *ncand, int *nsplit, double *eps_rel) {
    TcPlan p = tc_plan(N, M, d, k);
    SYM_REQUIRE(p.ok, SYM_ERR_UNSUPPORTED_SHAPE, "tcgen05 scan: d=%d k=%d does not fit", d, k);
    SYM_REQUIRE(workspace_bytes >= p.qpack_bytes + p.cand_bytes + p.thr_bytes, SYM_ERR_WORKSPACE_TOO_SMALL,
                "tcgen05 scan: workspace too small");
    unsigned char *w = (unsigned char *)workspace;
    unsigned char *qpack = w; w += (p.qpack_bytes + 255) & ~(size_t)255;
    unsigned long long *cand_w = (unsigned long long *)w; w += p.cand_bytes;
    float *thr_w = (float *)w;
    {
        const int64_t rows_pad = p.qtiles * p.qt * TC_BM;
        dim3 grid((unsigned)ceil_div(rows_pad, 256), (unsigned)(p.kp / 8));
        tc_pack_kernel<<<grid, 256, 0, st>>>(Q, ldq, N, rows_pad, d, p.kp, 1, qpack);
        SYM_LAUNCH_CHECK("tc_pack_kernel(query)");
    }
    dim3 grid((unsigned)p.qtiles, (unsigned)p.nsplit);
    if (p.qt == 2) {
        SYM_CUDA(cudaFuncSetAttribute(knn_scan_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        knn_scan_tc_kernel<2><<<grid, 64 + 128 * 2, p.smem, st>>>(qpack, pk.tc, N, M, p.kp, p.kc, p.stages,
                                                                  p.tiles_per_split, p.ntiles, cand_w, thr_w);
    } else {
        SYM_CUDA(cudaFuncSetAttribute(knn_scan_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        knn_scan_tc_kernel<1><<<grid, 64 + 128 * 1, p.smem, st>>>(qpack, pk.tc, N, M, p.kp, p.kc, p.stages,
                                                                  p.tiles_per_split, p.ntiles, cand_w, thr_w);
    }
    SYM_LAUNCH_CHECK("knn_scan_tc_kernel");
    *cand = cand_w;
    *thr = thr_w;
    *ncand = p.nsplit * p.kc;
    *nsplit = p.nsplit;
    // BF16x3: dropped lo.lo / residual terms are <= ~3 * 2^-16 of |q_i r_i| each; FP32 accumulation on top
    *eps_rel = 6.0e-5;
    return SYM_OK;
}

}  // namespace sym

// Test-only: scores of the first query tile against the first reference tile, through the same packing,
// descriptors and MMA chain as the scan.  Q [<=128, d], Ref [<=128, d] -> out [128,128]; scratch >= 2 tiles.
extern "C" int sym_debug_tc_scores(const float *Q, int32_t nq, const float *Ref, int32_t nr, int32_t d, float *out,
                                   void *scratch, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(nq > 0 && nq <= 128 && nr > 0 && nr <= 128 && d > 0 && d <= 84, SYM_ERR_INVALID_ARGUMENT,
                "sym_debug_tc_scores: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = tc_kp(d);
    unsigned char *qp = (unsigned char *)scratch, *rp = qp + tc_tile_bytes(kp);
    dim3 grid(1, (unsigned)(kp / 8));
    tc_pack_kernel<<<grid, 256, 0, st>>>(Q, d, nq, 128, d, kp, 1, qp);
    SYM_LAUNCH_CHECK("tc_pack_kernel");
    tc_pack_kernel<<<grid, 256, 0, st>>>(Ref, d, nr, 128, d, kp, 0, rp);
    SYM_LAUNCH_CHECK("tc_pack_kernel");
    const size_t smem = 2 * tc_tile_bytes(kp) + 1024;
    SYM_CUDA(cudaFuncSetAttribute(tc_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_debug_kernel<<<1, 128, smem, st>>>(qp, rp, kp, out);
    SYM_LAUNCH_CHECK("tc_debug_kernel");
    return SYM_OK;
}

