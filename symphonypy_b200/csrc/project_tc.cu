// (1b) projection on the tensor cores                             — symphonypy/_utils.py:298-308
//
//   Z = clip((X - mu) * inv_sd, +-clip) @ U          X [N,G] f32 (contiguous genes), U [G,d]
//
// The FFMA kernel (project.cu) spends d FMAs per expression value and is FFMA-bound at d >= 30 (4.25 ms for
// the 8 GB of config2, 29 % of the HBM rate).  Here the GEMM runs on tcgen05 with FP16x3 split operands
// (t = hi + lo, u = hi + lo, each an fp16: 22 significant bits;  hi.hi + hi.lo + lo.hi in the FP32 accumulator,
// FP32-class results — a BF16 split (16 bits) kept Z within 1e-4 but moved the soft assignments R by 3e-5).
// |t| <= clip = 10 fits fp16; the loadings are pre-scaled by a power of two (max |u| into [512, 1024)) so that
// their lo parts stay in the normal fp16 range, and the accumulator is scaled back in the epilogue.  A value
// costs ~8 ALU instructions (scale, clip, split, pack) whatever d is and the kernel is bound by streaming X.
//
//   CTA = 128 cells; genes in chunks of 32 (two K=16 MMA steps), 3-stage ring.
//   warps 0-7  producers: 128-bit streaming loads of X (next chunk prefetched into registers), scale + clip +
//              split in registers, the hi and lo tiles written to shared memory directly in the K-major
//              no-swizzle UMMA layout (one 16-byte vector = 8 consecutive genes of a cell = one core-matrix row);
//              warp 0 also issues the bulk copy of the chunk's pre-packed U tiles (hi | lo, packed once per
//              set of loadings by sym_project_pack)
//   warp 8     allocates tensor memory (two 128 x DPAD FP32 accumulators: main term, cross terms) and issues
//              6 MMAs per chunk
//   warps 0-3  epilogue: tcgen05.ld of the accumulator rows, store Z
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace sym {

constexpr int PT_BM = 128;          // cells per CTA (UMMA M)
constexpr int PT_GK = 32;           // genes per stage
constexpr int PT_STAGES = 3;
#ifndef PT_AHEAD
#define PT_AHEAD 1                  // chunks of X prefetched into registers per producer thread
#endif
constexpr int PT_PRODUCERS = 256;   // producer threads
constexpr int PT_THREADS = PT_PRODUCERS + 32;
constexpr uint32_t PT_A_TILE = PT_BM * PT_GK * 2;   // one 16-bit operand tile of the cells: 8 KB

__host__ __device__ inline int pt_dpad(int d) { return d <= 32 ? 32 : d <= 64 ? 64 : 128; }
__host__ __device__ inline size_t pt_b_tile(int dpad) { return (size_t)dpad * PT_GK * 2; }   // one 16-bit tile of U

// Streaming loads of 8 consecutive floats (one thread's share of a chunk).  One 256-bit load when the row pitch
// allows (32-byte aligned): with two 128-bit loads every 32-byte sector is requested twice, by different
// instructions, and without L1 allocation both requests go to L2 — the kernel was then bound by L2 -> SM sector
// traffic at twice the bytes of X.  The 128-bit fallback therefore allocates in L1.
__device__ __forceinline__ void pt_ldg8(const float *p, bool wide, float4 &a, float4 &b) {
    if (wide) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(p));
    } else {
        a = __ldg(reinterpret_cast<const float4 *>(p));
        b = __ldg(reinterpret_cast<const float4 *>(p + 4));
    }
}
// two floats -> packed fp16 pair (.x = a in the low half), and back
__device__ __forceinline__ uint32_t pt_pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float2 pt_unpack(uint32_t w) { return __half22float2(*reinterpret_cast<__half2 *>(&w)); }
// hi = fp16(t), lo = fp16(t - hi): 22 significant bits between them (FP32-class products hi.hi + hi.lo + lo.hi)
__device__ __forceinline__ void pt_split(float a, float b, uint32_t &hi, uint32_t &lo) {
    hi = pt_pack(a, b);
    const float2 h = pt_unpack(hi);
    lo = pt_pack(a - h.x, b - h.y);
}

// Packed loadings: chunk c of 32 genes -> [hi tile | lo tile]; element (output n, gene k of the chunk) of a tile
// at (k/8)*(DPAD*16) + (n/8)*128 + (n%8)*16 + (k%8)*2 bytes (K-major, no swizzle).  One thread = one 16-byte vector.
__global__ void __launch_bounds__(256)
project_pack_u_kernel(const float *__restrict__ U, int ldu, int G, int d, int dpad, float u_scale,
                      unsigned char *__restrict__ out) {
    const int nchunks = (G + PT_GK - 1) / PT_GK;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nchunks * dpad * 4) return;
    const int c = (int)(idx / (dpad * 4)), rem = (int)(idx % (dpad * 4));
    const int n = rem / 4, ku = rem % 4;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int g = c * PT_GK + ku * 8 + 2 * e + h;
            v[h] = (g < G && n < d) ? U[(int64_t)g * ldu + n] : 0.f;
        }
        pt_split(v[0] * u_scale, v[1] * u_scale, hi[e], lo[e]);
    }
    const size_t bt = pt_b_tile(dpad);
    unsigned char *dst = out + (size_t)c * 2 * bt + (size_t)ku * (dpad * 16) + (n / 8) * 128 + (n % 8) * 16;
    *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(dst + bt) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int DPAD>
__global__ void __launch_bounds__(PT_THREADS, 2)
project_tc_kernel(const float *__restrict__ X, int64_t N, int64_t ldx, int G, const float *__restrict__ mu,
                  const float *__restrict__ inv_sd, const unsigned char *__restrict__ upack, int d, float clip,
                  float out_scale, float *__restrict__ Z, int ldz) {
    constexpr uint32_t B_TILE = DPAD * PT_GK * 2;
    constexpr uint32_t STAGE = 2 * PT_A_TILE + 2 * B_TILE;
    // kind::f16, D = F32 (bit 4), A = B = F16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(DPAD >> 3) << 17) | ((PT_BM >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char psm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(psm + PT_STAGES * STAGE);
    uint64_t *a_full = bars, *b_full = bars + PT_STAGES, *empty = bars + 2 * PT_STAGES, *acc_full = bars + 3 * PT_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * PT_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * PT_BM;
    const int nchunks = (G + PT_GK - 1) / PT_GK;

    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) {
            mbar_init(&a_full[s], PT_PRODUCERS / 32);
            mbar_init(&b_full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(2 * DPAD))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ===== producers: item j = (cell j*64 + 8*warp + lane%8, k-unit lane/8): 8 consecutive genes of one cell =====
        // lane = ku * 8 + (row % 8): the 8 lanes that a 128-bit shared-memory store handles together write the
        // same k-unit of 8 consecutive cells = 128 contiguous bytes (with ku fastest they hit the same banks four
        // at a time: 16 wavefronts per store instead of 4, and the shared-memory pipe became the bottleneck)
        const int ku = lane >> 3;
        const int rl[2] = {warp * 8 + (lane & 7), 64 + warp * 8 + (lane & 7)};
        const bool wide = (ldx % 8) == 0 && (reinterpret_cast<uintptr_t>(X) & 31) == 0;
        float4 nxbuf[PT_AHEAD][2][2];   // the next PT_AHEAD chunks, in flight
        float4 msbuf[PT_AHEAD][4];      // their per-gene offsets and scales (mu x2, inv_sd x2): loaded ahead as well,
                                        // an L2-latency load right before its use was the top stall of this kernel
        auto load = [&](int c, float4 (&nx)[2][2], float4 (&ms)[4]) {
            const int g = c * PT_GK + ku * 8;
            if (g + 8 <= G) {   // g is a multiple of 8: 32-byte aligned
                ms[0] = __ldg(reinterpret_cast<const float4 *>(mu + g));
                ms[1] = __ldg(reinterpret_cast<const float4 *>(mu + g + 4));
                ms[2] = __ldg(reinterpret_cast<const float4 *>(inv_sd + g));
                ms[3] = __ldg(reinterpret_cast<const float4 *>(inv_sd + g + 4));
            } else {
                float t[16];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool ok = g + e < G;
                    t[e] = ok ? __ldg(mu + g + e) : 0.f;
                    t[8 + e] = ok ? __ldg(inv_sd + g + e) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) ms[q] = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int64_t row = row0 + rl[j];
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (row < N) {
                    const float *p = X + row * ldx + g;
                    if (g + 8 <= G) {
                        pt_ldg8(p, wide, a, b);
                    } else {
                        float t[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) t[e] = g + e < G ? p[e] : 0.f;
                        a = make_float4(t[0], t[1], t[2], t[3]);
                        b = make_float4(t[4], t[5], t[6], t[7]);
                    }
                }
                nx[j][0] = a;
                nx[j][1] = b;
            }
        };
#pragma unroll
        for (int a = 0; a < PT_AHEAD; ++a)
            if (a < nchunks) load(a, nxbuf[a], msbuf[a]);
        auto step = [&](int c, float4 (&nx)[2][2], float4 (&ms)[4]) {
            const int s = c % PT_STAGES;
            const uint32_t ph = (uint32_t)((c / PT_STAGES) & 1);
            float x[2][8];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                x[j][0] = nx[j][0].x; x[j][1] = nx[j][0].y; x[j][2] = nx[j][0].z; x[j][3] = nx[j][0].w;
                x[j][4] = nx[j][1].x; x[j][5] = nx[j][1].y; x[j][6] = nx[j][1].z; x[j][7] = nx[j][1].w;
            }
            const float m[8] = {ms[0].x, ms[0].y, ms[0].z, ms[0].w, ms[1].x, ms[1].y, ms[1].z, ms[1].w};
            const float is[8] = {ms[2].x, ms[2].y, ms[2].z, ms[2].w, ms[3].x, ms[3].y, ms[3].z, ms[3].w};
            if (c + PT_AHEAD < nchunks) load(c + PT_AHEAD, nx, ms);
            mbar_wait(&empty[s], ph ^ 1);   // the MMAs that read this stage last time are done
            unsigned char *stage = psm + (size_t)s * STAGE;
            if (tid == 0) {   // this chunk's loadings: [hi | lo] in one bulk copy
                mbar_arrive_expect_tx(&b_full[s], 2 * B_TILE);
                bulk_g2s(stage + 2 * PT_A_TILE, upack + (size_t)c * 2 * B_TILE, 2 * B_TILE, &b_full[s]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * e + h;
                        // genes with inv_sd == 0 (absent from the query or zero std) contribute exactly 0
                        t[h] = is[q] != 0.f ? fminf(fmaxf((x[j][q] - m[q]) * is[q], -clip), clip) : 0.f;
                    }
                    pt_split(t[0], t[1], hi[e], lo[e]);
                }
                unsigned char *dst = stage + (size_t)ku * 2048 + (rl[j] >> 3) * 128 + (rl[j] & 7) * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(dst + PT_A_TILE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
        };
        for (int c0 = 0; c0 < nchunks; c0 += PT_AHEAD) {
#pragma unroll
            for (int a = 0; a < PT_AHEAD; ++a)
                if (c0 + a < nchunks) step(c0 + a, nxbuf[a], msbuf[a]);
        }
        // ===== epilogue: accumulator rows -> Z =====
        if (warp < 4) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int64_t row = row0 + warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c0 = 0; c0 < DPAD; c0 += 32) {
                uint32_t v[32], w[32];
                tc_ld32(taddr + (uint32_t)c0, v);
                tc_ld32(taddr + (uint32_t)(DPAD + c0), w);
                tc_wait_ld();
                if (row < N) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (c0 + e < d) Z[row * ldz + c0 + e] = (__uint_as_float(v[e]) + __uint_as_float(w[e])) * out_scale;
                }
            }
        }
    } else {
        // ===== MMA issuer: per chunk and K step  t_hi.u_hi + t_hi.u_lo + t_lo.u_hi =====
        const bool leader = elect_one();
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % PT_STAGES;
            const uint32_t ph = (uint32_t)((c / PT_STAGES) & 1);
            mbar_wait(&a_full[s], ph);
            mbar_wait(&b_full[s], ph);
            tc_fence_after();
            if (leader) {
                const uint32_t a_hi = smem_u32(psm + (size_t)s * STAGE), a_lo = a_hi + PT_A_TILE;
                const uint32_t b_hi = a_hi + 2 * PT_A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t ah = umma_desc(a_hi + ks * 4096), al = umma_desc(a_lo + ks * 4096);
                    const uint64_t bh = umma_desc(b_hi + ks * 2 * DPAD * 16, DPAD * 16);
                    const uint64_t bl = umma_desc(b_lo + ks * 2 * DPAD * 16, DPAD * 16);
                    // the two cross terms are 2^-12 of the main one: they get their own accumulator, so that the main
                    // accumulator sees a third of the (truncating) tensor-core additions
                    const uint32_t first = (c > 0 || ks > 0) ? 1u : 0u;
                    tc_mma_bf16(tmem_base, ah, bh, IDESC, first);
                    tc_mma_bf16(tmem_base + DPAD, ah, bl, IDESC, first);
                    tc_mma_bf16(tmem_base + DPAD, al, bh, IDESC, 1u);
                }
                tc_commit(&empty[s]);
                if (c == nchunks - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * DPAD))
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------ TMA-staged variant
// Same arithmetic, different way in for X.  The register-prefetch kernel above keeps one chunk of X in flight per
// producer thread: 2 CTAs x 256 threads x 64 B = 32 KB per SM, while HBM at 6.5 TB/s and ~1.5 us of loaded latency
// needs ~66 KB per SM (it sat at 54 % of the copy rate).  Here a dedicated warp lands RAW tiles of X — 128 cells x
// 32 genes of float32, 16 KB — in shared memory with cp.async.bulk.tensor (one instruction per tile, 128-byte
// swizzle) through a ring of RAW stages: 2 CTAs x 4 stages x 16 KB = 128 KB in flight per SM, independent of what
// the transform threads are doing.  The 8 transform warps read a raw tile conflict-free (the swizzle XORs the
// 16-byte unit of a row with row % 8: the 8 cells that share a k-unit hit 8 different bank groups), scale + clip +
// split in registers and write the hi / lo operand tiles in the UMMA layout exactly as above.
constexpr int PT_OP_STAGES = 2;
__host__ __device__ inline int pt_raw_stages(int dpad) { return dpad == 32 ? 4 : 3; }
constexpr uint32_t PT_RAW_TILE = PT_BM * PT_GK * 4;   // 16 KB
constexpr int PT_TMA_THREADS = PT_PRODUCERS + 64;      // + MMA issuer warp + TMA warp

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int DPAD>
__global__ void __launch_bounds__(PT_TMA_THREADS, 2)
project_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t N, int G, const float *__restrict__ mu,
                   const float *__restrict__ inv_sd, const unsigned char *__restrict__ upack, int d, float clip,
                   float out_scale, float *__restrict__ Z, int ldz) {
    constexpr uint32_t B_TILE = DPAD * PT_GK * 2;
    constexpr uint32_t OP_STAGE = 2 * PT_A_TILE + 2 * B_TILE;
    constexpr int RAW = DPAD == 32 ? 4 : 3;
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(DPAD >> 3) << 17) | ((PT_BM >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char psm[];
    unsigned char *raw_s = psm;                                   // [RAW][16 KB], 1024-byte aligned (128-byte swizzle)
    unsigned char *op_s = psm + RAW * PT_RAW_TILE;                // [PT_OP_STAGES][A hi | A lo | U hi | U lo]
    uint64_t *bars = reinterpret_cast<uint64_t *>(op_s + PT_OP_STAGES * OP_STAGE);
    uint64_t *raw_full = bars, *raw_empty = bars + RAW, *a_full = bars + 2 * RAW, *b_full = a_full + PT_OP_STAGES;
    uint64_t *op_empty = b_full + PT_OP_STAGES, *acc_full = op_empty + PT_OP_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * PT_BM;
    const int nchunks = (G + PT_GK - 1) / PT_GK;

    if (tid == 0) {
        for (int s = 0; s < RAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], PT_PRODUCERS / 32); }
        for (int s = 0; s < PT_OP_STAGES; ++s) {
            mbar_init(&a_full[s], PT_PRODUCERS / 32);
            mbar_init(&b_full[s], 1);
            mbar_init(&op_empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(2 * DPAD))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 9) {
        // ===== TMA warp: one raw tile of X per chunk =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait_sleepy(&raw_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&raw_full[s], PT_RAW_TILE);
                tma_load_2d(raw_s + (size_t)s * PT_RAW_TILE, &tmap, c * PT_GK, (int)row0, &raw_full[s]);
                if (++s == RAW) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 8) {
        // ===== MMA issuer: per chunk and K step  t_hi.u_hi + t_hi.u_lo + t_lo.u_hi =====
        const bool leader = elect_one();
        for (int c = 0; c < nchunks; ++c) {
            const int t = c % PT_OP_STAGES;
            const uint32_t ph = (uint32_t)((c / PT_OP_STAGES) & 1);
            mbar_wait(&a_full[t], ph);
            mbar_wait(&b_full[t], ph);
            tc_fence_after();
            if (leader) {
                const uint32_t a_hi = smem_u32(op_s + (size_t)t * OP_STAGE), a_lo = a_hi + PT_A_TILE;
                const uint32_t b_hi = a_hi + 2 * PT_A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t ah = umma_desc(a_hi + ks * 4096), al = umma_desc(a_lo + ks * 4096);
                    const uint64_t bh = umma_desc(b_hi + ks * 2 * DPAD * 16, DPAD * 16);
                    const uint64_t bl = umma_desc(b_lo + ks * 2 * DPAD * 16, DPAD * 16);
                    const uint32_t first = (c > 0 || ks > 0) ? 1u : 0u;
                    tc_mma_bf16(tmem_base, ah, bh, IDESC, first);           // main term
                    tc_mma_bf16(tmem_base + DPAD, ah, bl, IDESC, first);    // cross terms: their own accumulator
                    tc_mma_bf16(tmem_base + DPAD, al, bh, IDESC, 1u);
                }
                tc_commit(&op_empty[t]);
                if (c == nchunks - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ===== transform warps: item j = (cell j*64 + 8*warp + lane%8, k-unit lane/8) =====
        const int ku = lane >> 3;
        const int rl[2] = {warp * 8 + (lane & 7), 64 + warp * 8 + (lane & 7)};
        // raw tile: row r at r*128 bytes, its 16-byte unit u stored at unit (u ^ (r % 8)); rl[j] % 8 == lane % 8
        const uint32_t sw = (uint32_t)(lane & 7);
        const uint32_t off0 = (((uint32_t)(2 * ku)) ^ sw) << 4, off1 = (((uint32_t)(2 * ku + 1)) ^ sw) << 4;
        float4 ms[4];   // mu x2, inv_sd x2 of this thread's 8 genes, one chunk ahead
        auto load_ms = [&](int c) {
            const int g = c * PT_GK + ku * 8;
            if (g + 8 <= G) {
                ms[0] = __ldg(reinterpret_cast<const float4 *>(mu + g));
                ms[1] = __ldg(reinterpret_cast<const float4 *>(mu + g + 4));
                ms[2] = __ldg(reinterpret_cast<const float4 *>(inv_sd + g));
                ms[3] = __ldg(reinterpret_cast<const float4 *>(inv_sd + g + 4));
            } else {   // the tail chunk: genes past G read as x = 0 (TMA zero fill) and get inv_sd = 0
                float t[16];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool ok = g + e < G;
                    t[e] = ok ? __ldg(mu + g + e) : 0.f;
                    t[8 + e] = ok ? __ldg(inv_sd + g + e) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) ms[q] = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
            }
        };
        load_ms(0);
        int s = 0, t = 0;
        uint32_t ph_raw = 0, ph_op = 0;
        for (int c = 0; c < nchunks; ++c) {
            // t = clip(x * inv_sd - mu * inv_sd): one FMA per value.  Genes with inv_sd == 0 (zero std; absent genes cannot
            // occur on this path, the genes are in reference order) give x * 0 - 0 = 0 exactly for every finite x.
            const float is[8] = {ms[2].x, ms[2].y, ms[2].z, ms[2].w, ms[3].x, ms[3].y, ms[3].z, ms[3].w};
            const float m[8] = {-ms[0].x * is[0], -ms[0].y * is[1], -ms[0].z * is[2], -ms[0].w * is[3],
                                -ms[1].x * is[4], -ms[1].y * is[5], -ms[1].z * is[6], -ms[1].w * is[7]};
            if (c + 1 < nchunks) load_ms(c + 1);
            mbar_wait(&raw_full[s], ph_raw);
            const unsigned char *raw = raw_s + (size_t)s * PT_RAW_TILE;
            float4 xv[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                xv[j][0] = *reinterpret_cast<const float4 *>(raw + rl[j] * 128 + off0);
                xv[j][1] = *reinterpret_cast<const float4 *>(raw + rl[j] * 128 + off1);
            }
            uint32_t hi[2][4], lo[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float x[8] = {xv[j][0].x, xv[j][0].y, xv[j][0].z, xv[j][0].w, xv[j][1].x, xv[j][1].y, xv[j][1].z, xv[j][1].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float tt[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * e + h;
                        tt[h] = fminf(fmaxf(fmaf(x[q], is[q], m[q]), -clip), clip);
                    }
                    pt_split(tt[0], tt[1], hi[j][e], lo[j][e]);
                }
            }
            // Release the raw stage only now, after the values have been USED: the loads above are merely issued when
            // the next instruction runs, and the TMA write that refills the stage goes through the async proxy, which is
            // not ordered behind generic-proxy loads still in flight.  Arriving right after the loads (with or without a
            // membar.cta) let ~0.3 % of the tiles read one 32-byte unit of the NEXT use of the stage: one cell per ~300
            // tiles off by a few per cent, differently on every launch (found by repeated launches on the same input —
            // `tests: test_project_tensor_core_large_launches`).  Either measure fixes it alone: the arrive after the uses, or
            // a generic -> async proxy fence before it; both are kept (the compiler may sink register arithmetic below
            // the barrier instructions, the fence does not depend on instruction placement).
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[s]);
            mbar_wait(&op_empty[t], ph_op ^ 1);   // the MMAs that read this operand stage last time are done
            unsigned char *stage = op_s + (size_t)t * OP_STAGE;
            if (tid == 0) {   // this chunk's loadings: [hi | lo] in one bulk copy
                mbar_arrive_expect_tx(&b_full[t], 2 * B_TILE);
                bulk_g2s(stage + 2 * PT_A_TILE, upack + (size_t)c * 2 * B_TILE, 2 * B_TILE, &b_full[t]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                unsigned char *dst = stage + (size_t)ku * 2048 + (rl[j] >> 3) * 128 + (rl[j] & 7) * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[j][0], hi[j][1], hi[j][2], hi[j][3]);
                *reinterpret_cast<uint4 *>(dst + PT_A_TILE) = make_uint4(lo[j][0], lo[j][1], lo[j][2], lo[j][3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[t]);
            if (++s == RAW) { s = 0; ph_raw ^= 1; }
            if (++t == PT_OP_STAGES) { t = 0; ph_op ^= 1; }
        }
        // ===== epilogue: accumulator rows -> Z =====
        if (warp < 4) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int64_t row = row0 + warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c0 = 0; c0 < DPAD; c0 += 32) {
                uint32_t v[32], w[32];
                tc_ld32(taddr + (uint32_t)c0, v);
                tc_ld32(taddr + (uint32_t)(DPAD + c0), w);
                tc_wait_ld();
                if (row < N) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (c0 + e < d) Z[row * ldz + c0 + e] = (__uint_as_float(v[e]) + __uint_as_float(w[e])) * out_scale;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * DPAD))
                     : "memory");
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*pt_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pt_encode_fn pt_encoder() {
    static pt_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (pt_encode_fn)p;
    }
    return fn;
}

// 1 = launched, 0 = not applicable (no encoder / layout the tensor map cannot describe): the caller falls back
template <int DPAD>
static int launch_project_tma(const float *X, int64_t N, int64_t ldx, int G, const float *mu, const float *inv_sd,
                              const unsigned char *upack, int d, float clip, float out_scale, float *Z, int ldz,
                              cudaStream_t st, int *launched) {
    *launched = 0;
    pt_encode_fn enc = pt_encoder();
    if (enc == nullptr || (ldx % 4) != 0 || ((uintptr_t)X & 15) != 0 || N >= (1ll << 31)) return SYM_OK;
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)G, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * 4};
    const cuuint32_t box[2] = {(cuuint32_t)PT_GK, (cuuint32_t)PT_BM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return SYM_OK;
    const size_t smem = (size_t)pt_raw_stages(DPAD) * PT_RAW_TILE + (size_t)PT_OP_STAGES * (2 * PT_A_TILE + 2 * pt_b_tile(DPAD)) + 256;
    SYM_CUDA(cudaFuncSetAttribute(project_tma_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    project_tma_kernel<DPAD><<<(unsigned)ceil_div(N, PT_BM), PT_TMA_THREADS, smem, st>>>(tmap, N, G, mu, inv_sd, upack, d, clip,
                                                                                        out_scale, Z, ldz);
    SYM_LAUNCH_CHECK("project_tma_kernel");
    *launched = 1;
    return SYM_OK;
}

template <int DPAD>
static int launch_project_tc(const float *X, int64_t N, int64_t ldx, int G, const float *mu, const float *inv_sd,
                             const unsigned char *upack, int d, float clip, float out_scale, float *Z, int ldz,
                             cudaStream_t st) {
    const size_t smem = (size_t)PT_STAGES * (2 * PT_A_TILE + 2 * pt_b_tile(DPAD)) + 256;
    SYM_CUDA(cudaFuncSetAttribute(project_tc_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    project_tc_kernel<DPAD><<<(unsigned)ceil_div(N, PT_BM), PT_THREADS, smem, st>>>(X, N, ldx, G, mu, inv_sd, upack, d,
                                                                                   clip, out_scale, Z, ldz);
    SYM_LAUNCH_CHECK("project_tc_kernel");
    return SYM_OK;
}

}  // namespace sym

extern "C" size_t sym_project_pack_bytes(int32_t G, int32_t d) {
    if (G <= 0 || d <= 0 || d > 128) return 0;
    return (size_t)sym::ceil_div(G, sym::PT_GK) * 2 * sym::pt_b_tile(sym::pt_dpad(d));
}

extern "C" int sym_project_pack(const float *U, int32_t ldu, int32_t G, int32_t d, float u_scale, void *packed,
                                sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(G > 0 && d > 0 && d <= 128 && ldu >= d, SYM_ERR_INVALID_ARGUMENT, "sym_project_pack: bad sizes");
    SYM_REQUIRE(((uintptr_t)packed & 15) == 0, SYM_ERR_INVALID_ARGUMENT, "sym_project_pack: packed must be 16-byte aligned");
    const int dpad = pt_dpad(d);
    const int64_t items = (int64_t)ceil_div(G, PT_GK) * dpad * 4;
    project_pack_u_kernel<<<(unsigned)ceil_div(items, 256), 256, 0, (cudaStream_t)stream>>>(U, ldu, G, d, dpad,
                                                                                           u_scale,
                                                                                           (unsigned char *)packed);
    SYM_LAUNCH_CHECK("project_pack_u_kernel");
    return SYM_OK;
}

extern "C" int sym_project_dense_tc(const float *X, int64_t N, int64_t ldx, int32_t G, const float *mu,
                                    const float *inv_sd, const void *packed, int32_t d, float clip, float out_scale,
                                    float *Z, int32_t ldz, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && G > 0 && d > 0 && d <= 128 && ldx >= G && ldz >= d, SYM_ERR_INVALID_ARGUMENT,
                "sym_project_dense_tc: bad sizes");
    SYM_REQUIRE((ldx % 4) == 0 && ((uintptr_t)X & 15) == 0, SYM_ERR_INVALID_ARGUMENT,
                "sym_project_dense_tc: X rows must be 16-byte aligned (use sym_project_dense otherwise)");
    SYM_REQUIRE(((uintptr_t)packed & 15) == 0, SYM_ERR_INVALID_ARGUMENT, "sym_project_dense_tc: packed must be 16-byte aligned");
    if (N == 0) return SYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned char *up = (const unsigned char *)packed;
    const int dpad = pt_dpad(d);
    const char *no_tma = getenv("SYM_PROJECT_TMA");   // "0": the register-prefetch kernel (experiments / comparison)
    if (no_tma == nullptr || atoi(no_tma) != 0) {
        int launched = 0, rc;
        if (dpad == 32) rc = launch_project_tma<32>(X, N, ldx, G, mu, inv_sd, up, d, clip, out_scale, Z, ldz, st, &launched);
        else if (dpad == 64) rc = launch_project_tma<64>(X, N, ldx, G, mu, inv_sd, up, d, clip, out_scale, Z, ldz, st, &launched);
        else rc = launch_project_tma<128>(X, N, ldx, G, mu, inv_sd, up, d, clip, out_scale, Z, ldz, st, &launched);
        if (rc != SYM_OK || launched) return rc;
    }
    if (dpad == 32) return launch_project_tc<32>(X, N, ldx, G, mu, inv_sd, up, d, clip, out_scale, Z, ldz, st);
    if (dpad == 64) return launch_project_tc<64>(X, N, ldx, G, mu, inv_sd, up, d, clip, out_scale, Z, ldz, st);
    return launch_project_tc<128>(X, N, ldx, G, mu, inv_sd, up, d, clip, out_scale, Z, ldz, st);
}
