// Microbenchmark: tcgen05.mma (kind::f16, bf16, M=128) issue/execute rate from shared memory operands in the
// K-major no-swizzle layout used by knn_tc.cu, for N = 64 / 128 / 256 and for 1..4 accumulators in rotation.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo) {
    return (uint64_t)((a & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// A operand from tensor memory (TS form)
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool TS>
__global__ void bench(int tiles, int kp, int nacc, long long *cycles, const unsigned char *gsrc, int copy_bytes, int mma_on, int commit_every) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint64_t cbar[4];
    __shared__ uint64_t dummy_bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + N) * kp * 2 / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((128 >> 4) << 24);
    if (warp == 1 && copy_bytes > 0 && threadIdx.x == 32) {
        // concurrent producer: `tiles` bulk copies of copy_bytes each, 4 in flight, into smem beyond the operands
        unsigned char *dst0 = sm + (size_t)(128 + N) * kp * 2;
        long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const int s = t & 3;
            if (t >= 4) asm volatile("{\n\t.reg .pred P1;\n\tWC:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DC;\n\tbra WC;\n\tDC:\n\t}" ::"r"(smem_u32(&cbar[s])), "r"((uint32_t)(((t >> 2) - 1) & 1)) : "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"((uint32_t)copy_bytes), "r"(smem_u32(&cbar[s])) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst0 + (size_t)s * copy_bytes)), "l"(gsrc + ((size_t)(blockIdx.x * 64 + (t & 63)) * copy_bytes)), "r"((uint32_t)copy_bytes), "r"(smem_u32(&cbar[s])) : "memory");
        }
        for (int s = 0; s < 4; ++s) { int t = tiles - 4 + s; asm volatile("{\n\t.reg .pred P1;\n\tWE:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DE;\n\tbra WE;\n\tDE:\n\t}" ::"r"(smem_u32(&cbar[t & 3])), "r"((uint32_t)((t >> 2) & 1)) : "memory"); }
        cycles[148 + blockIdx.x] = clock64() - t0;
    }
    if (warp == 0 && mma_on) {
        const uint64_t a0 = umma_desc(smem_u32(sm), 128 * 16), b0 = umma_desc(smem_u32(sm) + 128 * kp * 2, N * 16);
        const uint32_t bstep = (2 * N * 16) >> 4;
        const int ks = kp / 16;
        long long t0 = clock64();
        int acc = 0;
        for (int t = 0; t < tiles; ++t) {
            if (elect_one()) {
                // accumulators occupy the top of TMEM; in TS mode the A operand sits in columns [0, kp/2)
                const uint32_t d = tmem + (uint32_t)(512 - (acc + 1) * N);
                for (int k = 0; k < ks; ++k) {
                    if (TS) mma_ts(d, tmem + (uint32_t)(k * 8), b0 + (uint64_t)(k * bstep), idesc, k > 0);
                    else mma(d, a0 + (uint64_t)(k * 256), b0 + (uint64_t)(k * bstep), idesc, k > 0);
                }
                if (commit_every && (t % commit_every) == commit_every - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy_bar)) : "memory");
            }
            __syncwarp();
            acc = acc + 1 == nacc ? 0 : acc + 1;
        }
        if (elect_one()) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool TS>
void run(int kp, int nacc, int copy_bytes = 0, int mma_on = 1, int commit_every = 0) {
    const int tiles = 2000, grid = 148;
    long long *cyc; cudaMalloc(&cyc, 2 * grid * 8); cudaMemset(cyc, 0, 2 * grid * 8);
    unsigned char *gsrc; cudaMalloc(&gsrc, (size_t)148 * 64 * 49152); cudaMemset(gsrc, 0, (size_t)148 * 64 * 49152);
    size_t smem = (size_t)(128 + N) * kp * 2 + 1024 + 4 * (size_t)copy_bytes;
    cudaFuncSetAttribute(bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<N, TS><<<grid, 128, smem>>>(tiles, kp, nacc, cyc, gsrc, copy_bytes, mma_on, commit_every);
    bench<N, TS><<<grid, 128, smem>>>(tiles, kp, nacc, cyc, gsrc, copy_bytes, mma_on, commit_every);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, cyc, 2 * grid * 8, cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (tiles * (kp / 16));
    printf("[commit every %d tiles] %s N=%3d kp=%3d nacc=%d: %.1f cycles per MMA (ideal %d)  [%s]\n", commit_every, TS ? "TS" : "SS", N, kp, nacc, per, N / 2, cudaGetErrorString(e));
    if (copy_bytes) printf("      + concurrent bulk copies of %d B: %.0f cycles per copy = %.1f B/clk/SM (mma_on=%d)\n", copy_bytes, (double)h[148] / tiles, copy_bytes * (double)tiles / h[148], mma_on);
    cudaFree(cyc); cudaFree(gsrc);
}

int main() {
    run<128, true>(96, 2, 0, 1, 0);
    run<128, true>(96, 2, 0, 1, 1);
    run<128, true>(96, 2, 0, 1, 2);
    run<128, true>(96, 2, 0, 1, 4);
    run<128, true>(96, 3, 0, 1, 1);
    run<128, true>(96, 4, 0, 1, 1);
    run<128, false>(96, 2, 0, 1, 1);
    return 0;
}
