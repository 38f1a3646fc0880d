// Microbenchmark: tcgen05.ld throughput / latency per SM as a function of warps and loads in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int INFLIGHT>
__global__ void bench(int iters, long long *cycles, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[INFLIGHT][32];
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f) ld32(base + ((it * INFLIGHT + f) * 32) % 512, v[f]);
        wait_ld();
        // light consumption (two of the 32 registers per load): the ALU pipe must not be what is measured
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f) acc ^= v[f][0] ^ v[f][31];
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cycles[blockIdx.x * (blockDim.x / 32) + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int INFLIGHT>
void run(int warps) {
    const int iters = 2000, grid = 148;
    long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, grid * warps * 8); cudaMalloc(&sink, grid * warps * 32 * 4);
    bench<INFLIGHT><<<grid, warps * 32>>>(iters, cyc, sink);
    bench<INFLIGHT><<<grid, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, cyc, warps * 8, cudaMemcpyDeviceToHost);
    double c = (double)h[0] / (iters * INFLIGHT);
    printf("warps=%2d inflight=%d: %.1f cycles per LDTM.x32 per warp -> %.1f B/cycle/SM  (%s)\n", warps, INFLIGHT, c,
           warps * 4096.0 / c, cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int w : {1, 4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); }
    return 0;
}
