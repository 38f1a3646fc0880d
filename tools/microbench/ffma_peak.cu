// Microbenchmark: sustained FP32 FFMA rate of a register-tile outer product (the inner loop of any CUDA-core GEMM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_peak ffma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int TM, int TN>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int iters) {
    float a[TM], b[TN], acc[TM][TN];
    for (int i = 0; i < TM; ++i) a[i] = in[threadIdx.x + i];
    for (int j = 0; j < TN; ++j) b[j] = in[threadIdx.x + 64 + j];
    for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        // perturb operands so the loop is not hoisted
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] += 1e-7f;
    }
    float s = 0.f;
    for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int TM, int TN> void run(int blocks_per_sm) {
    float *in, *out; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000, grid = 148 * blocks_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<TM, TN><<<grid, 256>>>(out, in, iters);
    cudaEventRecord(e0); k<TM, TN><<<grid, 256>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * TM * TN * (double)iters * grid * 256;
    printf("tile %dx%d, %d blocks/SM: %.1f TFLOP/s\n", TM, TN, blocks_per_sm, flops / ms / 1e9);
}
int main() { run<8, 8>(2); run<8, 8>(4); run<8, 4>(4); run<4, 4>(8); run<16, 4>(2); return 0; }
