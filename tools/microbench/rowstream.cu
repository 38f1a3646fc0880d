// How fast can 128-row CTAs stream a row-major [N, G] f32 matrix when every CTA walks its rows in steps of W bytes
// per row?  (The projection reads X this way: W = 128.)  nvcc -arch=sm_100a -O3 rowstream.cu -o rowstream
#include <cstdio>
#include <cuda_runtime.h>

template <int W, int AHEAD>   // W = floats per row per step (32, 64, 128); thread = 8 floats (32 B)
__global__ void __launch_bounds__(256) stream_kernel(const float *X, long N, long ldx, int G, float *out) {
    constexpr int TPR = W / 8;              // threads per row per step
    constexpr int ROWS_PER_PASS = 256 / TPR;
    constexpr int PASSES = 128 / ROWS_PER_PASS;
    const int tid = threadIdx.x;
    const long row0 = (long)blockIdx.x * 128;
    float acc = 0.f;
    float4 buf[AHEAD][PASSES][2];
    auto load = [&](int c, float4 (&b)[PASSES][2]) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const long row = row0 + p * ROWS_PER_PASS + tid / TPR;
            const float *q = X + row * ldx + c * W + (tid % TPR) * 8;
            asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(b[p][0].x), "=f"(b[p][0].y), "=f"(b[p][0].z), "=f"(b[p][0].w), "=f"(b[p][1].x),
                           "=f"(b[p][1].y), "=f"(b[p][1].z), "=f"(b[p][1].w)
                         : "l"(q));
        }
    };
    const int nch = G / W;
#pragma unroll
    for (int a = 0; a < AHEAD; ++a) load(a, buf[a]);
    for (int c0 = 0; c0 < nch; c0 += AHEAD) {
#pragma unroll
        for (int a = 0; a < AHEAD; ++a) {
            const int c = c0 + a;
            if (c < nch) {
#pragma unroll
                for (int p = 0; p < PASSES; ++p)
                    acc += buf[a][p][0].x + buf[a][p][0].w + buf[a][p][1].y + buf[a][p][1].z;
                if (c + AHEAD < nch) load(c + AHEAD, buf[a]);
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int W, int AHEAD>
void run(const float *X, long N, long ldx, int G, float *out, const char *name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    stream_kernel<W, AHEAD><<<(unsigned)(N / 128), 256>>>(X, N, ldx, G, out);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) stream_kernel<W, AHEAD><<<(unsigned)(N / 128), 256>>>(X, N, ldx, G, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("%-28s %.3f ms  %.2f TB/s  (%s)\n", name, ms, (double)N * G * 4 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long N = 1000064 / 128 * 128, G = 2048, ldx = 2048;
    float *X, *out;
    cudaMalloc(&X, N * ldx * 4);
    cudaMalloc(&out, 4);
    cudaMemset(X, 0, N * ldx * 4);
    run<32, 1>(X, N, ldx, G, out, "W=128B ahead 1 (2 CTA regs)");
    run<32, 2>(X, N, ldx, G, out, "W=128B ahead 2");
    run<32, 4>(X, N, ldx, G, out, "W=128B ahead 4");
    run<64, 1>(X, N, ldx, G, out, "W=256B ahead 1");
    run<64, 2>(X, N, ldx, G, out, "W=256B ahead 2");
    run<128, 1>(X, N, ldx, G, out, "W=512B ahead 1");
    run<128, 2>(X, N, ldx, G, out, "W=512B ahead 2");
    return 0;
}
