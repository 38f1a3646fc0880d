// per-cell mapping confidence              — symphonypy/tl.py:33-111, _utils.py:468-496
//
//   reference side (once per reference, cached):  R-weighted mean and covariance of every Harmony cluster
//       v1_k = sum_m R_km, v2_k = sum_m R_km^2, mean_k = sum_m R_km x_m / v1_k,
//       cov_k = sum_m R_km (x_m - mean_k)(x_m - mean_k)^T * v1_k / (v1_k^2 - v2_k)
//   query side:  out_n = sum_k R_nk * sqrt( (x_n - mean_k)^T inv(cov_k) (x_n - mean_k) )
//
// The reference materialises [K, d, N_ref] and [K, N_query, d] float64 temporaries (12 GB / 24 GB at the
// BASELINE shapes); here the sums are streamed, FP32 partials with FP64 atomics on the reference side, and the
// query side is one pass over the cells with the K inverse covariances staged through shared memory.
#include "common.cuh"
#include "work_chunks.cuh"

namespace sym {

constexpr int CC_THREADS = 256;
constexpr int CC_CHUNK = 2048;  // reference cells per block

// sums[k] = {v1, v2, sum_m R_km x_m[0..d)}   (double, [K][d+2])
__global__ void __launch_bounds__(CC_THREADS)
cluster_moments_kernel(const float *__restrict__ X, int ldx, int64_t M, int d, const float *__restrict__ R /*[K][M]*/,
                       int K, double *__restrict__ sums) {
    const int k = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * CC_CHUNK;
    const int64_t m1 = m0 + CC_CHUNK < M ? m0 + CC_CHUNK : M;
    extern __shared__ __align__(16) float csm[];   // [CC_THREADS/32][d+2] warp partials
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // lane <-> column (strided), warps stride over cells: coalesced reads of X rows
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float a1 = 0.f, a2 = 0.f;
    const float *Rk = R + (int64_t)k * M;
    for (int64_t m = m0 + warp; m < m1; m += CC_THREADS / 32) {
        const float r = Rk[m];
        a1 += r;
        a2 = fmaf(r, r, a2);
        const float *x = X + m * ldx;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < d) acc[q] = fmaf(r, x[lane + 32 * q], acc[q]);
    }
    float *wp = csm + warp * (d + 2);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < d) wp[2 + lane + 32 * q] = acc[q];
    if (lane == 0) { wp[0] = a1; wp[1] = a2; }
    __syncthreads();
    for (int c = threadIdx.x; c < d + 2; c += CC_THREADS) {
        double s = 0.0;
        for (int w = 0; w < CC_THREADS / 32; ++w) s += (double)csm[w * (d + 2) + c];
        atomicAdd(&sums[(int64_t)k * (d + 2) + c], s);
    }
}

// cov[k][i][j] += sum_{m in chunk} R_km (x_mi - mean_ki)(x_mj - mean_kj)
__global__ void __launch_bounds__(CC_THREADS)
cluster_cov_kernel(const float *__restrict__ X, int ldx, int64_t M, int d, const float *__restrict__ R, int K,
                   const float *__restrict__ mean /*[K][d]*/, double *__restrict__ cov /*[K][d][d]*/) {
    const int k = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * CC_CHUNK;
    const int64_t m1 = m0 + CC_CHUNK < M ? m0 + CC_CHUNK : M;
    extern __shared__ __align__(16) float csm[];
    float *xs = csm;             // [32][d+1] centred cells
    float *rs = csm + 32 * (d + 1);  // [32]
    const int tid = threadIdx.x;
    const int nent = d * d;
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    const float *Rk = R + (int64_t)k * M;
    const float *mu = mean + (int64_t)k * d;
    for (int64_t s0 = m0; s0 < m1; s0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * d; i += CC_THREADS) {
            const int c = i / d, j = i % d;
            xs[c * (d + 1) + j] = (s0 + c < m1) ? X[(s0 + c) * ldx + j] - mu[j] : 0.f;
        }
        if (tid < 32) rs[tid] = (s0 + tid < m1) ? Rk[s0 + tid] : 0.f;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int ent = tid + e * CC_THREADS;
            if (ent < nent) {
                const int i = ent / d, j = ent % d;
                float a = acc[e];
#pragma unroll 8
                for (int c = 0; c < 32; ++c) a = fmaf(rs[c] * xs[c * (d + 1) + i], xs[c * (d + 1) + j], a);
                acc[e] = a;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int ent = tid + e * CC_THREADS;
        if (ent < nent) atomicAdd(&cov[(int64_t)k * nent + ent], (double)acc[e]);
    }
}

// out[n] = sum_k R[n,k] sqrt(xc^T S_k xc), xc = x_n - mean_k.   One thread per cell, x in registers.
template <int DPAD, int TPB>
__global__ void __launch_bounds__(TPB)
cell_confidence_kernel(const float *__restrict__ Xq, int ldq, int64_t N, int d, const float *__restrict__ Rq, int K,
                       const float *__restrict__ mean, const float *__restrict__ inv_cov, float *__restrict__ out) {
    __shared__ __align__(16) float S[DPAD * DPAD];
    __shared__ float mu[DPAD];
    __shared__ float xcs[DPAD * TPB];   // centred cell, [i][thread]: lets the outer loop index it at run time
    const int64_t n = (int64_t)blockIdx.x * TPB + threadIdx.x;
    float x[DPAD];
#pragma unroll
    for (int i = 0; i < DPAD; ++i) x[i] = (n < N && i < d) ? Xq[n * ldq + i] : 0.f;
    float total = 0.f;
    for (int k = 0; k < K; ++k) {
        __syncthreads();
        for (int i = threadIdx.x; i < DPAD * DPAD; i += TPB) {
            const int r = i / DPAD, c = i % DPAD;
            S[i] = (r < d && c < d) ? inv_cov[((int64_t)k * d + r) * d + c] : 0.f;
        }
        if (threadIdx.x < DPAD) mu[threadIdx.x] = threadIdx.x < d ? mean[(int64_t)k * d + threadIdx.x] : 0.f;
        __syncthreads();
        float xc[DPAD];
#pragma unroll
        for (int i = 0; i < DPAD; ++i) {
            xc[i] = x[i] - mu[i];
            xcs[i * TPB + threadIdx.x] = xc[i];
        }
        float q = 0.f;
#pragma unroll 4
        for (int i = 0; i < DPAD; ++i) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < DPAD; j += 4) {
                const float4 s4 = *reinterpret_cast<const float4 *>(S + i * DPAD + j);
                t = fmaf(s4.x, xc[j], t);
                t = fmaf(s4.y, xc[j + 1], t);
                t = fmaf(s4.z, xc[j + 2], t);
                t = fmaf(s4.w, xc[j + 3], t);
            }
            q = fmaf(xcs[i * TPB + threadIdx.x], t, q);   // own column: no synchronisation needed
        }
        if (n < N) total = fmaf(Rq[n * K + k], sqrtf(fmaxf(q, 0.f)), total);
    }
    if (n < N) out[n] = total;
}

// ------------------------------------------------------------------------------------------
// per-cluster confidence                    — symphonypy/_utils.py:422-465
// Query cells grouped by a user-given cluster code (sym_batch_order): per group the column sums and the centred
// scatter matrix sum (x - mean)(x - mean)^T; one chunk of <= 256 members of ONE group per block, FP32 partials,
// FP64 atomics.  The reference slices the AnnData per group and runs numpy on each slice.
constexpr int GC_THREADS = 256;

__global__ void __launch_bounds__(GC_THREADS)
group_sum_kernel(const float *__restrict__ X, int ldx, int64_t N, int d, int G, const int32_t *__restrict__ perm,
                 const int64_t *__restrict__ seg, const int32_t *__restrict__ chunk_start,
                 double *__restrict__ sums /*[G][d]*/) {
    const Chunk ch = find_chunk(blockIdx.x, G, seg, chunk_start, N);
    if (ch.n == 0) return;
    extern __shared__ __align__(16) float csm[];   // [GC_THREADS/32][d] warp partials
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = warp; c < ch.n; c += GC_THREADS / 32) {
        const int64_t id = perm ? (int64_t)perm[ch.pos + c] : ch.pos + c;
        const float *x = X + id * ldx;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < d) acc[q] += x[lane + 32 * q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < d) csm[warp * d + lane + 32 * q] = acc[q];
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += GC_THREADS) {
        double t = 0.0;
        for (int w = 0; w < GC_THREADS / 32; ++w) t += (double)csm[w * d + j];
        atomicAdd(&sums[(int64_t)ch.level * d + j], t);
    }
}

__global__ void __launch_bounds__(GC_THREADS)
group_cov_kernel(const float *__restrict__ X, int ldx, int64_t N, int d, int G, const int32_t *__restrict__ perm,
                 const int64_t *__restrict__ seg, const int32_t *__restrict__ chunk_start,
                 const float *__restrict__ mean /*[G][d]*/, double *__restrict__ cov /*[G][d][d]*/) {
    const Chunk ch = find_chunk(blockIdx.x, G, seg, chunk_start, N);
    if (ch.n == 0) return;
    extern __shared__ __align__(16) float csm[];
    float *xs = csm;                                                    // [32][d+1] centred members
    int32_t *ids = reinterpret_cast<int32_t *>(csm + 32 * (d + 1));     // [MOE_CH] member cell ids
    const int tid = threadIdx.x;
    for (int c = tid; c < ch.n; c += GC_THREADS) ids[c] = perm ? perm[ch.pos + c] : (int32_t)(ch.pos + c);
    const int nent = d * d;
    const float *mu = mean + (int64_t)ch.level * d;
    // x - mean is exact in FP32 (both are floats within a factor of two for any cluster worth scoring); the
    // products are summed in FP64: small clusters have ill-conditioned covariances whose inverse amplifies
    // FP32 accumulation error beyond the 1e-4 parity bar (2 d^2 FP64 flop per cell is nothing here)
    double acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
    for (int s0 = 0; s0 < ch.n; s0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * d; i += GC_THREADS) {
            const int c = i / d, j = i - c * d;
            xs[c * (d + 1) + j] = (s0 + c < ch.n) ? X[(int64_t)ids[s0 + c] * ldx + j] - mu[j] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int ent = tid + e * GC_THREADS;
            if (ent < nent) {
                const int i = ent / d, j = ent - i * d;
                double a = acc[e];
#pragma unroll 8
                for (int c = 0; c < 32; ++c) a = fma((double)xs[c * (d + 1) + i], (double)xs[c * (d + 1) + j], a);
                acc[e] = a;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int ent = tid + e * GC_THREADS;
        if (ent < nent) atomicAdd(&cov[(int64_t)ch.level * nent + ent], acc[e]);
    }
}

}  // namespace sym

extern "C" int sym_cluster_moments(const float *X, int32_t ldx, int64_t M, int32_t d, const float *R, int32_t K,
                                   double *sums, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(M > 0 && d > 0 && d <= 128 && K > 0 && ldx >= d, SYM_ERR_INVALID_ARGUMENT, "sym_cluster_moments: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    SYM_CUDA(cudaMemsetAsync(sums, 0, (size_t)K * (d + 2) * sizeof(double), st));
    dim3 grid((unsigned)ceil_div(M, CC_CHUNK), (unsigned)K);
    const size_t smem = (size_t)(CC_THREADS / 32) * (d + 2) * sizeof(float);
    cluster_moments_kernel<<<grid, CC_THREADS, smem, st>>>(X, ldx, M, d, R, K, sums);
    SYM_LAUNCH_CHECK("cluster_moments_kernel");
    return SYM_OK;
}

extern "C" int sym_cluster_cov(const float *X, int32_t ldx, int64_t M, int32_t d, const float *R, int32_t K,
                               const float *mean, double *cov, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(M > 0 && d > 0 && d <= 64 && K > 0 && ldx >= d, SYM_ERR_INVALID_ARGUMENT,
                "sym_cluster_cov: bad sizes (d <= 64)");
    cudaStream_t st = (cudaStream_t)stream;
    SYM_CUDA(cudaMemsetAsync(cov, 0, (size_t)K * d * d * sizeof(double), st));
    dim3 grid((unsigned)ceil_div(M, CC_CHUNK), (unsigned)K);
    const size_t smem = (size_t)(32 * (d + 1) + 32) * sizeof(float);
    cluster_cov_kernel<<<grid, CC_THREADS, smem, st>>>(X, ldx, M, d, R, K, mean, cov);
    SYM_LAUNCH_CHECK("cluster_cov_kernel");
    return SYM_OK;
}

extern "C" int sym_cell_confidence(const float *Xq, int32_t ldq, int64_t N, int32_t d, const float *Rq, int32_t K,
                                   const float *mean, const float *inv_cov, float *out, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && d > 0 && d <= 64 && K > 0 && ldq >= d, SYM_ERR_INVALID_ARGUMENT,
                "sym_cell_confidence: bad sizes (d <= 64)");
    if (N == 0) return SYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (d <= 32)
        cell_confidence_kernel<32, 128><<<(unsigned)ceil_div(N, 128), 128, 0, st>>>(Xq, ldq, N, d, Rq, K, mean, inv_cov, out);
    else
        cell_confidence_kernel<64, 64><<<(unsigned)ceil_div(N, 64), 64, 0, st>>>(Xq, ldq, N, d, Rq, K, mean, inv_cov, out);
    SYM_LAUNCH_CHECK("cell_confidence_kernel");
    return SYM_OK;
}

static int group_args_ok(const int32_t *perm, const int64_t *seg, const int32_t *chunk_start, int32_t G) {
    return ((perm == nullptr) == (seg == nullptr)) && ((perm == nullptr) == (chunk_start == nullptr)) &&
           (perm != nullptr || G == 1);
}

extern "C" int sym_group_sums(const float *X, int32_t ldx, int64_t N, int32_t d, int32_t G, const int32_t *perm,
                              const int64_t *seg, const int32_t *chunk_start, double *sums, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && d > 0 && d <= 128 && G > 0 && ldx >= d, SYM_ERR_INVALID_ARGUMENT, "sym_group_sums: bad sizes");
    SYM_REQUIRE(group_args_ok(perm, seg, chunk_start, G), SYM_ERR_INVALID_ARGUMENT,
                "sym_group_sums: several groups need the perm, seg and chunk_start of sym_batch_order");
    cudaStream_t st = (cudaStream_t)stream;
    SYM_CUDA(cudaMemsetAsync(sums, 0, (size_t)G * d * sizeof(double), st));
    if (N == 0) return SYM_OK;
    const int64_t grid = ceil_div(N, MOE_CH) + (perm ? G : 0);
    group_sum_kernel<<<(unsigned)grid, GC_THREADS, (size_t)(GC_THREADS / 32) * d * sizeof(float), st>>>(
        X, ldx, N, d, G, perm, seg, chunk_start, sums);
    SYM_LAUNCH_CHECK("group_sum_kernel");
    return SYM_OK;
}

extern "C" int sym_group_cov(const float *X, int32_t ldx, int64_t N, int32_t d, int32_t G, const int32_t *perm,
                             const int64_t *seg, const int32_t *chunk_start, const float *mean, double *cov,
                             sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && d > 0 && d <= 64 && G > 0 && ldx >= d, SYM_ERR_INVALID_ARGUMENT,
                "sym_group_cov: bad sizes (d <= 64)");
    SYM_REQUIRE(group_args_ok(perm, seg, chunk_start, G), SYM_ERR_INVALID_ARGUMENT,
                "sym_group_cov: several groups need the perm, seg and chunk_start of sym_batch_order");
    cudaStream_t st = (cudaStream_t)stream;
    SYM_CUDA(cudaMemsetAsync(cov, 0, (size_t)G * d * d * sizeof(double), st));
    if (N == 0) return SYM_OK;
    const int64_t grid = ceil_div(N, MOE_CH) + (perm ? G : 0);
    const size_t smem = (size_t)(32 * (d + 1)) * sizeof(float) + MOE_CH * sizeof(int32_t);
    group_cov_kernel<<<(unsigned)grid, GC_THREADS, smem, st>>>(X, ldx, N, d, G, perm, seg, chunk_start, mean, cov);
    SYM_LAUNCH_CHECK("group_cov_kernel");
    return SYM_OK;
}
