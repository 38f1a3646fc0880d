// (2) assign: soft cluster membership of each query cell       — symphonypy/_utils.py:154-178
//
//   zc = sign(rowmax(z)) * z / ||z||     (the reference divides by the SIGNED row max before
//                                          the L2 normalisation, _utils.py:168-170, which flips
//                                          rows whose entries are all negative)
//   R[n,k] = softmax_k( -2 (1 - Y_k . zc) * inv_sigma_k )
//
// HBM-bound (reads 4d, writes 4K bytes per cell).  One warp owns CT cells at a time; lane l
// owns centroids l, l+32, ...; the row reductions (max, L2 norm, softmax max and sum) are
// warp shuffles.  Y^T lives in shared memory for the whole CTA lifetime (persistent grid).
#include "common.cuh"

namespace sym {

constexpr int AS_THREADS = 256;
constexpr int AS_WARPS = AS_THREADS / 32;

template <int KJ, int CT>
__global__ void __launch_bounds__(AS_THREADS)
assign_kernel(const float *__restrict__ Z, int ldz, int64_t N, int d, const float *__restrict__ Y, int K,
              const float *__restrict__ inv_sigma, float *__restrict__ R) {
    extern __shared__ __align__(16) float smem[];
    const int KP = KJ * 32;
    float *Yt = smem;                         // [d][KP]  (transposed, zero padded)
    float *zs = smem + (size_t)d * KP;        // [AS_WARPS][CT][d]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < d * KP; i += AS_THREADS) {
        const int dd = i / KP, k = i % KP;
        Yt[i] = k < K ? Y[(int64_t)k * d + dd] : 0.f;
    }
    float isg[KJ];
#pragma unroll
    for (int j = 0; j < KJ; ++j) isg[j] = (lane + 32 * j < K) ? inv_sigma[lane + 32 * j] : 0.f;
    __syncthreads();

    float *zw = zs + (size_t)warp * CT * d;
    const int64_t groups = ceil_div(N, CT);
    for (int64_t grp = (int64_t)blockIdx.x * AS_WARPS + warp; grp < groups; grp += (int64_t)gridDim.x * AS_WARPS) {
        const int64_t n0 = grp * CT;
        // normalise CT rows into the warp's scratch
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int64_t n = n0 + c;
            float mx = -INFINITY, ss = 0.f;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int dd = lane + 32 * q;
                v[q] = (n < N && dd < d) ? Z[n * ldz + dd] : 0.f;
                if (dd < d) mx = fmaxf(mx, v[q]);
                ss = fmaf(v[q], v[q], ss);
            }
            mx = warp_max(mx);
            ss = warp_sum(ss);
            // x / max / ||x / max||  ==  sign(max) * x / ||x||
            const float sc = (mx < 0.f ? -1.f : 1.f) / sqrtf(ss);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int dd = lane + 32 * q;
                if (dd < d) zw[c * d + dd] = v[q] * sc;
            }
        }
        __syncwarp();
        float acc[CT][KJ];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int j = 0; j < KJ; ++j) acc[c][j] = 0.f;
        for (int dd = 0; dd < d; ++dd) {
            float y[KJ];
#pragma unroll
            for (int j = 0; j < KJ; ++j) y[j] = Yt[dd * KP + lane + 32 * j];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float z = zw[c * d + dd];
#pragma unroll
                for (int j = 0; j < KJ; ++j) acc[c][j] = fmaf(z, y[j], acc[c][j]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int64_t n = n0 + c;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                acc[c][j] = -2.f * (1.f - acc[c][j]) * isg[j];
                if (lane + 32 * j < K) mx = fmaxf(mx, acc[c][j]);
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                acc[c][j] = (lane + 32 * j < K) ? expf(acc[c][j] - mx) : 0.f;
                sum += acc[c][j];
            }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            if (n < N) {
#pragma unroll
                for (int j = 0; j < KJ; ++j)
                    if (lane + 32 * j < K) R[n * K + lane + 32 * j] = acc[c][j] * inv;
            }
        }
    }
}

template <int KJ, int CT>
static int launch_assign(const float *Z, int ldz, int64_t N, int d, const float *Y, int K, const float *inv_sigma,
                         float *R, cudaStream_t st) {
    auto kern = assign_kernel<KJ, CT>;
    const size_t smem = ((size_t)d * KJ * 32 + (size_t)AS_WARPS * CT * d) * sizeof(float);
    SYM_REQUIRE(smem <= 220 * 1024, SYM_ERR_UNSUPPORTED_SHAPE, "sym_assign: K=%d d=%d needs %zu B of shared memory", K, d,
                smem);
    SYM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = smem > 100 * 1024 ? 1 : (smem > 50 * 1024 ? 2 : 4);
    int64_t grid = ceil_div(ceil_div(N, CT), AS_WARPS);
    if (grid > (int64_t)kNumSMs * per_sm) grid = (int64_t)kNumSMs * per_sm;
    kern<<<(unsigned)grid, AS_THREADS, smem, st>>>(Z, ldz, N, d, Y, K, inv_sigma, R);
    SYM_LAUNCH_CHECK("assign_kernel");
    return SYM_OK;
}

}  // namespace sym

extern "C" int sym_assign(const float *Z, int32_t ldz, int64_t N, int32_t d, const float *Y, int32_t K,
                          const float *inv_sigma, float *R, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && d > 0 && K > 0 && ldz >= d, SYM_ERR_INVALID_ARGUMENT, "sym_assign: bad sizes");
    SYM_REQUIRE(d <= 128 && K <= 1024, SYM_ERR_UNSUPPORTED_SHAPE, "sym_assign: d=%d (<=128) K=%d (<=1024) unsupported", d,
                K);
    if (N == 0) return SYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 128) return launch_assign<4, 4>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
    if (K <= 256) return launch_assign<8, 4>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
    if (K <= 512) return launch_assign<16, 2>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
    return launch_assign<32, 1>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
}
