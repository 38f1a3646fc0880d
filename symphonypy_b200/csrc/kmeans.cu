// k-means++ seeding for the no-Harmony fallback                  — symphonypy/_utils.py:323-325
//
// The reference calls sklearn's KMeans(init="k-means++"): greedy k-means++ (sklearn/cluster/_kmeans.py,
// _kmeans_plusplus).  Per new centre: draw `trials` = 2 + int(ln K) candidates with probability proportional to
// the squared distance to the nearest centre so far (a searchsorted into the running sum of those distances), take
// the candidate that lowers the total potential most.  Given the same uniform draws the result is the same set of
// rows (tests compare with a numpy restatement pinned to sklearn itself).
//
// Everything stays on the device, four small launches per centre, no host round trip:
//   sample : block sums of `closest` -> the block a draw falls into -> scan inside that block   (1 CTA)
//   trial  : per point min(closest, ||x - cand_t||^2) for the T candidates, block partial sums   (N / 1024 CTAs)
//   select : potentials of the candidates, argmin                                                (1 CTA)
//   commit : closest = min(closest, ||x - chosen||^2), new block sums                            (N / 1024 CTAs)
// Distances use sklearn's formula ||x||^2 - 2 x.c + ||c||^2 in float64 (clamped at 0); X is float32, so they equal
// sklearn's on data that is exactly representable in float32.  At N = 500k, d = 30 the two passes read 60 MB each
// (L2 resident): a seeding of K = 100 centres is ~400 launches, a few ms (torch matmuls before: 40 ms).
#include "common.cuh"

namespace sym {

constexpr int KPP_THREADS = 256;
constexpr int KPP_ITEMS = 4;                         // points per thread
constexpr int KPP_BLOCK = KPP_THREADS * KPP_ITEMS;   // points per CTA and per block sum
constexpr int KPP_MAXT = 16;                         // candidates per centre (2 + ln K: 8 at K = 1024)
constexpr int KPP_DMAX = 128;

__device__ __forceinline__ double block_sum(double v, double *sh) {   // sum over the CTA (KPP_THREADS threads)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < KPP_THREADS / 32; ++w) t += sh[w];
    return t;
}

// ||x_n - c||^2 by sklearn's expansion, float64
__device__ __forceinline__ double kpp_dist(const float *__restrict__ x, int d, const float *__restrict__ c, double x2n,
                                           double c2) {
    double dot = 0.0;
    for (int j = 0; j < d; ++j) dot = fma((double)x[j], (double)c[j], dot);
    return fmax(x2n - 2.0 * dot + c2, 0.0);
}

// x2[n] = ||x_n||^2; closest[n] = ||x_n - x_first||^2; blocksum[b]
__global__ void __launch_bounds__(KPP_THREADS)
kpp_init_kernel(const float *__restrict__ X, int ldx, int64_t N, int d, int64_t first, double *__restrict__ x2,
                double *__restrict__ closest, double *__restrict__ blocksum, int32_t *__restrict__ centre_ids) {
    __shared__ float c[KPP_DMAX];
    __shared__ double sh[KPP_THREADS / 32];
    for (int j = threadIdx.x; j < d; j += KPP_THREADS) c[j] = X[first * ldx + j];
    __syncthreads();
    double c2 = 0.0;
    for (int j = 0; j < d; ++j) c2 = fma((double)c[j], (double)c[j], c2);
    double acc = 0.0;
    for (int i = 0; i < KPP_ITEMS; ++i) {
        const int64_t n = (int64_t)blockIdx.x * KPP_BLOCK + i * KPP_THREADS + threadIdx.x;
        if (n < N) {
            const float *x = X + n * ldx;
            double s = 0.0;
            for (int j = 0; j < d; ++j) s = fma((double)x[j], (double)x[j], s);
            x2[n] = s;
            const double dd = kpp_dist(x, d, c, s, c2);
            closest[n] = dd;
            acc += dd;
        }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        blocksum[blockIdx.x] = acc;
        if (blockIdx.x == 0) centre_ids[0] = (int32_t)first;
    }
}

// candidate ids of one centre: searchsorted(cumsum(closest), rand[t] * pot), clipped to N - 1
__global__ void __launch_bounds__(KPP_THREADS)
kpp_sample_kernel(const double *__restrict__ closest, const double *__restrict__ blocksum, int nb, int64_t N,
                  const double *__restrict__ rand, int T, int32_t *__restrict__ cand_ids) {
    __shared__ double part[KPP_THREADS];      // sums of contiguous chunks of block sums, then their exclusive prefix
    __shared__ double sh[KPP_THREADS / 32];
    __shared__ double wsum[KPP_THREADS / 32];
    __shared__ int sel_block[KPP_MAXT];
    __shared__ double sel_before[KPP_MAXT], sel_target[KPP_MAXT];
    __shared__ int found;
    const int tid = threadIdx.x;
    const int chunk = (nb + KPP_THREADS - 1) / KPP_THREADS;
    double s = 0.0;
    for (int b = tid * chunk; b < min(nb, (tid + 1) * chunk); ++b) s += blocksum[b];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {   // exclusive prefix over the chunks (256 sequential adds)
        double run = 0.0;
        for (int i = 0; i < KPP_THREADS; ++i) {
            const double v = part[i];
            part[i] = run;
            run += v;
        }
        sh[0] = run;   // the potential
    }
    __syncthreads();
    const double pot = sh[0];
    if (tid < T) {   // the block a draw falls into
        const double target = rand[tid] * pot;
        int c = 0;
        while (c + 1 < KPP_THREADS && part[c + 1] < target) ++c;   // last chunk whose exclusive prefix is below the target
        double run = part[c];
        int b = c * chunk;
        const int bend = min(nb, (c + 1) * chunk);
        while (b + 1 < bend && run + blocksum[b] < target) {
            run += blocksum[b];
            ++b;
        }
        if (b >= nb) b = nb - 1;
        sel_block[tid] = b;
        sel_before[tid] = run;
        sel_target[tid] = target;
    }
    __syncthreads();
    for (int t = 0; t < T; ++t) {   // inside the block: first index whose running sum reaches the target
        const int b = sel_block[t];
        const double before = sel_before[t], target = sel_target[t];
        const int64_t base = (int64_t)b * KPP_BLOCK + (int64_t)tid * KPP_ITEMS;   // each thread: KPP_ITEMS consecutive points
        double v[KPP_ITEMS], loc = 0.0;
#pragma unroll
        for (int i = 0; i < KPP_ITEMS; ++i) {
            v[i] = base + i < N ? closest[base + i] : 0.0;
            loc += v[i];
        }
        // exclusive prefix of the per-thread sums over the CTA
        double inc = loc;
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (tid == 0) found = 0x7fffffff;
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        double wbefore = 0.0;
        for (int w = 0; w < warp; ++w) wbefore += wsum[w];
        double run = before + wbefore + inc - loc;
#pragma unroll
        for (int i = 0; i < KPP_ITEMS; ++i) {
            run += v[i];
            if (run >= target && base + i < N) {
                atomicMin(&found, tid * KPP_ITEMS + i);
                break;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int64_t id = found == 0x7fffffff ? N - 1 : (int64_t)b * KPP_BLOCK + found;
            // (rounding can leave the target above the block's total: the draw then belongs to a later point)
            if (found == 0x7fffffff && b + 1 < nb) id = (int64_t)(b + 1) * KPP_BLOCK;
            cand_ids[t] = (int32_t)(id < N ? id : N - 1);
        }
        __syncthreads();
    }
}

// partial[b][t] = sum over the block's points of min(closest, ||x - cand_t||^2)
__global__ void __launch_bounds__(KPP_THREADS)
kpp_trial_kernel(const float *__restrict__ X, int ldx, int64_t N, int d, const double *__restrict__ x2,
                 const double *__restrict__ closest, const int32_t *__restrict__ cand_ids, int T,
                 double *__restrict__ partial) {
    extern __shared__ float cand[];   // [T][d]
    __shared__ double c2[KPP_MAXT];
    __shared__ double sh[KPP_THREADS / 32];
    for (int i = threadIdx.x; i < T * d; i += KPP_THREADS) cand[i] = X[(int64_t)cand_ids[i / d] * ldx + (i % d)];
    if (threadIdx.x < T) c2[threadIdx.x] = x2[cand_ids[threadIdx.x]];
    __syncthreads();
    double acc[KPP_MAXT];
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) acc[t] = 0.0;
    for (int i = 0; i < KPP_ITEMS; ++i) {
        const int64_t n = (int64_t)blockIdx.x * KPP_BLOCK + i * KPP_THREADS + threadIdx.x;
        if (n < N) {
            const float *x = X + n * ldx;
            const double x2n = x2[n], cl = closest[n];
#pragma unroll
            for (int t = 0; t < KPP_MAXT; ++t)
                if (t < T) acc[t] += fmin(cl, kpp_dist(x, d, cand + t * d, x2n, c2[t]));
        }
    }
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) {
        if (t < T) {   // (T is uniform over the CTA)
            const double s = block_sum(acc[t], sh);
            if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * T + t] = s;
        }
    }
}

// the candidate with the lowest potential (first one on ties, as np.argmin) becomes centre c
__global__ void __launch_bounds__(KPP_THREADS)
kpp_select_kernel(const double *__restrict__ partial, int nb, int T, const int32_t *__restrict__ cand_ids, int c,
                  int32_t *__restrict__ centre_ids) {
    __shared__ double sh[KPP_THREADS / 32];
    __shared__ double pots[KPP_MAXT];
    for (int t = 0; t < T; ++t) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nb; b += KPP_THREADS) s += partial[(int64_t)b * T + t];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) pots[t] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        for (int t = 1; t < T; ++t)
            if (pots[t] < pots[best]) best = t;
        centre_ids[c] = cand_ids[best];
    }
}

// closest = min(closest, ||x - X[centre_ids[c]]||^2), and the block sums of the new values
__global__ void __launch_bounds__(KPP_THREADS)
kpp_commit_kernel(const float *__restrict__ X, int ldx, int64_t N, int d, const double *__restrict__ x2,
                  double *__restrict__ closest, const int32_t *__restrict__ centre_ids, int c,
                  double *__restrict__ blocksum) {
    __shared__ float ctr[KPP_DMAX];
    __shared__ double sh[KPP_THREADS / 32];
    const int64_t id = centre_ids[c];
    for (int j = threadIdx.x; j < d; j += KPP_THREADS) ctr[j] = X[id * ldx + j];
    __syncthreads();
    const double c2 = x2[id];
    double acc = 0.0;
    for (int i = 0; i < KPP_ITEMS; ++i) {
        const int64_t n = (int64_t)blockIdx.x * KPP_BLOCK + i * KPP_THREADS + threadIdx.x;
        if (n < N) {
            const double v = fmin(closest[n], kpp_dist(X + n * ldx, d, ctr, x2[n], c2));
            closest[n] = v;
            acc += v;
        }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = acc;
}

}  // namespace sym

using namespace sym;

extern "C" size_t sym_kmeanspp_workspace(int64_t N, int32_t trials) {
    if (N <= 0 || trials <= 0) return 256;
    const int64_t nb = ceil_div(N, KPP_BLOCK);
    return (size_t)(2 * N + nb + nb * trials) * sizeof(double) + KPP_MAXT * sizeof(int32_t) + 1024;
}

extern "C" int sym_kmeanspp(const float *X, int32_t ldx, int64_t N, int32_t d, int32_t K, int32_t trials, int64_t first,
                            const double *rand, int32_t *centre_ids, void *workspace, size_t workspace_bytes,
                            sym_stream_t stream) {
    SYM_REQUIRE(N > 0 && d > 0 && d <= KPP_DMAX && ldx >= d && K >= 1 && K <= N, SYM_ERR_INVALID_ARGUMENT,
                "sym_kmeanspp: bad sizes");
    SYM_REQUIRE(trials >= 1 && trials <= KPP_MAXT, SYM_ERR_INVALID_ARGUMENT, "sym_kmeanspp: 1 <= trials <= %d", KPP_MAXT);
    SYM_REQUIRE(first >= 0 && first < N && N < (1ll << 31), SYM_ERR_INVALID_ARGUMENT, "sym_kmeanspp: bad first index / N");
    SYM_REQUIRE(workspace_bytes >= sym_kmeanspp_workspace(N, trials), SYM_ERR_WORKSPACE_TOO_SMALL,
                "sym_kmeanspp: workspace too small");
    SYM_REQUIRE(((uintptr_t)workspace & 7) == 0, SYM_ERR_INVALID_ARGUMENT, "sym_kmeanspp: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (int)ceil_div(N, KPP_BLOCK);
    double *x2 = (double *)workspace, *closest = x2 + N, *blocksum = closest + N, *partial = blocksum + nb;
    int32_t *cand_ids = (int32_t *)(partial + (size_t)nb * trials);
    kpp_init_kernel<<<nb, KPP_THREADS, 0, st>>>(X, ldx, N, d, first, x2, closest, blocksum, centre_ids);
    SYM_LAUNCH_CHECK("kpp_init_kernel");
    for (int c = 1; c < K; ++c) {
        kpp_sample_kernel<<<1, KPP_THREADS, 0, st>>>(closest, blocksum, nb, N, rand + (size_t)(c - 1) * trials, trials,
                                                      cand_ids);
        SYM_LAUNCH_CHECK("kpp_sample_kernel");
        kpp_trial_kernel<<<nb, KPP_THREADS, (size_t)trials * d * sizeof(float), st>>>(X, ldx, N, d, x2, closest, cand_ids,
                                                                                      trials, partial);
        SYM_LAUNCH_CHECK("kpp_trial_kernel");
        kpp_select_kernel<<<1, KPP_THREADS, 0, st>>>(partial, nb, trials, cand_ids, c, centre_ids);
        SYM_LAUNCH_CHECK("kpp_select_kernel");
        kpp_commit_kernel<<<nb, KPP_THREADS, 0, st>>>(X, ldx, N, d, x2, closest, centre_ids, c, blocksum);
        SYM_LAUNCH_CHECK("kpp_commit_kernel");
    }
    return SYM_OK;
}
