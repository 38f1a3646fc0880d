// (2) assign: soft cluster membership of each query cell       — symphonypy/_utils.py:154-178
//
//   zc = sign(rowmax(z)) * z / ||z||     (the reference divides by the SIGNED row max before
//                                          the L2 normalisation, _utils.py:168-170, which flips
//                                          rows whose entries are all negative)
//   R[n,k] = softmax_k( -2 (1 - Y_k . zc) * inv_sigma_k )
//
// 2 N K d flop in FP32 on 4(d + K) bytes per cell: at the BASELINE shapes (d = 30, K = 100) the FP32 pipe, not HBM, is
// the floor (94 warp-FFMAs per cell).  What the first version of this kernel was actually bound by is the shared-memory
// pipe — it re-read the centroid matrix for every 4 cells and reduced every row with 20 shuffles — so this one
//   * gives a warp CT cells at a time (16 for K <= 128), lane l owning centroids 4l .. 4l+3 (+128 v): one LDS.128 per
//     component fetches the lane's centroid values for all CT cells, one broadcast LDS.128 fetches 4 components of a
//     cell; per 4 components that is 4 + CT shared-memory instructions for 16 CT FFMAs,
//   * reduces the CT row sums / row maxima together (`warp_reduce_rows`: CT + 1 shuffles instead of 5 CT) and
//     needs no row maximum for the sign (a ballot: the row flips iff no entry is >= 0),
//   * writes R with 16-byte stores when K is a multiple of 4.
// Y^T lives in shared memory for the whole CTA lifetime (persistent grid).
#include "common.cuh"

namespace sym {

constexpr int AS_THREADS = 256;
constexpr int AS_WARPS = AS_THREADS / 32;

struct OpAdd {
    __device__ __forceinline__ float operator()(float a, float b) const { return a + b; }
};
struct OpMax {
    __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
struct FinSame {
    __device__ __forceinline__ float operator()(float a) const { return a; }
};
struct FinInvSqrt {
    __device__ __forceinline__ float operator()(float a) const { return 1.f / sqrtf(a); }
};
struct FinInv {
    __device__ __forceinline__ float operator()(float a) const { return 1.f / a; }
};
// 2^x, 2 ulp (MUFU.EX2); -inf and anything below -126 give 0
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// v[c] (c < CT, one value per lane and row) -> the reduction over the 32 lanes of row c, for all rows at once:
// each halving step sends away the half of the rows this lane will not finish (CT/2 + CT/4 + ... shuffles), the
// lanes that end up with the same row finish with a plain butterfly.  Lane l ends with row  l >> (5 - log2 CT)  in
// v[0]; the totals are then passed through `red` (CT floats of this warp's shared memory) so that every lane has all.
// `fin` is applied to a row's total once, by the lane that publishes it (a square root or a division per row, not per lane).
template <int CT, class Op, class Fin>
__device__ __forceinline__ void warp_reduce_rows(float (&v)[CT], float *red, int lane, Op op, Fin fin) {
    static_assert(CT >= 1 && CT <= 32 && (CT & (CT - 1)) == 0, "CT must be a power of two");
    int w = 16;
#pragma unroll
    for (int n = CT; n > 1; n >>= 1, w >>= 1) {
        const bool upper = lane & w;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, w));
        }
    }
#pragma unroll
    for (; w >= 1; w >>= 1) v[0] = op(v[0], __shfl_xor_sync(0xffffffffu, v[0], w));
    constexpr int SHARE = 32 / CT;               // lanes that hold the same row
    __syncwarp();
    if ((lane & (SHARE - 1)) == 0) red[lane / SHARE] = fin(v[0]);
    __syncwarp();
    if constexpr (CT >= 4) {
#pragma unroll
        for (int c = 0; c < CT; c += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(red + c);
            v[c] = t.x, v[c + 1] = t.y, v[c + 2] = t.z, v[c + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < CT; ++c) v[c] = red[c];
    }
}

// KV: 128-centroid slabs per lane (lane l owns centroids 128 v + 4 l + j, j < 4); CT cells per warp step
template <int KV, int CT>
__global__ void __launch_bounds__(AS_THREADS, KV == 1 ? 2 : 1)
assign_kernel(const float *__restrict__ Z, int ldz, int64_t N, int d, const float *__restrict__ Y, int K,
              const float *__restrict__ inv_sigma, float *__restrict__ R) {
    extern __shared__ __align__(16) float smem[];
    constexpr int KP = KV * 128;
    const int DP = (d + 3) & ~3;
    float *Yt = smem;                                   // [DP][KP]  (transposed, zero padded)
    float *zs = Yt + (size_t)DP * KP;                   // [AS_WARPS][CT][DP]
    float *reds = zs + (size_t)AS_WARPS * CT * DP;      // [AS_WARPS][CT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < DP * KP; i += AS_THREADS) {
        const int dd = i / KP, k = i % KP;
        Yt[i] = (k < K && dd < d) ? Y[(int64_t)k * d + dd] : 0.f;
    }
    // logit_k * log2(e) = (cos_k - 1) * g_k with g_k = 2 log2(e) / sigma_k.  The padding centroids (zero columns of
    // Yt: cos = 0) get a huge g, i.e. a logit of -3e38: they never win the row maximum and their exponential is 0.
    float g[KV][4];
#pragma unroll
    for (int v = 0; v < KV; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = v * 128 + lane * 4 + j;
            g[v][j] = k < K ? 2.f * 1.4426950408889634f * inv_sigma[k] : 3.0e38f;
        }
    float *zw = zs + (size_t)warp * CT * DP;
    float *red = reds + warp * CT;
    for (int i = lane; i < CT * DP; i += 32) zw[i] = 0.f;      // (the padding components stay zero)
    __syncthreads();
    const bool vec_store = (K & 3) == 0;

    const int64_t groups = ceil_div(N, CT);
    for (int64_t grp = (int64_t)blockIdx.x * AS_WARPS + warp; grp < groups; grp += (int64_t)gridDim.x * AS_WARPS) {
        const int64_t n0 = grp * CT;
        // ---- normalise CT rows into the warp's scratch: sign(rowmax) / ||z||
        float ss[CT];
        unsigned nonneg[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) ss[c] = 0.f, nonneg[c] = 0u;
        for (int q = 0; q * 32 < d; ++q) {
            const int dd = lane + 32 * q;
            float x[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) x[c] = (n0 + c < N && dd < d) ? Z[(n0 + c) * ldz + dd] : 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                ss[c] = fmaf(x[c], x[c], ss[c]);
                nonneg[c] |= __ballot_sync(0xffffffffu, dd < d && !(x[c] < 0.f));
                if (dd < d) zw[c * DP + dd] = x[c];
            }
        }
        warp_reduce_rows<CT>(ss, red, lane, OpAdd(), FinInvSqrt());
        // x / max / ||x / max||  ==  sign(max) * x / ||x||; the row max is negative iff no entry is >= 0
        // (a NaN entry counts as "not negative": it only matters for rows that come out as NaN anyway)
#pragma unroll
        for (int c = 0; c < CT; ++c) ss[c] = nonneg[c] ? ss[c] : -ss[c];
        __syncwarp();

        float acc[CT][KV][4];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int v = 0; v < KV; ++v)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[c][v][j] = 0.f;
#pragma unroll 1
        for (int d4 = 0; d4 < DP; d4 += 4) {
            if constexpr (KV <= 2) {
                // few centroids per lane, many cells: 4 components of the lane's centroids in registers, cells streamed
                float4 y[4][KV];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < KV; ++v)
                        y[u][v] = *reinterpret_cast<const float4 *>(Yt + (size_t)(d4 + u) * KP + v * 128 + lane * 4);
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const float4 z = *reinterpret_cast<const float4 *>(zw + c * DP + d4);
#pragma unroll
                    for (int v = 0; v < KV; ++v) {
                        acc[c][v][0] = fmaf(z.w, y[3][v].x, fmaf(z.z, y[2][v].x, fmaf(z.y, y[1][v].x, fmaf(z.x, y[0][v].x, acc[c][v][0]))));
                        acc[c][v][1] = fmaf(z.w, y[3][v].y, fmaf(z.z, y[2][v].y, fmaf(z.y, y[1][v].y, fmaf(z.x, y[0][v].y, acc[c][v][1]))));
                        acc[c][v][2] = fmaf(z.w, y[3][v].z, fmaf(z.z, y[2][v].z, fmaf(z.y, y[1][v].z, fmaf(z.x, y[0][v].z, acc[c][v][2]))));
                        acc[c][v][3] = fmaf(z.w, y[3][v].w, fmaf(z.z, y[2][v].w, fmaf(z.y, y[1][v].w, fmaf(z.x, y[0][v].w, acc[c][v][3]))));
                    }
                }
            } else {
                // many centroids per lane, few cells: the cells' 4 components in registers, centroid components streamed
                float z[CT][4];
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const float4 t = *reinterpret_cast<const float4 *>(zw + c * DP + d4);
                    z[c][0] = t.x, z[c][1] = t.y, z[c][2] = t.z, z[c][3] = t.w;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < KV; ++v) {
                        const float4 y = *reinterpret_cast<const float4 *>(Yt + (size_t)(d4 + u) * KP + v * 128 + lane * 4);
#pragma unroll
                        for (int c = 0; c < CT; ++c) {
                            acc[c][v][0] = fmaf(z[c][u], y.x, acc[c][v][0]);
                            acc[c][v][1] = fmaf(z[c][u], y.y, acc[c][v][1]);
                            acc[c][v][2] = fmaf(z[c][u], y.z, acc[c][v][2]);
                            acc[c][v][3] = fmaf(z[c][u], y.w, acc[c][v][3]);
                        }
                    }
            }
        }
        __syncwarp();

        // ---- softmax over the K centroids of every row (base 2: the logits carry the factor log2 e)
        float mx[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            mx[c] = -INFINITY;
#pragma unroll
            for (int v = 0; v < KV; ++v) {
#pragma unroll
                for (int j = 0; j < 4; ++j)   // (the scale of the row is applied to the dot product: Y_k . (s z) = s (Y_k . z))
                    acc[c][v][j] = fmaf(acc[c][v][j] * ss[c], g[v][j], -g[v][j]);
                mx[c] = fmaxf(fmaxf(mx[c], fmaxf(acc[c][v][0], acc[c][v][1])), fmaxf(acc[c][v][2], acc[c][v][3]));
            }
        }
        warp_reduce_rows<CT>(mx, red, lane, OpMax(), FinSame());
        float sum[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            sum[c] = 0.f;
#pragma unroll
            for (int v = 0; v < KV; ++v)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[c][v][j] = fast_exp2(acc[c][v][j] - mx[c]);
                    sum[c] += acc[c][v][j];
                }
        }
        warp_reduce_rows<CT>(sum, red, lane, OpAdd(), FinInv());
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int64_t n = n0 + c;
            if (n >= N) break;
            const float inv = sum[c];
            float *row = R + n * K;
#pragma unroll
            for (int v = 0; v < KV; ++v) {
                const int k = v * 128 + lane * 4;
                if (vec_store) {
                    if (k < K)
                        *reinterpret_cast<float4 *>(row + k) = make_float4(acc[c][v][0] * inv, acc[c][v][1] * inv,
                                                                           acc[c][v][2] * inv, acc[c][v][3] * inv);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (k + j < K) row[k + j] = acc[c][v][j] * inv;
                }
            }
        }
        __syncwarp();
    }
}

template <int KV, int CT>
static int launch_assign(const float *Z, int ldz, int64_t N, int d, const float *Y, int K, const float *inv_sigma,
                         float *R, cudaStream_t st) {
    auto kern = assign_kernel<KV, CT>;
    const int DP = (d + 3) & ~3;
    const size_t smem = ((size_t)DP * KV * 128 + (size_t)AS_WARPS * CT * DP + AS_WARPS * CT) * sizeof(float);
    SYM_REQUIRE(smem <= 220 * 1024, SYM_ERR_UNSUPPORTED_SHAPE, "sym_assign: K=%d d=%d needs %zu B of shared memory", K, d,
                smem);
    SYM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int by_regs = KV == 1 ? 2 : 1;
    int per_sm = smem > 110 * 1024 ? 1 : 2;
    if (per_sm > by_regs) per_sm = by_regs;
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
    if (K <= 128) return launch_assign<1, 16>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
    if (K <= 256) return launch_assign<2, 8>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
    if (K <= 512) return launch_assign<4, 4>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
    return launch_assign<8, 2>(Z, ldz, N, d, Y, K, inv_sigma, R, st);
}
