// (1) project: Z = clip((X - mu) * inv_sd, +-clip) @ U      — symphonypy/_utils.py:298-308
//
// Dense kernel (FP32 FFMA): one CTA owns BM cells and all d outputs and streams the G genes
// in chunks of 32.  X is read exactly once from HBM with 128-bit streaming loads into
// registers (next chunk prefetched while the current one is multiplied), scaled and clipped
// in registers (the clip makes the affine part non-factorisable, so it has to sit on the
// A-operand path) and parked in shared memory; the U chunk arrives by cp.async (L2 resident:
// U is G*d*4 <= 1 MB).  Thread tile: 8 cells x 4 outputs.
//
// CSR kernel: Z = z0 + sum_nnz (clip((x-mu)/sd) - clip(-mu/sd)) * U[g,:], one warp per cell,
// which never materialises the dense matrix the reference builds at _utils.py:240-243.
#include "common.cuh"

namespace sym {

constexpr int PJ_THREADS = 256;
constexpr int PJ_GK = 32;       // genes per stage
constexpr int PJ_XS = PJ_GK + 4;  // padded row stride of the X tile (floats)

template <int DPAD>
struct ProjCfg {
    static constexpr int TX = DPAD / 4;            // threads across outputs
    static constexpr int TY = PJ_THREADS / TX;     // thread rows
    static constexpr int BM = TY * 8;              // cells per CTA
    static constexpr int XLD = BM / 32;            // float4 X loads per thread per stage
    static constexpr size_t smem = 2 * (size_t)(BM * PJ_XS + PJ_GK * DPAD) * sizeof(float);
};

__device__ __forceinline__ float scale_clip(float x, float mu, float is, float clip) {
    return fminf(fmaxf((x - mu) * is, -clip), clip);
}

template <int DPAD, bool GATHER>
__global__ void __launch_bounds__(PJ_THREADS, 2)
project_dense_kernel(const float *__restrict__ X, int64_t N, int64_t ldx, const int32_t *__restrict__ gene_col,
                     int G, const float *__restrict__ mu, const float *__restrict__ inv_sd,
                     const float *__restrict__ U, int ldu, int d, float clip, float *__restrict__ Z, int ldz) {
    using C = ProjCfg<DPAD>;
    extern __shared__ __align__(16) float smem[];
    float *Xs = smem;                          // [2][BM][PJ_XS]
    float *Us = smem + 2 * C::BM * PJ_XS;      // [2][PJ_GK][DPAD]

    const int tid = threadIdx.x;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const int64_t row0 = (int64_t)blockIdx.x * C::BM;
    const int nchunks = (G + PJ_GK - 1) / PJ_GK;

    // X staging role: float4 column chunk lc of tile row lr + 32*j
    const int lc = tid % 8, lr = tid / 8;
    float4 xr[C::XLD];

    auto load_x = [&](int g0) {
        const int g = g0 + lc * 4;
#pragma unroll
        for (int j = 0; j < C::XLD; ++j) {
            const int64_t row = row0 + lr + 32 * j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < N) {
                if (!GATHER) {
                    if (g + 3 < G) {
                        v = ldg_stream(reinterpret_cast<const float4 *>(X + row * ldx + g));
                    } else {
                        const float *p = X + row * ldx;
                        if (g < G) v.x = p[g];
                        if (g + 1 < G) v.y = p[g + 1];
                        if (g + 2 < G) v.z = p[g + 2];
                    }
                } else {
                    const float *p = X + row * ldx;
                    float t[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        int c = -1;
                        if (g + e < G) c = gene_col ? gene_col[g + e] : g + e;
                        t[e] = c >= 0 ? __ldg(p + c) : 0.f;
                    }
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            xr[j] = v;
        }
    };
    auto store_x = [&](int g0, int stage) {
        const int g = g0 + lc * 4;
        float m[4], s[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool ok = g + e < G;
            m[e] = ok ? __ldg(mu + g + e) : 0.f;
            s[e] = ok ? __ldg(inv_sd + g + e) : 0.f;
        }
        float *base = Xs + stage * C::BM * PJ_XS;
#pragma unroll
        for (int j = 0; j < C::XLD; ++j) {
            float4 v = xr[j];
            // genes with inv_sd == 0 (absent from the query or zero std) contribute exactly 0
            v.x = s[0] != 0.f ? scale_clip(v.x, m[0], s[0], clip) : 0.f;
            v.y = s[1] != 0.f ? scale_clip(v.y, m[1], s[1], clip) : 0.f;
            v.z = s[2] != 0.f ? scale_clip(v.z, m[2], s[2], clip) : 0.f;
            v.w = s[3] != 0.f ? scale_clip(v.w, m[3], s[3], clip) : 0.f;
            *reinterpret_cast<float4 *>(base + (lr + 32 * j) * PJ_XS + lc * 4) = v;
        }
    };
    auto load_u = [&](int g0, int stage) {
        float *base = Us + stage * PJ_GK * DPAD;
        constexpr int V4 = PJ_GK * DPAD / 4;
        for (int i = tid; i < V4; i += PJ_THREADS) {
            const int r = i / (DPAD / 4), c4 = (i % (DPAD / 4)) * 4;
            const bool ok = (g0 + r < G) && (c4 + 4 <= ldu);
            const float *src = ok ? U + (int64_t)(g0 + r) * ldu + c4 : U;
            cp_async16_zfill(base + r * DPAD + c4, src, ok);
        }
        cp_async_commit();
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[i][n] = 0.f;

    load_x(0);
    load_u(0, 0);
    store_x(0, 0);
    cp_async_wait<0>();
    __syncthreads();

    for (int it = 0; it < nchunks; ++it) {
        const int cur = it & 1;
        const bool more = it + 1 < nchunks;
        if (more) {
            load_x((it + 1) * PJ_GK);
            load_u((it + 1) * PJ_GK, cur ^ 1);
        }
        const float *xs = Xs + cur * C::BM * PJ_XS;
        const float *us = Us + cur * PJ_GK * DPAD;
#pragma unroll
        for (int kk = 0; kk < PJ_GK; kk += 4) {
            float4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = *reinterpret_cast<const float4 *>(us + (kk + j) * DPAD + tx * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 x = *reinterpret_cast<const float4 *>(xs + (ty + C::TY * i) * PJ_XS + kk);
                acc[i][0] = fmaf(x.x, u[0].x, acc[i][0]);
                acc[i][1] = fmaf(x.x, u[0].y, acc[i][1]);
                acc[i][2] = fmaf(x.x, u[0].z, acc[i][2]);
                acc[i][3] = fmaf(x.x, u[0].w, acc[i][3]);
                acc[i][0] = fmaf(x.y, u[1].x, acc[i][0]);
                acc[i][1] = fmaf(x.y, u[1].y, acc[i][1]);
                acc[i][2] = fmaf(x.y, u[1].z, acc[i][2]);
                acc[i][3] = fmaf(x.y, u[1].w, acc[i][3]);
                acc[i][0] = fmaf(x.z, u[2].x, acc[i][0]);
                acc[i][1] = fmaf(x.z, u[2].y, acc[i][1]);
                acc[i][2] = fmaf(x.z, u[2].z, acc[i][2]);
                acc[i][3] = fmaf(x.z, u[2].w, acc[i][3]);
                acc[i][0] = fmaf(x.w, u[3].x, acc[i][0]);
                acc[i][1] = fmaf(x.w, u[3].y, acc[i][1]);
                acc[i][2] = fmaf(x.w, u[3].z, acc[i][2]);
                acc[i][3] = fmaf(x.w, u[3].w, acc[i][3]);
            }
        }
        if (more) {
            store_x((it + 1) * PJ_GK, cur ^ 1);
            cp_async_wait<0>();
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = row0 + ty + C::TY * i;
        if (row >= N) continue;
        float *z = Z + row * ldz + tx * 4;
#pragma unroll
        for (int n = 0; n < 4; ++n)
            if (tx * 4 + n < d) z[n] = acc[i][n];
    }
}

// z0[j] = sum_g clip(-mu_g * inv_sd_g) * U[g, j]   (one block, d threads x gene slices)
__global__ void project_csr_bias_kernel(const float *__restrict__ mu, const float *__restrict__ inv_sd,
                                        const float *__restrict__ U, int ldu, int G, int d, float clip,
                                        float *__restrict__ z0) {
    __shared__ double part[8][128];
    const int j = threadIdx.x % 128, sl = threadIdx.x / 128;  // 1024 threads: 8 slices x 128 outputs
    double a = 0.0;
    if (j < d)
        for (int g = sl; g < G; g += 8) {
            const float s = inv_sd[g];
            if (s != 0.f) a += (double)scale_clip(0.f, mu[g], s, clip) * (double)U[(int64_t)g * ldu + j];
        }
    part[sl][j] = a;
    __syncthreads();
    if (sl == 0 && j < d) {
        for (int s = 1; s < 8; ++s) a += part[s][j];
        z0[j] = (float)a;
    }
}

// One warp per cell; lane l owns outputs l, l+32, l+64, l+96.
// `not_canonical` (optional): set to 1 when a row's column indices are not strictly increasing — a duplicated column
// has to be summed BEFORE the scale and clip (the reference densifies first, `_utils.py:284-285`), which this kernel
// cannot do term by term; the host then canonicalises the matrix and runs again.  The check rides on the index loads
// the kernel makes anyway (scipy's own `has_canonical_format` is an O(nnz) pass on one host core).
__global__ void __launch_bounds__(256)
project_csr_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices,
                   const int64_t *__restrict__ indptr, int64_t N, const int32_t *__restrict__ col_gene,
                   const float *__restrict__ mu, const float *__restrict__ inv_sd, const float *__restrict__ U,
                   int ldu, int d, float clip, const float *__restrict__ z0, float *__restrict__ Z, int ldz,
                   int32_t *__restrict__ not_canonical) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const int64_t beg = indptr[row], end = indptr[row + 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    bool bad = false;
    for (int64_t base = beg; base < end; base += 32) {
        // each lane fetches one non-zero and resolves its coefficient; the warp then walks them
        const int64_t j = base + lane;
        float coef = 0.f;
        int g = -1;
        const int col = j < end ? indices[j] : 0x7fffffff;
        if (not_canonical != nullptr) {
            int prev = __shfl_up_sync(0xffffffffu, col, 1);
            if (lane == 0) prev = base > beg ? indices[base - 1] : -1;
            bad |= j < end && col <= prev;
        }
        if (j < end) {
            g = col_gene[col];
            if (g >= 0) {
                const float s = inv_sd[g], m = mu[g];
                coef = s != 0.f ? scale_clip(data[j], m, s, clip) - scale_clip(0.f, m, s, clip) : 0.f;
                if (coef == 0.f) g = -1;
            }
        }
        unsigned live = __ballot_sync(0xffffffffu, g >= 0);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            const float c = __shfl_sync(0xffffffffu, coef, src);
            const int gg = __shfl_sync(0xffffffffu, g, src);
            const float *u = U + (int64_t)gg * ldu;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane + 32 * q < d) acc[q] = fmaf(c, __ldg(u + lane + 32 * q), acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < d) Z[row * ldz + lane + 32 * q] = z0[lane + 32 * q] + acc[q];
    if (not_canonical != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(not_canonical, 1);
}

template <int DPAD, bool GATHER>
static int launch_project(const float *X, int64_t N, int64_t ldx, const int32_t *gene_col, int G, const float *mu,
                          const float *inv_sd, const float *U, int ldu, int d, float clip, float *Z, int ldz,
                          cudaStream_t st) {
    using C = ProjCfg<DPAD>;
    auto kern = project_dense_kernel<DPAD, GATHER>;
    SYM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
    const int64_t grid = ceil_div(N, C::BM);
    kern<<<(unsigned)grid, PJ_THREADS, C::smem, st>>>(X, N, ldx, gene_col, G, mu, inv_sd, U, ldu, d, clip, Z, ldz);
    SYM_LAUNCH_CHECK("project_dense_kernel");
    return SYM_OK;
}

}  // namespace sym

extern "C" int sym_project_dense(const float *X, int64_t N, int64_t ldx, const int32_t *gene_col, int32_t G,
                                 const float *mu, const float *inv_sd, const float *U, int32_t ldu, int32_t d,
                                 float clip, float *Z, int32_t ldz, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && G > 0 && d > 0, SYM_ERR_INVALID_ARGUMENT, "sym_project_dense: bad sizes N=%lld G=%d d=%d",
                (long long)N, G, d);
    SYM_REQUIRE(d <= 128, SYM_ERR_UNSUPPORTED_SHAPE, "sym_project_dense: d=%d > 128 is not supported", d);
    SYM_REQUIRE(ldu >= d && ldu % 4 == 0 && ((uintptr_t)U & 15) == 0, SYM_ERR_INVALID_ARGUMENT,
                "sym_project_dense: U must be 16-byte aligned with ldu %% 4 == 0 and ldu >= d (ldu=%d d=%d)", ldu, d);
    SYM_REQUIRE(ldz >= d, SYM_ERR_INVALID_ARGUMENT, "sym_project_dense: ldz < d");
    SYM_REQUIRE(ceil_div(N, 64) < (1ll << 31), SYM_ERR_UNSUPPORTED_SHAPE, "sym_project_dense: N too large");
    if (N == 0) return SYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool gather = gene_col != nullptr || (ldx % 4) != 0 || ((uintptr_t)X & 15) != 0;
#define SYM_PJ(DP)                                                                                               \
    return gather ? launch_project<DP, true>(X, N, ldx, gene_col, G, mu, inv_sd, U, ldu, d, clip, Z, ldz, st)   \
                  : launch_project<DP, false>(X, N, ldx, gene_col, G, mu, inv_sd, U, ldu, d, clip, Z, ldz, st)
    if (d <= 32) { SYM_PJ(32); }
    if (d <= 64) { SYM_PJ(64); }
    SYM_PJ(128);
#undef SYM_PJ
}

extern "C" int sym_project_csr_bias(const float *mu, const float *inv_sd, const float *U, int32_t ldu, int32_t G,
                                    int32_t d, float clip, float *z0, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(G > 0 && d > 0 && d <= 128 && ldu >= d, SYM_ERR_INVALID_ARGUMENT, "sym_project_csr_bias: bad sizes");
    project_csr_bias_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mu, inv_sd, U, ldu, G, d, clip, z0);
    SYM_LAUNCH_CHECK("project_csr_bias_kernel");
    return SYM_OK;
}

extern "C" int sym_project_csr_checked(const float *data, const int32_t *indices, const int64_t *indptr, int64_t N,
                                       const int32_t *col_gene, const float *mu, const float *inv_sd, const float *U,
                                       int32_t ldu, int32_t d, float clip, const float *z0, float *Z, int32_t ldz,
                                       int32_t *not_canonical, sym_stream_t stream);

extern "C" int sym_project_csr(const float *data, const int32_t *indices, const int64_t *indptr, int64_t N,
                               const int32_t *col_gene, const float *mu, const float *inv_sd, const float *U,
                               int32_t ldu, int32_t d, float clip, const float *z0, float *Z, int32_t ldz,
                               sym_stream_t stream) {
    return sym_project_csr_checked(data, indices, indptr, N, col_gene, mu, inv_sd, U, ldu, d, clip, z0, Z, ldz, nullptr,
                                   stream);
}

extern "C" int sym_project_csr_checked(const float *data, const int32_t *indices, const int64_t *indptr, int64_t N,
                                       const int32_t *col_gene, const float *mu, const float *inv_sd, const float *U,
                                       int32_t ldu, int32_t d, float clip, const float *z0, float *Z, int32_t ldz,
                                       int32_t *not_canonical, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && d > 0 && d <= 128 && ldu >= d && ldz >= d, SYM_ERR_INVALID_ARGUMENT,
                "sym_project_csr: bad sizes");
    if (N == 0) return SYM_OK;
    const int64_t grid = ceil_div(N, 8);
    SYM_REQUIRE(grid < (1ll << 31), SYM_ERR_UNSUPPORTED_SHAPE, "sym_project_csr: N too large");
    project_csr_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(data, indices, indptr, N, col_gene, mu, inv_sd,
                                                                         U, ldu, d, clip, z0, Z, ldz, not_canonical);
    SYM_LAUNCH_CHECK("project_csr_kernel");
    return SYM_OK;
}
