// (3) mixture-of-experts correction                              — symphonypy/_utils.py:181-214
//
// The reference loops over the K clusters and, for each, streams all N cells through three
// dense GEMMs against the (B+1) x N design matrix Phi.  Phi is one-hot per batch variable
// (tl.py:368-378), so here it never exists: cells are grouped by level code with a counting
// sort (sym_batch_order) and
//   stats : per level c, cross[k,c,:] = sum_n R[n,k] z_n, gram[k,c,c] = sum_n R[n,k]   (one small
//           GEMM R^T [Z|1] per 256-cell chunk of a level, FP32 partials, FP64 atomics)
//   solve : K independent (B+1)x(B+1) ridge systems, FP64 Cholesky in shared memory
//   apply : z_n -= R[n,:] W[:,c(n),:]   (one [cells x K][K x d] GEMM per chunk)
// All K clusters are handled together, and the statistics are a [K,B1,B1+d] FP64 block that a
// multi-GPU run all-reduces between stats and solve.
#include "common.cuh"
#include "work_chunks.cuh"

namespace sym {

constexpr int BO_THREADS = 256;
constexpr int BO_ITEMS = 4096;    // cells per CTA in the ordering kernels
constexpr int BO_SMEM_BINS = 4096;

// ------------------------------------------------------------------------------------------
// batch order (counting sort by level code)
// workspace layout: int32 count[B] | int32 cursor[B]
// N < 2^31 per GPU is required (perm is int32).
__global__ void bo_hist_kernel(const int32_t *__restrict__ codes, int64_t N, int B, int32_t *__restrict__ count) {
    __shared__ int32_t h[BO_SMEM_BINS];
    const bool priv = B <= BO_SMEM_BINS;
    if (priv) {
        for (int i = threadIdx.x; i < B; i += blockDim.x) h[i] = 0;
        __syncthreads();
    }
    const int64_t beg = (int64_t)blockIdx.x * BO_ITEMS;
    const int64_t end = beg + BO_ITEMS < N ? beg + BO_ITEMS : N;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        int c = codes[i];
        c = c < 0 ? 0 : (c >= B ? B - 1 : c);
        if (priv) atomicAdd(&h[c], 1); else atomicAdd(&count[c], 1);
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < B; i += blockDim.x)
            if (h[i]) atomicAdd(&count[i], h[i]);
    }
}

__global__ void bo_scan_kernel(const int32_t *__restrict__ count, int B, int64_t *__restrict__ seg,
                               int32_t *__restrict__ chunk_start, int32_t *__restrict__ cursor) {
    // single thread block, serial over B in chunks (B is small: number of batch levels)
    if (threadIdx.x == 0) {
        int64_t s = 0;
        int32_t c = 0;
        for (int b = 0; b < B; ++b) {
            seg[b] = s;
            chunk_start[b] = c;
            cursor[b] = (int32_t)s;
            s += count[b];
            c += (count[b] + MOE_CH - 1) / MOE_CH;
        }
        seg[B] = s;
        chunk_start[B] = c;
    }
}

__global__ void bo_scatter_kernel(const int32_t *__restrict__ codes, int64_t N, int B, int32_t *__restrict__ cursor,
                                  int32_t *__restrict__ perm) {
    __shared__ int32_t h[BO_SMEM_BINS];  // local count, then local base
    const bool priv = B <= BO_SMEM_BINS;
    const int64_t beg = (int64_t)blockIdx.x * BO_ITEMS;
    const int64_t end = beg + BO_ITEMS < N ? beg + BO_ITEMS : N;
    if (!priv) {
        for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
            int c = codes[i];
            c = c < 0 ? 0 : (c >= B ? B - 1 : c);
            perm[atomicAdd(&cursor[c], 1)] = (int32_t)i;
        }
        return;
    }
    for (int i = threadIdx.x; i < B; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        int c = codes[i];
        c = c < 0 ? 0 : (c >= B ? B - 1 : c);
        atomicAdd(&h[c], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const int n = h[i];
        h[i] = n ? atomicAdd(&cursor[i], n) : 0;  // reserve a contiguous range per level
    }
    __syncthreads();
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        int c = codes[i];
        c = c < 0 ? 0 : (c >= B ? B - 1 : c);
        perm[atomicAdd(&h[c], 1)] = (int32_t)i;
    }
}

// ------------------------------------------------------------------------------------------
// stats: P[k, 0..d] = sum_{cells of chunk} R[n,k] * [z_n | 1]
// Thread tile: 8 clusters x 4 columns of [Z|1].  256 threads cover KT=128 clusters x 64 columns
// per pass; larger K / d+1 loop over (k0, c0) passes re-reading the staged sub-tiles.
constexpr int ST_THREADS = 256;
constexpr int ST_SUB = 32;  // cells staged per step
constexpr int MOE_MAX_VARS = 16;
struct VarOffsets {
    int off[MOE_MAX_VARS];
};

// Thread layout: 16 x 16 threads, 8 clusters x 4 columns each (128 clusters x 64 columns of [Z|1] per pass), or — NARROW,
// for d + 1 <= 32, where the wide layout leaves half of the threads without a column — 32 x 8 threads, 4 clusters x 4
// columns each (128 clusters x 32 columns).
template <bool NARROW>
__global__ void __launch_bounds__(ST_THREADS)
moe_stats_kernel(const float *__restrict__ Z, int ldz, const float *__restrict__ R, int64_t N, int d, int K,
                 const int32_t *__restrict__ codes, int V, VarOffsets var_offset, int v, int n_levels,
                 const int32_t *__restrict__ perm, const int64_t *__restrict__ seg,
                 const int32_t *__restrict__ chunk_start, int B1, double *__restrict__ gram,
                 double *__restrict__ cross) {
    const Chunk ch = find_chunk(blockIdx.x, n_levels, seg, chunk_start, N);
    if (ch.n == 0) return;
    extern __shared__ __align__(16) float smem[];
    const int KP = (K + 7) & ~7;           // R row stride in smem
    const int DP = ((d + 1) + 3) & ~3;     // [Z|1] row stride
    // two stages of [ST_SUB][KP] R rows and [ST_SUB][DP] [Z|1] rows, filled by cp.async straight from global
    // memory (no register round trip: the load -> STS dependency of a synchronous fill was 58 % of this kernel's
    // stalls), the next stage in flight while the current one is multiplied
    const int stage_floats = ST_SUB * KP + ST_SUB * DP;
    int32_t *ids = reinterpret_cast<int32_t *>(smem + 2 * stage_floats);  // [MOE_CH] cell ids of the chunk
    const int tid = threadIdx.x;
    for (int i = tid; i < ch.n; i += ST_THREADS) ids[i] = perm ? perm[ch.pos + i] : (int32_t)(ch.pos + i);
    __syncthreads();

    const int col = 1 + var_offset.off[v] + ch.level;  // design column of this level
    constexpr int TCN = NARROW ? 8 : 16, KT = NARROW ? 4 : 8, CW = NARROW ? 32 : 64;   // column threads, clusters per thread, columns per pass
    const int tk = tid / TCN, tc = tid % TCN;
    const bool r16 = (K % 4) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0;          // 16-byte copies of R rows
    const bool z8 = (d % 2) == 0 && (ldz % 2) == 0 && (reinterpret_cast<uintptr_t>(Z) & 7) == 0;   // 8-byte copies of Z rows
    const int kv = KP / 4, zv = z8 ? d / 2 : d;
    auto fill = [&](int stage, int s0) {
        float *Rs = smem + stage * stage_floats, *Zs = Rs + ST_SUB * KP;
        const int ns = ch.n - s0 < ST_SUB ? ch.n - s0 : ST_SUB;
        if (r16) {
            for (int i = tid; i < ST_SUB * kv; i += ST_THREADS) {
                const int c = i / kv, k = (i - c * kv) * 4;
                const bool ok = c < ns && k < K;
                cp_async16_zfill(Rs + c * KP + k, ok ? R + (int64_t)ids[s0 + c] * K + k : R, ok);
            }
        } else {
            for (int i = tid; i < ST_SUB * KP; i += ST_THREADS) {
                const int c = i / KP, k = i - c * KP;
                const bool ok = c < ns && k < K;
                cp_async4_zfill(Rs + i, ok ? R + (int64_t)ids[s0 + c] * K + k : R, ok);
            }
        }
        for (int i = tid; i < ST_SUB * zv; i += ST_THREADS) {
            const int c = i / zv, j = i - c * zv;
            const bool ok = c < ns;
            if (z8) cp_async8_zfill(Zs + c * DP + 2 * j, ok ? Z + (int64_t)ids[s0 + c] * ldz + 2 * j : Z, ok);
            else cp_async4_zfill(Zs + c * DP + j, ok ? Z + (int64_t)ids[s0 + c] * ldz + j : Z, ok);
        }
        for (int i = tid; i < ST_SUB * (DP - d); i += ST_THREADS) {   // the ones column and the padding
            const int c = i / (DP - d), j = d + (i - c * (DP - d));
            Zs[c * DP + j] = (j == d && c < ns) ? 1.f : 0.f;
        }
        cp_async_commit();
    };
    for (int k0 = 0; k0 < K; k0 += 128) {
        for (int c0 = 0; c0 < d + 1; c0 += CW) {
            float acc[KT][4];
#pragma unroll
            for (int i = 0; i < KT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            __syncthreads();   // everyone is done with both stages of the previous pass
            fill(0, 0);
            int stage = 0;
            for (int s0 = 0; s0 < ch.n; s0 += ST_SUB, stage ^= 1) {
                if (s0 + ST_SUB < ch.n) {
                    fill(stage ^ 1, s0 + ST_SUB);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const float *Rs = smem + stage * stage_floats, *Zs = Rs + ST_SUB * KP;
                const int kb = k0 + tk * KT, cb = c0 + tc * 4;
                if (kb < KP && cb < DP) {
#pragma unroll 4
                    for (int c = 0; c < ST_SUB; ++c) {
                        const float4 r0 = *reinterpret_cast<const float4 *>(Rs + c * KP + kb);
                        const float4 r1 = NARROW ? r0 : *reinterpret_cast<const float4 *>(Rs + c * KP + kb + 4);
                        const float4 z = *reinterpret_cast<const float4 *>(Zs + c * DP + cb);
                        const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int i = 0; i < KT; ++i) {
                            acc[i][0] = fmaf(r[i], z.x, acc[i][0]);
                            acc[i][1] = fmaf(r[i], z.y, acc[i][1]);
                            acc[i][2] = fmaf(r[i], z.z, acc[i][2]);
                            acc[i][3] = fmaf(r[i], z.w, acc[i][3]);
                        }
                    }
                }
                __syncthreads();   // before the stage just read is refilled (two fills from now)
            }
            const int kb = k0 + tk * KT, cb = c0 + tc * 4;
#pragma unroll
            for (int i = 0; i < KT; ++i) {
                const int k = kb + i;
                if (k >= K) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = cb + j;
                    if (c < d)
                        atomicAdd(&cross[((int64_t)k * B1 + col) * d + c], (double)acc[i][j]);
                    else if (c == d)
                        atomicAdd(&gram[((int64_t)k * B1 + col) * B1 + col], (double)acc[i][j]);
                }
            }
        }
    }
    // co-occurrence with the levels of later variables (off-diagonal blocks of the Gram)
    for (int v2 = v + 1; v2 < V; ++v2) {
        const int32_t *c2 = codes + (int64_t)v2 * N;
        const int off2 = 1 + var_offset.off[v2];
        for (int64_t i = tid; i < (int64_t)ch.n * K; i += ST_THREADS) {
            const int c = (int)(i / K), k = (int)(i % K);
            const int64_t n = ids[c];
            const int col2 = off2 + c2[n];
            const double r = (double)R[n * K + k];
            atomicAdd(&gram[((int64_t)k * B1 + col) * B1 + col2], r);
            atomicAdd(&gram[((int64_t)k * B1 + col2) * B1 + col], r);
        }
    }
}

// ------------------------------------------------------------------------------------------
// solve: one CTA per cluster.  A = gram_k (+ intercept row/col, + Nr_k, + diag(lamb)), Cholesky.
constexpr int SV_THREADS = 256;

__global__ void __launch_bounds__(SV_THREADS)
moe_solve_kernel(const double *__restrict__ gram, const double *__restrict__ cross, const double *__restrict__ Nr,
                 const double *__restrict__ C, const double *__restrict__ lamb, int K, int B1, int d, int n_lev0,
                 float *__restrict__ W, int32_t *__restrict__ info) {
    extern __shared__ __align__(16) double sm[];
    const int k = blockIdx.x, tid = threadIdx.x;
    const double *Ag = gram + (int64_t)k * B1 * B1;
    const double *Hk = cross + (int64_t)k * B1 * d;
    double *A = sm;                       // [B1][B1]
    double *rhs = sm + (size_t)B1 * B1;   // [B1][d]
    __shared__ int bad;
    if (tid == 0) bad = 0;
    for (int i = tid; i < B1 * B1; i += SV_THREADS) A[i] = Ag[i];
    for (int i = tid; i < B1 * d; i += SV_THREADS) rhs[i] = Hk[i];
    __syncthreads();

    // intercept row / column: Phi row 0 is all ones and variable 0 partitions the cells, so
    //   G[0,0] = sum_{levels c of var 0} G[c,c],  G[0,c] = G[c,0] = G[c,c]   (_utils.py:199)
    //   H[0,:] = sum_{levels c of var 0} H[c,:]                              (_utils.py:204)
    __shared__ double A0;
    if (tid == 0) {
        double s = 0.0;
        for (int c = 1; c <= n_lev0; ++c) s += A[c * B1 + c];
        A0 = s + Nr[k];
        A[0] = A0 + lamb[0];  // _utils.py:201; the reference forces lamb[0,0] = 0 (:191)
    }
    for (int c = tid; c < d; c += SV_THREADS) {
        double h0 = C[(int64_t)k * d + c];  // _utils.py:205
        for (int b = 1; b <= n_lev0; ++b) h0 += rhs[b * d + c];
        rhs[c] = h0;
    }
    __syncthreads();
    for (int c = 1 + tid; c < B1; c += SV_THREADS) {
        const double g = A[c * B1 + c];
        A[c] = g;
        A[c * B1] = g;
        A[c * B1 + c] = g + lamb[c];
    }
    __syncthreads();

    // right-looking Cholesky on the lower triangle: A = L L^T.  A pivot that cancels to rounding
    // noise of its original diagonal entry means the ridge system is singular (np.linalg.inv raises).
    for (int j = 0; j < B1; ++j) {
        const double piv = A[j * B1 + j];
        const double d0 = Ag[j * B1 + j] + lamb[j] + (j == 0 ? A0 : 0.0);
        if (!(piv > 1e-13 * d0)) {  // uniform: every thread reads the same value
            if (tid == 0) bad = j + 1;
            break;
        }
        const double l = sqrt(piv);
        __syncthreads();
        if (tid == 0) A[j * B1 + j] = l;
        for (int i = j + 1 + tid; i < B1; i += SV_THREADS) A[i * B1 + j] /= l;
        __syncthreads();
        const int m = B1 - j - 1;
        for (int t = tid; t < m * m; t += SV_THREADS) {
            const int i = j + 1 + t / m, c = j + 1 + t % m;
            if (c <= i) A[i * B1 + c] -= A[i * B1 + j] * A[c * B1 + j];
        }
        __syncthreads();
    }
    __syncthreads();
    if (bad) {
        if (tid == 0) info[k] = bad;
        for (int i = tid; i < B1 * d; i += SV_THREADS) W[(int64_t)k * B1 * d + i] = __int_as_float(0x7fc00000);
        return;
    }
    if (tid == 0) info[k] = 0;

    // one thread per right-hand-side column: forward then backward substitution
    for (int c = tid; c < d; c += SV_THREADS) {
        for (int i = 0; i < B1; ++i) {
            double s = rhs[i * d + c];
            for (int j = 0; j < i; ++j) s -= A[i * B1 + j] * rhs[j * d + c];
            rhs[i * d + c] = s / A[i * B1 + i];
        }
        for (int i = B1 - 1; i >= 0; --i) {
            double s = rhs[i * d + c];
            for (int j = i + 1; j < B1; ++j) s -= A[j * B1 + i] * rhs[j * d + c];
            rhs[i * d + c] = s / A[i * B1 + i];
        }
        for (int i = 0; i < B1; ++i)
            W[((int64_t)k * B1 + i) * d + c] = i == 0 ? 0.f : (float)rhs[i * d + c];  // _utils.py:209
    }
}

// ------------------------------------------------------------------------------------------
// apply: Zout[n,:] = Zin[n,:] - R[n,:] @ W[:, col, :]
// Thread tile 4 cells x 4 outputs; 256 threads = 64 cells x 64 outputs per pass, or — NARROW, for d <= 32, where the
// wide layout leaves half of the threads without an output — 128 cells x 32 outputs.
constexpr int AP_THREADS = 256;
constexpr int AP_KS = 32;  // clusters staged per step

template <bool NARROW>
__global__ void __launch_bounds__(AP_THREADS)
moe_apply_kernel(const float *__restrict__ Zin, int ldzin, const float *__restrict__ R, int64_t N, int d, int K,
                 int var_offset, int n_levels, const int32_t *__restrict__ perm, const int64_t *__restrict__ seg,
                 const int32_t *__restrict__ chunk_start, const float *__restrict__ W, int B1,
                 float *__restrict__ Zout, int ldzout) {
    const Chunk ch = find_chunk(blockIdx.x, n_levels, seg, chunk_start, N);
    if (ch.n == 0) return;
    extern __shared__ __align__(16) float smem[];
    const int DP = (d + 3) & ~3;
    const int KP = (K + 3) & ~3;
    float *Ws = smem;                 // [KP][DP]   W[:, col, :]
    constexpr int RP = NARROW ? 128 : 64, TCN = NARROW ? 8 : 16, TRN = AP_THREADS / TCN;   // rows per pass, column / row threads
    float *Rs0 = Ws + (size_t)KP * DP; // 2 stages of [RP][AP_KS + 4], filled by cp.async
    constexpr int RS = RP * (AP_KS + 4);
    int32_t *ids = reinterpret_cast<int32_t *>(Rs0 + 2 * RS);
    const int tid = threadIdx.x;
    const int col = 1 + var_offset + ch.level;
    for (int i = tid; i < ch.n; i += AP_THREADS) ids[i] = perm ? perm[ch.pos + i] : (int32_t)(ch.pos + i);
    for (int i = tid; i < KP * DP; i += AP_THREADS) {
        const int k = i / DP, j = i % DP;
        Ws[i] = (k < K && j < d) ? W[((int64_t)k * B1 + col) * d + j] : 0.f;
    }
    __syncthreads();
    const int tr = tid / TCN, tc = tid % TCN;  // rows tr + TRN*i (i<4), outputs tc*4..+3
    const bool r16 = (K % 4) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0;
    auto fill = [&](int stage, int r0, int k0) {   // R[rows r0..r0+RP-1, clusters k0..k0+31] -> stage
        float *Rs = Rs0 + stage * RS;
        if (r16) {
            for (int i = tid; i < RP * (AP_KS / 4); i += AP_THREADS) {
                const int r = i / (AP_KS / 4), k = (i % (AP_KS / 4)) * 4;
                const bool ok = r0 + r < ch.n && k0 + k < K;
                cp_async16_zfill(Rs + r * (AP_KS + 4) + k, ok ? R + (int64_t)ids[r0 + r] * K + k0 + k : R, ok);
            }
        } else {
            for (int i = tid; i < RP * AP_KS; i += AP_THREADS) {
                const int r = i / AP_KS, k = i % AP_KS;
                const bool ok = r0 + r < ch.n && k0 + k < K;
                cp_async4_zfill(Rs + r * (AP_KS + 4) + k, ok ? R + (int64_t)ids[r0 + r] * K + k0 + k : R, ok);
            }
        }
        cp_async_commit();
    };
    for (int r0 = 0; r0 < ch.n; r0 += RP) {
        for (int c0 = 0; c0 < d; c0 += TCN * 4) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            __syncthreads();   // both stages free
            fill(0, r0, 0);
            int stage = 0;
            for (int k0 = 0; k0 < K; k0 += AP_KS, stage ^= 1) {
                if (k0 + AP_KS < K) {
                    fill(stage ^ 1, r0, k0 + AP_KS);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const float *Rs = Rs0 + stage * RS;
                const int cb = c0 + tc * 4;
                if (cb < DP) {
#pragma unroll 8
                    for (int k = 0; k < AP_KS; ++k) {
                        if (k0 + k >= KP) break;
                        const float4 w = *reinterpret_cast<const float4 *>(Ws + (k0 + k) * DP + cb);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float r = Rs[(tr + TRN * i) * (AP_KS + 4) + k];
                            acc[i][0] = fmaf(r, w.x, acc[i][0]);
                            acc[i][1] = fmaf(r, w.y, acc[i][1]);
                            acc[i][2] = fmaf(r, w.z, acc[i][2]);
                            acc[i][3] = fmaf(r, w.w, acc[i][3]);
                        }
                    }
                }
                __syncthreads();   // before the stage just read is refilled
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + tr + TRN * i;
                if (r >= ch.n) continue;
                const int64_t n = ids[r];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c0 + tc * 4 + j;
                    if (c < d) Zout[n * ldzout + c] = Zin[n * ldzin + c] - acc[i][j];
                }
            }
        }
    }
}

}  // namespace sym

// ============================================================================================
extern "C" size_t sym_batch_order_workspace(int32_t B) { return (size_t)(B > 0 ? B : 1) * 2 * sizeof(int32_t) + 64; }

extern "C" int sym_batch_order(const int32_t *codes, int64_t N, int32_t B, int32_t *perm, int64_t *seg,
                               int32_t *chunk_start, void *workspace, size_t workspace_bytes, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && B > 0, SYM_ERR_INVALID_ARGUMENT, "sym_batch_order: bad sizes");
    SYM_REQUIRE(N < (1ll << 31), SYM_ERR_UNSUPPORTED_SHAPE, "sym_batch_order: N >= 2^31 cells per GPU");
    SYM_REQUIRE(workspace_bytes >= sym_batch_order_workspace(B), SYM_ERR_WORKSPACE_TOO_SMALL,
                "sym_batch_order: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *count = (int32_t *)workspace, *cursor = count + B;
    SYM_CUDA(cudaMemsetAsync(count, 0, (size_t)B * sizeof(int32_t), st));
    const unsigned grid = (unsigned)(N > 0 ? ceil_div(N, BO_ITEMS) : 1);
    bo_hist_kernel<<<grid, BO_THREADS, 0, st>>>(codes, N, B, count);
    SYM_LAUNCH_CHECK("bo_hist_kernel");
    bo_scan_kernel<<<1, 32, 0, st>>>(count, B, seg, chunk_start, cursor);
    SYM_LAUNCH_CHECK("bo_scan_kernel");
    bo_scatter_kernel<<<grid, BO_THREADS, 0, st>>>(codes, N, B, cursor, perm);
    SYM_LAUNCH_CHECK("bo_scatter_kernel");
    return SYM_OK;
}

extern "C" int sym_moe_stats(const float *Z, int32_t ldz, const float *R, int64_t N, int32_t d, int32_t K,
                             const int32_t *codes, int32_t V, const int32_t *var_offset, int32_t v,
                             int32_t n_levels_v, const int32_t *perm, const int64_t *seg,
                             const int32_t *chunk_start, int32_t B1, double *gram, double *cross,
                             sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(N >= 0 && d > 0 && K > 0 && V > 0 && v >= 0 && v < V && n_levels_v > 0 && B1 > 1,
                SYM_ERR_INVALID_ARGUMENT, "sym_moe_stats: bad sizes");
    SYM_REQUIRE((perm == nullptr) == (seg == nullptr) && (perm == nullptr) == (chunk_start == nullptr),
                SYM_ERR_INVALID_ARGUMENT, "sym_moe_stats: perm, seg and chunk_start must be given together");
    SYM_REQUIRE(perm != nullptr || n_levels_v == 1, SYM_ERR_INVALID_ARGUMENT,
                "sym_moe_stats: a variable with several levels needs sym_batch_order output");
    if (N == 0) return SYM_OK;
    const int KP = (K + 7) & ~7, DP = ((d + 1) + 3) & ~3;
    const size_t smem = 2 * (size_t)(ST_SUB * KP + ST_SUB * DP) * sizeof(float) + MOE_CH * sizeof(int32_t);
    SYM_REQUIRE(smem <= 200 * 1024, SYM_ERR_UNSUPPORTED_SHAPE, "sym_moe_stats: K=%d too large", K);
    const bool narrow = d + 1 <= 32;
    SYM_CUDA(cudaFuncSetAttribute(narrow ? moe_stats_kernel<true> : moe_stats_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SYM_REQUIRE(V <= MOE_MAX_VARS, SYM_ERR_UNSUPPORTED_SHAPE, "sym_moe_stats: more than %d batch variables", MOE_MAX_VARS);
    VarOffsets vo;
    for (int i = 0; i < MOE_MAX_VARS; ++i) vo.off[i] = i < V ? var_offset[i] : 0;
    const int64_t grid = ceil_div(N, MOE_CH) + (perm ? n_levels_v : 0);
    if (narrow)
        moe_stats_kernel<true><<<(unsigned)grid, ST_THREADS, smem, (cudaStream_t)stream>>>(
            Z, ldz, R, N, d, K, codes, V, vo, v, n_levels_v, perm, seg, chunk_start, B1, gram, cross);
    else
        moe_stats_kernel<false><<<(unsigned)grid, ST_THREADS, smem, (cudaStream_t)stream>>>(
            Z, ldz, R, N, d, K, codes, V, vo, v, n_levels_v, perm, seg, chunk_start, B1, gram, cross);
    SYM_LAUNCH_CHECK("moe_stats_kernel");
    return SYM_OK;
}

extern "C" int sym_moe_solve(const double *gram, const double *cross, const double *Nr, const double *C,
                             const double *lamb, int32_t K, int32_t B1, int32_t d, int32_t n_levels_var0, float *W,
                             int32_t *info, sym_stream_t stream) {
    using namespace sym;
    SYM_REQUIRE(K > 0 && B1 > 1 && d > 0 && n_levels_var0 > 0 && n_levels_var0 < B1, SYM_ERR_INVALID_ARGUMENT,
                "sym_moe_solve: bad sizes");
    const size_t need = ((size_t)B1 * B1 + (size_t)B1 * d) * sizeof(double);
    SYM_REQUIRE(need <= 200 * 1024, SYM_ERR_UNSUPPORTED_SHAPE,
                "sym_moe_solve: B+1=%d d=%d needs %zu B of shared memory (limit 200 KiB)", B1, d, need);
    SYM_CUDA(cudaFuncSetAttribute(moe_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    moe_solve_kernel<<<K, SV_THREADS, need, (cudaStream_t)stream>>>(gram, cross, Nr, C, lamb, K, B1, d, n_levels_var0,
                                                                    W, info);
    SYM_LAUNCH_CHECK("moe_solve_kernel");
    return SYM_OK;
}

extern "C" int sym_moe_apply(const float *Zin, int32_t ldzin, const float *R, int64_t N, int32_t d, int32_t K,
                             const int32_t *codes_v, int32_t var_offset, int32_t n_levels_v, const int32_t *perm,
                             const int64_t *seg, const int32_t *chunk_start, const float *W, int32_t B1,
                             float *Zout, int32_t ldzout, sym_stream_t stream) {
    using namespace sym;
    (void)codes_v;
    SYM_REQUIRE(N >= 0 && d > 0 && K > 0 && n_levels_v > 0 && B1 > 1, SYM_ERR_INVALID_ARGUMENT,
                "sym_moe_apply: bad sizes");
    SYM_REQUIRE((perm == nullptr) == (seg == nullptr) && (perm == nullptr) == (chunk_start == nullptr),
                SYM_ERR_INVALID_ARGUMENT, "sym_moe_apply: perm, seg and chunk_start must be given together");
    SYM_REQUIRE(perm != nullptr || n_levels_v == 1, SYM_ERR_INVALID_ARGUMENT,
                "sym_moe_apply: a variable with several levels needs sym_batch_order output");
    if (N == 0) return SYM_OK;
    const int DP = (d + 3) & ~3, KP = (K + 3) & ~3;
    const bool narrow = d <= 32;
    const size_t smem = ((size_t)KP * DP + 2 * (narrow ? 128 : 64) * (AP_KS + 4)) * sizeof(float) + MOE_CH * sizeof(int32_t);
    SYM_REQUIRE(smem <= 200 * 1024, SYM_ERR_UNSUPPORTED_SHAPE, "sym_moe_apply: K=%d d=%d too large", K, d);
    SYM_CUDA(cudaFuncSetAttribute(narrow ? moe_apply_kernel<true> : moe_apply_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = ceil_div(N, MOE_CH) + (perm ? n_levels_v : 0);
    if (narrow)
        moe_apply_kernel<true><<<(unsigned)grid, AP_THREADS, smem, (cudaStream_t)stream>>>(
            Zin, ldzin, R, N, d, K, var_offset, n_levels_v, perm, seg, chunk_start, W, B1, Zout, ldzout);
    else
        moe_apply_kernel<false><<<(unsigned)grid, AP_THREADS, smem, (cudaStream_t)stream>>>(
            Zin, ldzin, R, N, d, K, var_offset, n_levels_v, perm, seg, chunk_start, W, B1, Zout, ldzout);
    SYM_LAUNCH_CHECK("moe_apply_kernel");
    return SYM_OK;
}
