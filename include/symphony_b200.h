/*
 * symphony_b200 — C ABI of the B200 (sm_100a) query-mapping kernels.
 *
 * Drop-in boundary for the hot path of potulabe/symphonypy:
 *   symphonypy/tl.py:271-397   map_embedding      (project -> assign -> MoE correct)
 *   symphonypy/tl.py:400-436   transfer_labels_kNN (brute Euclidean kNN + majority vote)
 * The reference is pure Python (numpy / scikit-learn); it has no FFI of its own, so the
 * entry points below are the numerical call sites of that path, one export per reference
 * call cluster, each citing the lines it replaces.  `INTEGRATION.md` shows the ctypes
 * binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host"
 *   - all matrices are row-major, float32 unless stated; sizes in elements
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     allocates nothing, and returns 0 on success or a non-zero code
 *     (SYM_ERR_* below, or 1000 + cudaError_t); `sym_last_error()` gives the text
 *   - no C++ exceptions cross the boundary; no torch types appear in it
 */
#ifndef SYMPHONY_B200_H
#define SYMPHONY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYM_OK 0
#define SYM_ERR_INVALID_ARGUMENT 1
#define SYM_ERR_UNSUPPORTED_SHAPE 2
#define SYM_ERR_WORKSPACE_TOO_SMALL 3
#define SYM_ERR_CUDA_BASE 1000

typedef void *sym_stream_t;

/* Library version (major*10000 + minor*100 + patch). */
int sym_version(void);
/* Text of the last error raised on the calling host thread ("" if none). */
const char *sym_last_error(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t sym_launch_count(void);
/* 64-bit digest of `nbytes` of HOST memory, computed on up to `threads` host threads (the value does not depend on
 * `threads`).  Synchronous, no GPU work.  The host layer keys its device-side caches of reference objects (loadings,
 * packed kNN index) by the content of the arrays they were built from — the reference rebuilds them on every call
 * (symphonypy/tl.py:425-426 fits the classifier each time) — and re-hashes the arrays on every call: a foreign-function
 * call releases the interpreter lock, so the hashing overlaps the caller's other host work.  Not cryptographic. */
uint64_t sym_host_digest64(const void *data, uint64_t nbytes, int32_t threads);
/* memcpy of HOST memory on up to `threads` host threads (dst and src must not overlap).  Synchronous, no GPU work.  The
 * host layer stages pageable input arrays into page-locked memory with it when the process has been left a single
 * intra-op thread (torchrun exports OMP_NUM_THREADS=1): the staging copy is what bounds a pageable input then. */
void sym_host_copy(void *dst, const void *src, uint64_t nbytes, int32_t threads);
/* Compute capability (major*10+minor) the kernels were compiled for: 100. */
int sym_compiled_arch(void);

/* ---------------------------------------------------------------------------------------
 * (1) project — replaces symphonypy/_utils.py:298-308 (scale, clip, GEMM with loadings).
 *   Z[n,:] = sum_g clip((X[n,col(g)] - mu[g]) * inv_sd[g], -clip, +clip) * U[g,:]
 *   X        [N, ldx]  query expression; column col(g) holds reference gene g
 *   gene_col [G] or NULL   NULL: col(g) = g.  Entries < 0 mark genes absent from the query;
 *                      together with inv_sd[g] == 0 (absent or zero-std genes,
 *                      _utils.py:276) they contribute exactly 0 (_utils.py:240-243).
 *   U        [G, ldu]  loadings, ldu >= d, ldu % 4 == 0, columns d..ldu-1 zero
 *   Z        [N, ldz]  ldz >= d.   d <= 128.
 */
int sym_project_dense(const float *X, int64_t N, int64_t ldx, const int32_t *gene_col, int32_t G,
                      const float *mu, const float *inv_sd, const float *U, int32_t ldu, int32_t d,
                      float clip, float *Z, int32_t ldz, sym_stream_t stream);

/* Tensor-core form of the same projection for a query whose genes are already in reference order
 * (gene_col == NULL) and whose rows are 16-byte aligned (ldx % 4 == 0): tcgen05 MMAs with FP16x3 split operands
 * (hi + lo fp16 pairs, FP32-class results), bound by streaming X rather than by FFMA.
 *   sym_project_pack: packs u_scale * U [G, ldu] once into `packed` (sym_project_pack_bytes(G, d) bytes,
 *   16-byte aligned): per 32 genes a hi and a lo fp16 tile in the UMMA K-major layout.  u_scale: a power of two
 *   that brings max |U| into [512, 1024) (keeps the lo parts in the normal fp16 range).
 *   sym_project_dense_tc: out_scale = 1 / u_scale is applied to the accumulator.  clip <= 60000 (fp16 range).
 */
size_t sym_project_pack_bytes(int32_t G, int32_t d);
int sym_project_pack(const float *U, int32_t ldu, int32_t G, int32_t d, float u_scale, void *packed,
                     sym_stream_t stream);
int sym_project_dense_tc(const float *X, int64_t N, int64_t ldx, int32_t G, const float *mu, const float *inv_sd,
                         const void *packed, int32_t d, float clip, float out_scale, float *Z, int32_t ldz,
                         sym_stream_t stream);

/* CSR query (scipy layout) — replaces _utils.py:240-243,284-285 (densify) + :298-308.
 *   Z[n,:] = z0 + sum_{nnz j of row n, gene g=col_gene[indices[j]] >= 0}
 *                   (clip((data[j]-mu[g])*inv_sd[g]) - clip(-mu[g]*inv_sd[g])) * U[g,:]
 *   z0[d] = sum_g clip(-mu[g]*inv_sd[g]) * U[g,:]  is written by sym_project_csr_bias.
 *   col_gene [n_query_cols]  reference gene index of each query column, or -1 if unused.
 */
int sym_project_csr_bias(const float *mu, const float *inv_sd, const float *U, int32_t ldu, int32_t G,
                         int32_t d, float clip, float *z0, sym_stream_t stream);
int sym_project_csr(const float *data, const int32_t *indices, const int64_t *indptr, int64_t N,
                    const int32_t *col_gene, const float *mu, const float *inv_sd, const float *U,
                    int32_t ldu, int32_t d, float clip, const float *z0, float *Z, int32_t ldz,
                    sym_stream_t stream);
/* Same, and *not_canonical (device int32, zeroed by the caller) is set to 1 when some row's column indices are not
 * strictly increasing: duplicated columns must be summed before the scale + clip (the reference densifies first,
 * _utils.py:284-285), so the caller canonicalises the matrix and calls again.  not_canonical may be NULL. */
int sym_project_csr_checked(const float *data, const int32_t *indices, const int64_t *indptr, int64_t N,
                            const int32_t *col_gene, const float *mu, const float *inv_sd, const float *U,
                            int32_t ldu, int32_t d, float clip, const float *z0, float *Z, int32_t ldz,
                            int32_t *not_canonical, sym_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (2) assign — replaces symphonypy/_utils.py:154-178.
 *   zc = sign(max_j Z[n,j]) * Z[n,:] / ||Z[n,:]||     (signed row max: _utils.py:168-170)
 *   R[n,k] = softmax_k( -2 (1 - Y[k,:].zc) * inv_sigma[k] )
 *   Y [K,d] unit-norm centroids (tl.py:361-362), inv_sigma [K], R [N,K].  K <= 1024, d <= 128.
 */
int sym_assign(const float *Z, int32_t ldz, int64_t N, int32_t d, const float *Y, int32_t K,
               const float *inv_sigma, float *R, sym_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (3) MoE correction — replaces symphonypy/_utils.py:181-214.  The design matrix Phi of
 * tl.py:368-378 is one-hot per batch variable, so it is passed as integer level codes.
 * Design column of level c of variable v is  col = 1 + var_offset[v] + c  (column 0 = intercept).
 *
 * sym_batch_order: counting sort of the cells by level code of ONE variable.
 *   codes [N] in [0,B);  perm [N] cell ids grouped by level;  seg [B+1] int64 group offsets;
 *   chunk_start [B+1] int32 prefix of ceil(group/256) (work list of the stats/apply kernels).
 *   workspace: >= sym_batch_order_workspace(B) bytes.
 */
size_t sym_batch_order_workspace(int32_t B);
int sym_batch_order(const int32_t *codes, int64_t N, int32_t B, int32_t *perm, int64_t *seg,
                    int32_t *chunk_start, void *workspace, size_t workspace_bytes, sym_stream_t stream);

/* sym_moe_stats: accumulate (+=) the sufficient statistics of _utils.py:196-204 for variable v:
 *   gram [K,B1,B1] f64:  gram[k,c,c] += sum_{n in level} R[n,k]
 *                         gram[k,c,c'] and gram[k,c',c] += R[n,k] for the level c' of every
 *                         later variable v' > v of the same cell
 *   cross[K,B1,d]  f64:  cross[k,c,:] += sum_{n in level} R[n,k] Z[n,:]
 *   (row / column 0 and the priors Nr, C are filled in by sym_moe_solve, AFTER any
 *    cross-GPU all-reduce of gram and cross — _utils.py:201,205 add them once, globally.)
 *   codes [V,N] all variables;  var_offset [V] host array;  perm/seg/chunk_start: order of
 *   variable v from sym_batch_order, or perm == NULL when variable v has a single level.
 */
int sym_moe_stats(const float *Z, int32_t ldz, const float *R, int64_t N, int32_t d, int32_t K,
                  const int32_t *codes, int32_t V, const int32_t *var_offset_host, int32_t v,
                  int32_t n_levels_v, const int32_t *perm, const int64_t *seg, const int32_t *chunk_start,
                  int32_t B1, double *gram, double *cross, sym_stream_t stream);

/* sym_moe_solve: W_k = inv(gram_k + diag(lamb)) cross_k with row 0 zeroed (_utils.py:201-209).
 *   Builds row/column 0 (intercept) from the levels of variable 0, adds Nr[k] to gram[k,0,0] and
 *   C[k,:] to cross[k,0,:], factorises (Cholesky, float64, in shared memory) and solves.
 *   Requires ((B1*B1 + B1*d) * 8 bytes <= 200 KiB.
 *   lamb [B1] f64 (lamb[0] == 0);  W [K,B1,d] f32;  info [K] int32: 0 ok, j>0 = pivot j not positive.
 */
int sym_moe_solve(const double *gram, const double *cross, const double *Nr, const double *C,
                  const double *lamb, int32_t K, int32_t B1, int32_t d, int32_t n_levels_var0,
                  float *W, int32_t *info, sym_stream_t stream);

/* sym_moe_apply: Zout[n,:] = Zin[n,:] - sum_k R[n,k] W[k, 1+var_offset+code[n], :] (_utils.py:212)
 *   for ONE variable; call once per variable, chaining Zout -> Zin (in place allowed).
 */
int sym_moe_apply(const float *Zin, int32_t ldzin, const float *R, int64_t N, int32_t d, int32_t K,
                  const int32_t *codes_v, int32_t var_offset, int32_t n_levels_v, const int32_t *perm,
                  const int64_t *seg, const int32_t *chunk_start, const float *W, int32_t B1, float *Zout,
                  int32_t ldzout, sym_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (4) kNN + vote — replaces tl.py:425-433 (sklearn KNeighborsClassifier, brute Euclidean,
 * uniform weights).
 *
 * sym_knn_fit (KNeighborsClassifier.fit): packs the reference embedding once.
 *   Ref [M,ldr] -> packed (opaque, sym_knn_packed_bytes(M,d) bytes).
 *   order [M] or NULL: a permutation of the reference rows (packed position -> row).  Any permutation gives
 *   the same results; one that keeps spatially close rows together (rows sorted by nearest of a few hundred
 *   centroids, see sym_nearest_centroid) makes the scan several times faster.
 * sym_nearest_centroid: code[n] = index of the nearest of C centroids [C,d] (squared Euclidean, lowest index
 *   on ties) — the coarse spatial key used to order reference rows and query rows alike.
 * sym_knn (kneighbors): idx [N,k] int32 ascending by (squared distance, index), dist2 [N,k] f32.
 *   Candidates are selected by a float32 (or tensor-core) scan over-selecting k+pad per row and
 *   re-ranked in float64 with sklearn's formula ||x||^2 - 2x.y + ||y||^2 clamped at 0; ties go
 *   to the lower reference index.  k <= 120, d <= 128, k <= M.
 *   method: 0 = auto (2 when the shape fits, else 1), 1 = FP32 FFMA scan, 2 = tcgen05 scan with BF16x3 split
 *   operands (FP32-class scores), 3 = tcgen05 scan with single FP16 operands (a third of the MMAs; its score
 *   error is bounded per row from the exact rounding residuals, so more rows may fail the certificate).
 *   unsafe [N] uint8 or NULL: rows whose over-selection margin could not be certified
 *   (the caller re-runs those rows with the next more accurate method: 3 -> 2 -> 1).
 * sym_knn_method_supported: 1 if sym_knn accepts this shape with this method, else 0.
 *   Optional scan order (all three or none): q_perm [N] = query rows sorted by coarse cluster code,
 *   q_code [N] = cluster code of every query row, cl_tile [C] = first 128-row tile of the packed reference
 *   that holds rows of cluster c.  Results do not depend on them.
 * sym_vote (predict / predict_proba): label[n] = lowest class code among the most frequent
 *   codes of the k neighbours; conf[n] = winning votes / k.  ref_codes [M]; entries of idx outside [0, M) are
 *   ignored (label -1 if a row has no valid neighbour).  dist2 NULL: weights="uniform"; dist2 [N,k] (squared
 *   distances from sym_knn): weights="distance" (1/distance, sklearn's zero-distance rule), conf = winning weight share.
  * Exact tile pruning (opt-in): OR `SYM_KNN_PRUNE` into `method` (2 or 3).  sym_knn_fit stores a bounding ball for
 *   every 128-row tile of the packed reference; a scan CTA (256 query rows) then skips the tiles whose ball cannot hold
 *   a candidate of any of its rows:  max(0, |c_q - c_t| - r_q - r_t)^2 >= max over its rows of (threshold + ||q||^2).
 *   Results are identical to the full scan (every skipped reference is provably farther than the row's threshold;
 *   the certificate of the re-rank is unchanged).  The first two uint64 of `workspace` then hold
 *   {tiles scanned, tiles a full scan visits}, summed over the scan's CTAs.  Data dependent: nothing is skipped for
 *   unclustered data.
*/
size_t sym_knn_packed_bytes(int64_t M, int32_t d);
int sym_knn_fit(const float *Ref, int32_t ldr, int64_t M, int32_t d, const int32_t *order, void *packed,
                sym_stream_t stream);
int sym_nearest_centroid(const float *X, int32_t ldx, int64_t N, int32_t d, const float *centroids, int32_t C,
                         int32_t *code, sym_stream_t stream);
size_t sym_knn_workspace(int64_t N, int64_t M, int32_t d, int32_t k, int32_t method);
int sym_knn_method_supported(int64_t N, int64_t M, int32_t d, int32_t k, int32_t method);
#define SYM_KNN_PRUNE 0x100
int sym_knn(const float *Q, int32_t ldq, int64_t N, const float *Ref, int32_t ldr, int64_t M,
            const void *packed, int32_t d, int32_t k, int32_t method, const int32_t *q_perm, const int32_t *q_code,
            const int32_t *cl_tile, int32_t *idx, float *dist2, uint8_t *unsafe, void *workspace,
            size_t workspace_bytes, sym_stream_t stream);
/* Exact float64 brute force for a FEW query rows (n_rows <= sym_knn_exact_rows_max() = 64), the cheap way to finish
 * the rows sym_knn flagged `unsafe` when there are only a handful of them (a scan launched for 4 rows costs ~0.8 ms, a
 * reading of the reference ~0.03 ms).  rows [n_rows] int64: indices into Q.  ub [n_rows] f32: an upper bound on each
 * row's k-th squared distance — e.g. the k-th distance sym_knn returned for that row (uncertified results are still exact
 * distances of real reference rows), nudged up; +inf is allowed but useless.  idx_out / dist2_out [n_rows,k] as sym_knn
 * (same formula, same tie rule).  overflow [n_rows] uint8 = 1 for rows whose bound admitted more than 1024 reference rows
 * (or fewer than k: not a bound): nothing usable was written, send them through sym_knn with the next method. */
int32_t sym_knn_exact_rows_max(void);
size_t sym_knn_exact_rows_workspace(int32_t n_rows);
int sym_knn_exact_rows(const float *Q, int32_t ldq, const int64_t *rows, int32_t n_rows, const float *Ref, int32_t ldr,
                       int64_t M, int32_t d, int32_t k, const float *ub, int32_t *idx_out, float *dist2_out,
                       uint8_t *overflow, void *workspace, size_t workspace_bytes, sym_stream_t stream);
int sym_vote(const int32_t *idx, const float *dist2, int64_t N, int32_t k, const int32_t *ref_codes, int64_t M,
             int32_t *label, float *conf, sym_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (5) per-cell mapping confidence — replaces symphonypy/tl.py:88-111 and _utils.py:468-496
 *     (next row after the hot path).
 * sym_cluster_moments: sums [K, d+2] f64 = {sum_m R_km, sum_m R_km^2, sum_m R_km x_m[0..d)} over the
 *   reference (X [M,ldx], R [K,M] as uns["harmony"]["R"] stores it, float32).
 * sym_cluster_cov: cov [K,d,d] f64 = sum_m R_km (x_m - mean_k)(x_m - mean_k)^T  (mean [K,d] f32; d <= 64).
 *   The caller applies the v1/(v1^2 - v2) factor of _utils.py:493 and inverts the K small matrices.
 * sym_cell_confidence: out[n] = sum_k Rq[n,k] * sqrt((x_n - mean_k)^T inv_cov_k (x_n - mean_k))   (d <= 64).
 */
int sym_cluster_moments(const float *X, int32_t ldx, int64_t M, int32_t d, const float *R, int32_t K, double *sums,
                        sym_stream_t stream);
int sym_cluster_cov(const float *X, int32_t ldx, int64_t M, int32_t d, const float *R, int32_t K, const float *mean,
                    double *cov, sym_stream_t stream);
int sym_cell_confidence(const float *Xq, int32_t ldq, int64_t N, int32_t d, const float *Rq, int32_t K,
                        const float *mean, const float *inv_cov, float *out, sym_stream_t stream);

/* k-means++ seeding of the no-Harmony fallback — replaces sklearn's greedy k-means++ inside
 * KMeans(init="k-means++") at symphonypy/_utils.py:323-325 (sklearn/cluster/_kmeans.py, _kmeans_plusplus).
 *   centre_ids[0] = first;  for c = 1..K-1: `trials` candidate rows are drawn with probability proportional to the
 *   squared distance to the nearest centre so far — searchsorted(cumsum(closest), rand[c-1][t] * sum(closest)) — and
 *   the candidate that lowers the total potential most becomes centre c (first one on ties).
 *   rand [K-1, trials] f64 uniform draws in [0,1) (device);  centre_ids [K] i32 out (device): rows of X.
 *   Distances are ||x||^2 - 2 x.c + ||c||^2 in float64 (sklearn's formula).  N < 2^31, d <= 128, trials <= 16.
 */
size_t sym_kmeanspp_workspace(int64_t N, int32_t trials);
int sym_kmeanspp(const float *X, int32_t ldx, int64_t N, int32_t d, int32_t K, int32_t trials, int64_t first,
                 const double *rand, int32_t *centre_ids, void *workspace, size_t workspace_bytes, sym_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (6) per-cluster mapping confidence — replaces the per-group numpy of symphonypy/_utils.py:422-465
 *     (tl.per_cluster_confidence, tl.py:114-179).  Cells are grouped by a user-given
 *     cluster code with sym_batch_order (perm, seg, chunk_start; all NULL when G == 1).
 * sym_group_sums: sums [G,d] f64 = column sums of the members of every group (d <= 128).
 * sym_group_cov:  cov [G,d,d] f64 = sum over members of (x - mean_g)(x - mean_g)^T   (mean [G,d] f32; d <= 64).
 *   The caller divides by (n_g - 1), adds lamb*I, inverts the G small matrices and evaluates
 *   _utils.py:446-463 on them.
 */
int sym_group_sums(const float *X, int32_t ldx, int64_t N, int32_t d, int32_t G, const int32_t *perm,
                   const int64_t *seg, const int32_t *chunk_start, double *sums, sym_stream_t stream);
int sym_group_cov(const float *X, int32_t ldx, int64_t N, int32_t d, int32_t G, const int32_t *perm,
                  const int64_t *seg, const int32_t *chunk_start, const float *mean, double *cov,
                  sym_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (7) test hook (no reference counterpart): the FP32 scores S[q,r] = ||r||^2 - 2 q.r of ONE 128 x 128 tile as the
 *     tcgen05 scan computes them — same packing, descriptors and MMA chain — so that tests/test_gpu_parity.py can
 *     compare the tensor-core arithmetic with float64.  Q [nq<=128, d], Ref [nr<=128, d] -> out [128,128];
 *     scratch >= 3 packed tiles.  flags bit 0: query operand in tensor memory; bit 1: single-FP16 operands.
 */
int sym_debug_tc_scores(const float *Q, int32_t nq, const float *Ref, int32_t nr, int32_t d, float *out, void *scratch,
                        sym_stream_t stream, int32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* SYMPHONY_B200_H */
