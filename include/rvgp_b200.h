/* rvgp_b200.h -- C ABI of librvgp_b200.so: hand-written sm_100a kernels for RVGP's
 * geometry-and-spectral hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain pointers and sizes (no torch / numpy
 * types).  Unless stated otherwise every pointer is a DEVICE pointer, matrices are row-major FP64,
 * indices are int32, calls are asynchronous on the handle's stream, and no ownership is transferred:
 * the caller allocates inputs, outputs and workspaces (sizes given by the *_workspace_bytes queries).
 *
 * The reference's only FFI is the CPython extension `ptu_dijkstra`
 * (RVGP/lib/ptu_dijkstra.pyx: tangent_frames :33, connections :128, internal C signatures
 * _geodesic_neigborhood_tangents :300-311 and _parallel_transport_dijkstra :221-229); the rest of its
 * path calls NumPy/SciPy/sklearn/GPflow.  Each function below names the reference code it replaces.
 *
 * Return value: 0 (RVGP_OK) or a negative status; rvgp_last_error(h) gives the message.
 */
#ifndef RVGP_B200_H
#define RVGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rvgp_handle_t;

#define RVGP_OK 0
#define RVGP_ERR_BAD_ARG (-1)        /* <-> ValueError, ptu_dijkstra.pyx:68-82,156-160 */
#define RVGP_ERR_RANK_DEFICIENT (-2) /* <-> return code -1 / RuntimeError, pyx:119-123,427-428 */
#define RVGP_ERR_CUDA (-3)
#define RVGP_ERR_NCCL (-4)
#define RVGP_ERR_NOT_SPD (-5)        /* Cholesky breakdown (tf.linalg.cholesky raises InvalidArgumentError) */
#define RVGP_ERR_CAPACITY (-6)       /* caller-provided workspace too small */

/* ---- handle ------------------------------------------------------------------------------------ */
int rvgp_create(int device, rvgp_handle_t* out);
int rvgp_destroy(rvgp_handle_t h);
int rvgp_set_stream(rvgp_handle_t h, void* cuda_stream);
const char* rvgp_last_error(rvgp_handle_t h);
int rvgp_sm_count(rvgp_handle_t h);
/* kernels launched through this handle since creation (bench.py "gpu_launches") */
unsigned long long rvgp_launch_count(rvgp_handle_t h);
int rvgp_version(void);
/* tuning knobs for experiments: "spmm_lpr" = lanes per block row in rvgp_bsr_spmm_f64 (0 auto | 8 | 16 | 32);
 * "dgemm_dmma" = FP64 GEMM kernel family (0 DFMA register tile | 1 DMMA, register-staged | 2 (default) DMMA with the
 * cp.async-pipelined 128x64 tile for large products and the 32x32 tile for small ones); "mma_variant", "mma_variant_n2",
 * "mma_prefetch", "mma_stream_policy", "mma_gpw" = schedule of the MMA SpMM; "gp_graph" = CUDA-graph replay of K15b. */
int rvgp_set_option(rvgp_handle_t h, const char* key, int value);

/* ---- K9: block-CSR SpMM (replaces the ARPACK mat-vec inside geometry.py:73) ------------------------
 * Y = alpha * (A @ X) + beta * X + gamma * W      (the Chebyshev three-term step when beta,gamma != 0)
 * A: BSR, nbrows block rows, d x d FP64 blocks (row-major inside a block) in `vals` (may be NULL for
 * the scalar graph Laplacian pattern mode: d == 1, value = deg_i on the diagonal entry and -1 elsewhere,
 * geometry.py:61).  X, W, Y: (nbrows*d) x ncols row-major with leading dimensions ldx/ldw/ldy.
 * W may be NULL when gamma == 0.  ncols in [1, 64].  Y must not alias X or W. */
int rvgp_bsr_spmm_f64(rvgp_handle_t h, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                      const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw,
                      double* Y, int64_t ldy, int ncols, double alpha, double beta, double gamma);

/* rvgp_bsr_spmm_f64 restricted to the block rows in rowlist (nlist entries); the other rows of Y are untouched.
 * Recomputes the boundary rows of a row-sharded matrix once the halo exchange has landed (the full launch overlaps it). */
int rvgp_bsr_spmm_rows_f64(rvgp_handle_t h, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                           const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y,
                           int64_t ldy, int ncols, double alpha, double beta, double gamma, const int32_t* rowlist,
                           int nlist);

/* ROT2 storage for d == 2: when every block is a scaled 2x2 orthogonal matrix [[a,-s*b],[b,s*a]] (true for the
 * connection Laplacian: Procrustes blocks, diagonal deg*I) it is stored as (a,b) with s in the sign bit of the column
 * index -- half the matrix bytes and half the value loads.  bad_flag (device int32): bit0 set if some block does not
 * fit within rtol.  Use the outputs with rvgp_bsr_spmm_f64 / rvgp_cheb_filter_f64 by passing d = -2. */
int rvgp_bsr_compress_rot2(rvgp_handle_t h, int64_t nnzb, const double* vals, const int32_t* indices, double* ab,
                           int32_t* idx_flag, int32_t* bad_flag, double rtol);

/* Chebyshev filter of degree `degree` applied in place to the ncols columns of V (scaled three-term
 * recurrence, Zhou & Saad): damps [lo_cut, hi] and amplifies below lo_cut, normalised at `lo_spec`.
 * work0, work1: two (nrows x ncols) scratch block vectors with leading dimension ldw.
 * The whole degree loop runs inside this one call (degree SpMM launches on the handle's stream);
 * on return the filtered block is in V. */
int rvgp_cheb_filter_f64(rvgp_handle_t h, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                         const double* vals, double* V, int64_t ldv, double* work0, double* work1,
                         int64_t ldw, int ncols, int degree, double lo_spec, double lo_cut, double hi);

/* ---- row-group merge plan (input of the K9 v4 k-step plan below).  R (4 or 8) consecutive block rows are grouped and
 * the UNION of their sorted column lists is formed (uent: int32 pairs (column, R-bit row mask); gptr: ngroups+1 offsets).
 * rvgp_bsr_merge_plan is called twice: first with uent == NULL (fills gptr; gptr[ngroups] = number of union entries),
 * then with uent allocated.  (The measured-and-rejected K9 v2 / v3 kernels that also consumed such plans live in
 * tools/experiments/, outside the product library.) */
int rvgp_bsr_merge_plan(rvgp_handle_t h, int nbrows, const int32_t* indptr, const int32_t* indices, int R,
                        int32_t* gptr, int32_t* uent, void* workspace, int64_t workspace_bytes);
int64_t rvgp_bsr_merge_plan_workspace_bytes(int nbrows, int R);

/* ---- K9 v4: row-group SpMM on the FP64 tensor path (mma.sync.m8n8k4.f64), d in {1, 2}.  Groups of 8/d consecutive
 * block rows are the 8 MMA rows; the union of their column lists (rvgp_bsr_merge_plan with R = 8/d) is cut into k-steps of
 * 4/d columns.  rvgp_bsr_mma_pack writes, per k-step, the column indices (kcols) and the dense 8x4 slice of A in
 * fragment order (afrag, 32 doubles, zero where a row does not store the column); kptr (ngroups+1) is the exclusive
 * prefix of the k-steps per group, computed by the caller from gptr.  vals == NULL (d == 1) packs the unit-weight graph
 * Laplacian.  rvgp_bsr_spmm_mma_f64 has the contract of rvgp_bsr_spmm_f64; it needs ncols % 16 == 0, 32-byte aligned
 * X / W / Y and leading dimensions % 4 == 0.  rvgp_cheb_filter_mma_f64 = rvgp_cheb_filter_f64 on top of it. */
int rvgp_bsr_mma_pack(rvgp_handle_t h, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                      const double* vals, const int32_t* gptr, const int32_t* uent, const int32_t* kptr, int32_t* kcols,
                      double* afrag);
int rvgp_bsr_spmm_mma_f64(rvgp_handle_t h, int nbrows, int d, const int32_t* kptr, const int32_t* kcols,
                          const double* afrag, const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y,
                          int64_t ldy, int ncols, double alpha, double beta, double gamma);
/* PATTERN plan (rotc == 2 below): the scalar unit-weight graph Laplacian (vals == NULL, d == 1; geometry.py:55-63) runs on the
 * native kernel as L (x) I_2 -- a row-major (n x B) scalar panel IS a node-contiguous d = 2 panel with B/2 columns -- with no
 * matrix values streamed at all: column words carry the 4-bit row mask of the R = 4 merge plan in bits 27..30 and the diagonal
 * comes from deg[i] = row length - 1 (passed as `afrag`).  kptr as for d = 2 (ceil(ulen / 2) k-steps per group of 4 rows). */
int rvgp_bsr_mma_pack_pattern(rvgp_handle_t h, int nbrows, const int32_t* indptr, const int32_t* gptr, const int32_t* uent,
                              const int32_t* kptr, int32_t* kcols, int32_t* deg, int32_t* bad_flag);
/* node-contiguous panels (d == 2): element (2*node+q, 2*cp+e) at Xn[node*ns + (cp*2+q)*2 + e]; beta is folded into the
 * diagonal as beta/alpha (alpha != 0).  rvgp_bsr_mma_rotc compacts the plan when every block is a scaled rotation /
 * reflection (16 doubles + sign bits per k-step; bad_flag != 0: not applicable, re-pack kcols); pass rotc = 1 then.
 * reverse != 0 walks the row groups backwards (alternate between the steps of a recurrence for L2 reuse). */
int rvgp_bsr_spmm_mma_native_f64(rvgp_handle_t h, int nbrows, const int32_t* kptr, const int32_t* kcols,
                                 const double* afrag, int rotc, const double* Xn, int64_t nsx, const double* Wn,
                                 int64_t nsw, double* Yn, int64_t nsy, int ncols, double alpha, double beta, double gamma,
                                 int reverse);
int rvgp_bsr_mma_rotc(rvgp_handle_t h, int64_t nk, const double* afrag, int32_t* kcols, double* afrag_c,
                      int32_t* bad_flag, double rtol);
int rvgp_panel_native_f64(rvgp_handle_t h, int to_native, int nbrows, int ncols, double* V, int64_t ldv, double* Xn,
                          int64_t ns);
/* d == 2 and work2 != NULL: the recurrence runs on node-contiguous panels (work0..2 contiguous, ldw == ncols), V is
 * converted on entry / exit; otherwise the row-major MMA kernel is used and work2 / rotc must be NULL / 0. */
int rvgp_cheb_filter_mma_f64(rvgp_handle_t h, int nbrows, int d, const int32_t* kptr, const int32_t* kcols,
                             const double* afrag, int rotc, double* V, int64_t ldv, double* work0, double* work1,
                             double* work2, int64_t ldw, int ncols, int degree, double lo_spec, double lo_cut, double hi);

/* ---- Row-sharded operator with the halo exchange over NVLink peer memory (csrc/halo.cu) ------------
 * The extended block vectors E[0..2] ((n_loc + n_halo) * d rows x ncols, contiguous; node-contiguous panels when an MMA
 * plan is given) and the flag array live in memory allocated with rvgp_ipc_alloc and opened by the peers with
 * rvgp_ipc_open (CUDA IPC; the 64-byte handles travel through torch.distributed).  pull_src[s][i] = peer-mapped address
 * of the row of halo node i in its owner's E[s]; peer_slots[j] = peer-mapped address of THIS rank's slot in neighbour j's
 * flag array; wait_idx[j] = slot of neighbour j in this rank's flag array.  Every call only enqueues kernels on the
 * handle's stream (signal / wait / pull / SpMM per degree); epochs must grow monotonically and identically on all ranks:
 * a filter consumes epoch0 .. epoch0 + degree, a product epoch0 .. epoch0 + 1.  err (device int32) is set when a wait
 * times out (timeout_ms) -- the results are then undefined but nothing hangs. */
typedef struct rvgp_halo_ctx {
    int32_t n_loc, n_halo, d, ncols;
    const int32_t* indptr;          /* local block rows, columns renumbered to [local | halo] */
    const int32_t* indices;
    const double* vals;             /* NULL: unit-weight graph Laplacian pattern (d == 1) */
    const int32_t* kptr;            /* optional MMA plan of the local matrix (d == 2), else NULL */
    const int32_t* kcols;
    const double* afrag;
    int32_t rotc;
    int32_t n_peers;
    double* E[3];
    const int64_t* pull_src[3];
    uint64_t* flags;
    uint64_t** peer_slots;
    const int32_t* wait_idx;
    int32_t* err;
    int32_t timeout_ms;
    int32_t n_interior;             /* MMA plans only: row groups (4 block rows) that read local rows only ... */
    const int32_t* glist_interior;
    const int32_t* glist_boundary;  /* ... and those that also read halo rows.  Both NULL: one launch over all groups after the */
    int32_t n_boundary;             /* exchange.  Given: signal -> interior groups -> wait -> pull -> boundary groups, i.e. the */
    int32_t reserved;               /* neighbours' skew and the pull hide behind the interior launch. */
} rvgp_halo_ctx;
int rvgp_ipc_alloc(rvgp_handle_t h, int64_t bytes, void** dptr, uint8_t* handle64);
int rvgp_ipc_open(rvgp_handle_t h, const uint8_t* handle64, void** dptr);
int rvgp_ipc_close(rvgp_handle_t h, void* dptr);
int rvgp_ipc_free(rvgp_handle_t h, void* dptr);
int rvgp_halo_barrier(rvgp_handle_t h, const rvgp_halo_ctx* ctx, uint64_t epoch);
int rvgp_halo_spmm_f64(rvgp_handle_t h, const rvgp_halo_ctx* ctx, uint64_t epoch0, const double* X, int64_t ldx, double* Y,
                       int64_t ldy);
int rvgp_halo_cheb_filter_f64(rvgp_handle_t h, const rvgp_halo_ctx* ctx, uint64_t epoch0, double* V, int64_t ldv, int degree,
                              double lo_spec, double lo_cut, double hi);

/* ---- K10: dense FP64 kernels for orthogonalisation / Rayleigh-Ritz ---------------------------------
 * C (m x n, ldc) = alpha * op(A) * op(B).  Layout flags say which index of each operand is contiguous:
 *   a_kmajor = 0: A(i,k) = A[k*lda + i]  ("A is stored as K x M", e.g. V^T of a tall block vector)
 *   a_kmajor = 1: A(i,k) = A[i*lda + k]
 *   b_kmajor = 0: B(k,j) = B[k*ldb + j]
 *   b_kmajor = 1: B(k,j) = B[j*ldb + k]
 * scale_k (nullable, length K) multiplies A(i,k) by scale_k[k] (the spectral density S in
 * kernels.py:61).  When split_k > 1 the K range is split over split_k CTAs per tile, partial tiles go
 * to `workspace` (split_k*m*n doubles) and are reduced in a fixed order (deterministic). */
int rvgp_dgemm_f64(rvgp_handle_t h, int m, int n, int64_t k, double alpha, const double* A, int64_t lda,
                   int a_kmajor, const double* B, int64_t ldb, int b_kmajor, const double* scale_k,
                   double* C, int64_t ldc, int split_k, double* workspace);
int64_t rvgp_dgemm_workspace_bytes(int m, int n, int split_k);
/* symmetric results (Gram matrices): only tiles intersecting the lower triangle are computed; upper part undefined */
int rvgp_dgemm_lower_f64(rvgp_handle_t h, int m, int n, int64_t k, double alpha, const double* A, int64_t lda,
                         int a_kmajor, const double* B, int64_t ldb, int b_kmajor, double* C, int64_t ldc, int split_k,
                         double* workspace);

/* column-wise reductions over tall block vectors (deterministic two-stage) ---------------------------
 * out[c] = sum_r A[r*lda+c] * B[r*ldb+c]   (B == NULL: B = 1, i.e. column sums).
 * workspace: rvgp_coldot_workspace_bytes */
int rvgp_coldot_f64(rvgp_handle_t h, int64_t nrows, int ncols, const double* A, int64_t lda,
                    const double* B, int64_t ldb, double* out, double* workspace);
/* out[c] = sum_r (W[r,c] - theta[c] * V[r,c])^2    (squared residual norms of Ritz pairs; V == NULL: V = 1) */
int rvgp_resid_sq_f64(rvgp_handle_t h, int64_t nrows, int ncols, const double* W, int64_t ldw,
                      const double* V, int64_t ldv, const double* theta, double* out, double* workspace);
int64_t rvgp_coldot_workspace_bytes(int64_t nrows, int ncols);
/* A[r,c] *= s[c] */
int rvgp_colscale_f64(rvgp_handle_t h, int64_t nrows, int ncols, double* A, int64_t lda, const double* s);
/* deterministic counter-based uniform(-1,1) fill: element (r,c) depends only on (seed, row_offset+r, col_offset+c),
 * so a row-sharded block vector is the same global vector for any number of ranks */
int rvgp_fill_uniform_f64(rvgp_handle_t h, int64_t nrows, int ncols, double* A, int64_t lda,
                          uint64_t seed, int64_t col_offset, int64_t row_offset);
/* out[r, c] = in[perm[r], c] for r < nrows (row gather; used for the locality permutation) */
int rvgp_gather_rows_f64(rvgp_handle_t h, int64_t nrows, int ncols, const double* in, int64_t ldin,
                         const int32_t* perm, int block, double* out, int64_t ldout);

/* ---- K2: exact kNN (replaces sklearn kneighbors_graph, geometry.py:103-110) -------------------------
 * X (n, D) all candidates; queries are rows [q_begin, q_begin+q_count) (row sharding across GPUs).
 * out_idx (q_count, k) int32 in ascending (distance, index) order, self excluded by index;
 * out_d2 (nullable) squared distances.  Distances are sum_j (x_j-y_j)^2 in coordinate order, no FMA. */
int rvgp_knn_f64(rvgp_handle_t h, const double* X, int n, int D, int q_begin, int q_count, int k,
                 int32_t* out_idx, double* out_d2);

/* grid-accelerated exact kNN for D <= 3: identical outputs (bit for bit) to rvgp_knn_f64, candidates pruned through a
 * uniform cell grid; lo/hi are HOST arrays (D) with the bounding box of X. */
int rvgp_knn_grid_f64(rvgp_handle_t h, const double* X, int n, int D, const double* lo, const double* hi, int q_begin,
                      int q_count, int k, int32_t* out_idx, double* out_d2, void* workspace, int64_t workspace_bytes);
int64_t rvgp_knn_grid_workspace_bytes(int n, int D, const double* lo, const double* hi);

/* ---- K3: kNN lists -> symmetric CSR with self loops (geometry.py:111-112, ptu_dijkstra.pyx:84-103) ----- */
int rvgp_knn_to_csr(rvgp_handle_t h, const int32_t* knn, int n, int k, int32_t* indptr, int32_t* indices,
                    int32_t* nnz_out, void* workspace, int64_t workspace_bytes);
int64_t rvgp_knn_to_csr_workspace_bytes(int n, int k);
/* typ='affinity' graph of manifold_graph (geometry.py:114-118): dense Gaussian-kernel weights
 * A[i][j] = exp(-dist(i,j)^2 / (2 sigma^2)) with sklearn's pairwise_distances expansion, diagonal distance forced to 0.
 * A: n x n row-major; workspace: n doubles; n <= 65535. */
int rvgp_affinity_f64(rvgp_handle_t h, const double* X, int n, int D, double sigma, double* A, double* workspace);

/* locality (Morton) ordering of the points and symmetric permutation of a CSR pattern (eigensolver layout) */
int rvgp_morton_order(rvgp_handle_t h, const double* X, int n, int D, int32_t* order, int32_t* inv,
                      void* workspace, int64_t workspace_bytes);
int64_t rvgp_morton_order_workspace_bytes(int n);
int rvgp_csr_permute(rvgp_handle_t h, int n, const int32_t* indptr, const int32_t* indices, const int32_t* order,
                     const int32_t* inv, int32_t* new_indptr, int32_t* new_indices, void* workspace,
                     int64_t workspace_bytes);
int64_t rvgp_csr_permute_workspace_bytes(int n);

/* ---- K4: geodesic neighbourhoods = literal Fibonacci-heap Dijkstra (ptu_dijkstra.pyx:361-394, 444-690) --
 * seq (n, K+1) popped ids in pop order; counts (n); flags: device int32 (bit0 short component -> stale tail
 * reproduced, bit1 decrease_val would have fired, bit2 node-pool overflow). */
int rvgp_geodesic_neighbourhoods(rvgp_handle_t h, const int32_t* indptr, const int32_t* indices, int n, int K,
                                 int maxdeg, int32_t* seq, int32_t* counts, int32_t* flags, void* workspace,
                                 int64_t workspace_bytes);
int64_t rvgp_geodesic_workspace_bytes(rvgp_handle_t h, int n, int K, int maxdeg);
/* multi-GPU form: only the sources [src_begin, src_begin + src_count) (rows of the full-size seq / counts), no stale-tail pass;
 * all-gather the rows, OR the flags, then call rvgp_geodesic_fix_stale once on the complete arrays (pyx:350 copies the stale
 * tail of a short component from the PREVIOUS source's row). */
int rvgp_geodesic_neighbourhoods_range(rvgp_handle_t h, const int32_t* indptr, const int32_t* indices, int n, int K,
                                       int maxdeg, int src_begin, int src_count, int32_t* seq, int32_t* counts,
                                       int32_t* flags, void* workspace, int64_t workspace_bytes);
int rvgp_geodesic_fix_stale(rvgp_handle_t h, int n, int K, int32_t* seq, const int32_t* counts, const int32_t* flags);

/* ---- K5/K6: tangent frames (ptu_dijkstra.pyx:396-434) and dimension statistic (geometry.py:83-97) ------ */
int rvgp_tangent_frames(rvgp_handle_t h, const double* X, int n, int D, const int32_t* seq, int Kp1, int dcheck,
                        double* tangents, double* Sigma, int32_t* flag);
int rvgp_sigma_cumvar(rvgp_handle_t h, const double* Sigma, int n, int D, double* cum);
int rvgp_slice_frames(rvgp_handle_t h, const double* T, int64_t n, int D, int dfull, int d, double* G);

/* ---- K7/K8: Procrustes connections + connection-Laplacian assembly (ptu_dijkstra.pyx:259-293,
 * geometry.py:35-42).  For every stored CSR entry e=(i,j) incl. the diagonal: R = U V^T of svd(T_i^T T_j);
 * Lc block = deg_i * R on the diagonal entry, -R elsewhere (deg_i = row length - 1).
 * gauges (n, D, d); Lc_vals, R_vals (each nullable): (nnzb, d, d) in CSR entry order. */
int rvgp_connections(rvgp_handle_t h, const double* gauges, int n, int D, int d, const int32_t* indptr,
                     const int32_t* indices, int64_t nnzb, double* Lc_vals, double* R_vals);

/* initial block for the Lc eigensolver built from scalar-Laplacian eigenvectors U (n x kL):
 * out[(i*d+q), c] = gauges[i][c % D][q] * U[i][c / D]  (no reference counterpart; ARPACK starts from a random vector) */
int rvgp_lift_guess(rvgp_handle_t h, const double* gauges, int64_t n, int D, int d, const double* U, int64_t ldu, int kL,
                    double* out, int64_t ldo, int ncols);

/* ---- K11/a15: frame contractions ------------------------------------------------------------------------
 * mode 0: out(n,d)   = G^T x      express_in_local_frame          (geometry.py:171-176)
 * mode 1: out(n,D)   = G x        express_in_local_frame(reverse) / eigenvector lift (dataclass.py:57-59)
 * x and out carry `ncols` trailing columns: x (n, d|D, ncols), out (n, D|d, ncols). */
int rvgp_frame_apply(rvgp_handle_t h, const double* gauges, int64_t n, int D, int d, const double* x, double* out,
                     int ncols, int mode, double scale);

/* ---- K1: furthest-point sampling (geometry.py:126-162), one persistent cooperative kernel ---------------
 * N > 0: exactly N samples; N == 0: until lambdas[i]/diam < spacing (diam = max pairwise distance, computed
 * here with sklearn's euclidean_distances expansion).  perm/lambdas capacity: N, or n when N == 0.  D <= 1024; for D > 64
 * (feature-space sampling of SGPR inducing points, main.py:60) one warp owns a point and lanes stride over coordinates. */
int rvgp_fps_f64(rvgp_handle_t h, const double* X, int n, int D, int N, double spacing, int start_idx,
                 int32_t* perm, double* lambdas, int32_t* count_out, void* workspace, int64_t workspace_bytes);
int64_t rvgp_fps_workspace_bytes(rvgp_handle_t h, int n);

/* ---- K12 helpers: vector-diffusion smoothing (smoothing.py:37-64), random field (dataclass.py:91-103) ---
 * Y += a X ; out[i] = ||x[i,:]|| ; x[i,:] *= out_abs[i] / (ind[i] * ||x[i,:]||)  (out_abs/ind nullable = 1) */
int rvgp_axpy_f64(rvgp_handle_t h, int64_t nrows, int ncols, double a, const double* X, int64_t ldx, double* Y,
                  int64_t ldy);
int rvgp_row_norms_f64(rvgp_handle_t h, int64_t n, int d, const double* x, double* out);
int rvgp_renorm_rows_f64(rvgp_handle_t h, int64_t n, int d, double* x, const double* out_abs, const double* ind);

/* C = alpha * op(A) op(B) + beta * C (layout flags as rvgp_dgemm_f64; no split-K, no K scaling) */
int rvgp_dgemm_acc_f64(rvgp_handle_t h, int m, int n, int64_t k, double alpha, const double* A, int64_t lda,
                       int a_kmajor, const double* B, int64_t ldb, int b_kmajor, double beta, double* C, int64_t ldc);

/* ---- K13/K14: GP solve (replaces GPflow GPR's TF ops reached from main.py:55-58,77,80,111) -------------
 * rvgp_potrf_f64: in-place lower Cholesky (blocked, right-looking) of an SPD row-major matrix; the strict upper
 *   triangle is scratch.  flag (device int32): bit0 set when a pivot is not positive (-> RVGP_ERR_NOT_SPD at
 *   the caller, like tf.linalg.cholesky failing).  workspace: rvgp_potrf_workspace_bytes(n); it keeps the
 *   inverses of the diagonal blocks for rvgp_trsm_f64.
 * rvgp_trsm_f64: solve L X = B (trans 0) or L^T X = B (trans 1) in place, B (n x nrhs) row-major;
 *   scratch: 64*nrhs doubles.
 * rvgp_kdiag_f64: K_diag[i] = sum_j S[j] X[i,j]^2 (kernels.py:63-67) without forming the N* x N* matrix. */
int rvgp_potrf_f64(rvgp_handle_t h, double* A, int64_t lda, int n, int32_t* flag, void* workspace,
                   int64_t workspace_bytes);
int64_t rvgp_potrf_workspace_bytes(int n);
int rvgp_trsm_f64(rvgp_handle_t h, const double* L, int64_t ldl, int n, double* B, int64_t ldb, int nrhs, int trans,
                  const void* potrf_workspace, double* scratch);
int rvgp_add_diag_f64(rvgp_handle_t h, double* A, int64_t lda, int n, double v);
int rvgp_logdiag_sum_f64(rvgp_handle_t h, const double* A, int64_t lda, int n, double* out);
int rvgp_kdiag_f64(rvgp_handle_t h, const double* X, int64_t ldx, int64_t n, int k, const double* S, double* out);

/* ---- K15b: one rank-k GP evaluation for k > 64 enqueued as a whole (build B = I + S^1/2 G S^1/2 / noise, blocked Cholesky,
 * ONE forward-substitution kernel over 8-column slices of [S^1/2 G | S^1/2 b] that also reduces the column norms, ONE
 * back-substitution / log-determinant / packing kernel; the panel-by-panel rvgp_trsm_f64 route remains for k > 3200), captured
 * once per fit and replayed as a CUDA graph: the L-BFGS-B loop of train_gp (main.py:87-95) then costs one small
 * host->device copy (par = [S (k), noise]), this call and one device->host copy of out per evaluation.
 * out (2 + 2k doubles): [not-SPD flag, sum log L_ii, z = B^-1 (S^1/2 b), qs_j = ||L^-1 (S^1/2 G) e_j||^2]. */
int rvgp_gp_lowrank_eval_f64(rvgp_handle_t h, int k, const double* G, const double* b, const double* par, double* out,
                             void* workspace, int64_t workspace_bytes);
int64_t rvgp_gp_lowrank_eval_workspace_bytes(int k);

/* ---- K16: fused rank-k GP evaluation for small k (<= 64), one CTA per problem (many time frames per launch; the
 * EEG-shaped workload of examples/eeg_example/eeg_utils.py:26-35,107-112).  Inputs per problem: G = Phi^T Phi (k x k),
 * b = Phi^T y, yy = y^T y, Mrows, spectral density s (k), noise.  Strides (in doubles) of 0 share an array between
 * problems.  out per problem: [lml, dLML/dnoise, dLML/dS (k), posterior mean weights (k), Q = L_b^-1 S^1/2 (k*k)
 * when want_predict].  lml is NaN when B is not positive definite. */
int rvgp_gp_lowrank_small_f64(rvgp_handle_t h, int nprob, int k, const double* G, int64_t g_stride, const double* b,
                              int64_t b_stride, const double* yy, const double* Mrows, const double* s, int64_t s_stride,
                              const double* noise, double* out, int64_t out_stride, int want_predict);

/* ---- K17: squared-exponential kernel for train_gp(kernel='rbf') (main.py:33-37 -> gpflow.kernels.RBF())
 * and the dense pieces of the SGPR path (main.py:59-67,119-137 -> gpflow.models.SGPR).
 * rvgp_rbf_from_dot_f64: K[i,j] = variance * exp(-0.5 * (xa2[i] + xb2[j] - 2 P[i,j]) / lengthscale^2), P = XA XB^T from
 *   rvgp_dgemm_f64, xa2/xb2 = squared row norms (rvgp_kdiag_f64 with S = 1).  Kout may alias P.
 * rvgp_rbf_adjoint_f64: reverse mode of the same map for an adjoint Gbar (m x n): H = Gbar o K (Hout nullable, may alias
 *   Gbar or P), sums[0] = sum H, sums[1] = sum H * r2 (r2 scaled by 1/l^2), rowsum[i] = sum_j H[i,j] (nullable).
 *   dF/dvariance = sums[0] / variance, dF/dlengthscale = sums[1] / lengthscale.  Deterministic two-stage reduction.
 * rvgp_rbf_dx_f64: out = (HX - rowsum o XA) / l^2 = dF/dXA (gradient w.r.t. inducing points).
 * rvgp_scale_shift_f64: A = alpha * A + beta * I (rectangular allowed). */
int rvgp_rbf_from_dot_f64(rvgp_handle_t h, int m, int n, const double* P, int64_t ldp, const double* xa2,
                          const double* xb2, double variance, double lengthscale, double* Kout, int64_t ldk);
int rvgp_rbf_adjoint_f64(rvgp_handle_t h, int m, int n, const double* Gbar, int64_t ldg, const double* P, int64_t ldp,
                         const double* xa2, const double* xb2, double variance, double lengthscale, double* Hout,
                         int64_t ldh, double* rowsum, double* sums, void* workspace, int64_t workspace_bytes);
int64_t rvgp_rbf_adjoint_workspace_bytes(int m, int n);
int rvgp_rbf_dx_f64(rvgp_handle_t h, int64_t m, int k, const double* HX, int64_t ldhx, const double* rowsum,
                    const double* XA, int64_t ldx, double lengthscale, double* out, int64_t ldo);
int rvgp_scale_shift_f64(rvgp_handle_t h, int64_t nrows, int ncols, double alpha, double beta, double* A, int64_t lda);

/* ---- K19: divergence / curl features of a vector field on a 3-D point cloud (examples/eeg_example/eeg_utils.py:46-80,
 * compute_vectorfield_features).  knn (n x k, int32): neighbour lists from rvgp_knn_f64 / rvgp_knn_grid_f64
 * (self excluded, ascending distance = KDTree.query order).  ref_row0 = 1 reproduces the reference's use of
 * normalized_vectors[0] as the subtracted vector (eeg_utils.py:67); 0 subtracts the point's own vector.
 * div (n), curl (n x 3); both averaged over the k neighbours. */
int rvgp_vectorfield_features_f64(rvgp_handle_t h, int n, int k, const double* positions, const double* vectors,
                                  const int32_t* knn, int ref_row0, double* div, double* curl);

/* ---- K20: consistent orientation of d = 2 gauges -> complex-Hermitian form of the connection Laplacian (no reference
 * counterpart; the eigenpairs of D Lc D, D = diag(1, s_i), are those of Lc = compute_connection_laplacian, geometry.py:14-52,
 * up to the sign flip D, and every eigenvalue is exactly double -- the pairs eigsh returns at geometry.py:73).
 * rvgp_orient_steps: `steps` sweeps of pull-style label propagation over the block-CSR pattern; labels (n, int32): 0 =
 *   unlabelled, +-1 = node sign; seed at least one node before the first call; *changed is set when any label was written.
 * rvgp_orient_check: out2[0] = number of stored off-diagonal blocks with s_i s_j det(block) < 0, out2[1] = unlabelled nodes.
 * rvgp_rot90_nodes_f64: out = J V, (J V)[2i] = -V[2i+1], (J V)[2i+1] = V[2i] (out must not alias V).
 * rvgp_flip_odd_rows_f64: V[2i+1, :] *= s[i] (apply D to a block vector, or to the second gauge coordinate).
 * rvgp_pair_panel_f64: out (2n x 2b) = [V | J V]; rvgp_pair_combine_f64: W = beta W + alpha (T[:, :b] + J T[:, b:2b]) -- the two
 *   halves of the complex block algebra of the paired eigensolver, so that C^H W and V C are ONE real GEMM each (krylov.py). */
int rvgp_orient_steps(rvgp_handle_t h, int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                      int32_t* labels, int32_t* changed, int steps);
int rvgp_orient_check(rvgp_handle_t h, int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                      const int32_t* labels, int32_t* out2);
int rvgp_rot90_nodes_f64(rvgp_handle_t h, int64_t nnodes, int ncols, const double* V, int64_t ldv, double* out, int64_t ldo);
int rvgp_flip_odd_rows_f64(rvgp_handle_t h, int64_t nnodes, int ncols, const int32_t* s, double* V, int64_t ldv);
int rvgp_pair_panel_f64(rvgp_handle_t h, int64_t nnodes, int ncols, const double* V, int64_t ldv, double* out, int64_t ldo);
int rvgp_pair_combine_f64(rvgp_handle_t h, int64_t nnodes, int ncols, const double* T, int64_t ldt, double alpha, double beta,
                          double* W, int64_t ldw);

#ifdef __cplusplus
}
#endif
#endif /* RVGP_B200_H */
