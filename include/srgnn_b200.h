/*
 * srgnn_b200.h — C ABI of libsrgnn_b200.so, the B200 (sm_100a) implementation of the
 * K-hop feature-propagation path of yyysyyy/Scalable-Roubust-GNN.
 *
 * Reference interfaces replaced (paths relative to "Scalable Spectral Robust GNN/" = SSRG/):
 *   - SSRG/operators/csrc/matmul.h:5        void FloatCSRMulDenseOMP(...)   (the CPU SpMM that runs today)
 *   - SSRG/operators/csrc/cudamatmul.c:28   int  FloatCSRMulDense(...)      (dead cuSPARSE sibling)
 *   - SSRG/operators/utils.py:17-47         csr_sparse_dense_matmul         (ctypes marshalling around the former)
 *   - SSRG/operators/utils.py:81-93         adj_to_symmetric_norm           (scipy fp64 normalisation)
 *   - SSRG/operators/base_operator.py:19-36 GraphOp.propagate               (K-hop loop)
 *   - wavelet/src/utils.py:89-104,125-138   WaveletSparsifier               (Chebyshev heat filter + threshold)
 *   - SSRG/data_process.py:35-67, SSRG/data_augument.py:28,99-102           (mask application / edge gather)
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - every entry point returns 0 on success or a negative errno-style code; srg_last_error()
 *     returns a thread-local human readable message for the last failure.
 *   - "device" entry points take DEVICE pointers and a cudaStream_t passed as void* (NULL = legacy
 *     default stream); they never synchronise the device and never allocate unless stated.
 *   - "host" entry points take HOST pointers (pinned or pageable), do H2D / kernels / D2H
 *     themselves on library-owned streams, and return after the results are in host memory.
 *   - feature matrices are row-major fp32 with an explicit leading dimension (in floats).
 *     The fast path needs ld % 4 == 0 and 16-byte aligned base pointers (128-bit loads);
 *     anything else takes the scalar kernels.
 *   - CSR: int32 indptr (n+1) and int32 indices, exactly what scipy hands the reference.
 *     Products n*ld are formed in 64 bit (the reference's int offsets overflow at N*F >= 2^31,
 *     SSRG/operators/csrc/matmul.c:29,33).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SRG_ERR_NODEV.
 */
#ifndef SRGNN_B200_H
#define SRGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRG_ABI_VERSION 1

#define SRG_OK 0
#define SRG_ERR_INVALID (-22) /* EINVAL: bad argument                     */
#define SRG_ERR_NOMEM (-12)   /* ENOMEM: device / pinned allocation failed */
#define SRG_ERR_CUDA (-5)     /* EIO:    CUDA runtime / kernel failure     */
#define SRG_ERR_NODEV (-19)   /* ENODEV: no CUDA device                    */
#define SRG_ERR_RANGE (-34)   /* ERANGE: size exceeds an int32 CSR limit   */
#define SRG_ERR_UNSUPPORTED (-95) /* ENOTSUP: input needs a path not built yet (see message) */

/* dtype tags for the adjacency value array */
#define SRG_VAL_ONES 0 /* data == NULL, every stored entry is 1.0 (unweighted graph) */
#define SRG_VAL_F32 1
#define SRG_VAL_F64 2
#define SRG_VAL_HAS_ZEROS 0x100 /* OR into val_dtype: the value array may hold explicit zeros (compacting kernels) */

/* bits reported in the int32 flags word written by the normalisation entry points */
#define SRG_FLAG_UNSORTED 1       /* input rows not strictly increasing (duplicates / unsorted)      */
#define SRG_FLAG_ASYMMETRIC 2     /* pattern of (A+I) not symmetric: general transpose path required */
#define SRG_FLAG_ZERO_PRODUCT 4   /* a normalised value is exactly 0: scipy would drop the entry     */
#define SRG_FLAG_BAD_INDEX 8      /* a column index is outside [0, n)                                */
#define SRG_FLAG_WEIGHTED 16      /* informational: some stored value differs from 1.0              */
#define SRG_FLAG_EXPLICIT_ZERO 32  /* a stored value is exactly 0: retry with SRG_VAL_HAS_ZEROS          */

/* ---- library / diagnostics ------------------------------------------------------------- */
int srg_abi_version(void);
const char *srg_last_error(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int srg_device_count(void);
/* counts kernels launched by this library in the calling process (bench "gpu_launches") */
int64_t srg_launch_count(void);

/* tuning / experiment knobs (profiles/ records which values the defaults come from):
 *   "spmm_variant" 0 group kernel | 1 stream kernel;  "stream_batch" 4|8;  "stream_rows" rows per
 *   warp task;  "group_unroll" 4|8;  "gather_l2_64" 0|1;  "long_row" split threshold (0 = never);
 *   "bulk_gather" -1 auto | 0 never | 1 always (rows <= 512 B fetched by TMA bulk copies), "bulk_auto" / "bulk_min"
 *   widest / narrowest padded row (floats) of the automatic choice, "bulk_stages" 2..4, "bulk_rows" rows per warp
 *   task, "bulk_tile" 0|1 finished rows leave as one bulk store per destination;  "push_tma" 0|1 bulk-store tile
 *   epilogue of the LDGSTS push hop;  "push_rows_tma" 0|1 input exchange through the TMA unit, "push_rows_tma_blocks"
 *   / "push_rows_blocks" its grid / the grid of the st.global form;  "exact_sym_check" 0|1 per-entry mirror lookup
 *   instead of the hashed symmetry check of the normalisation.  SRG_TUNE="key=value,..." sets them from the
 *   environment when the Python package loads the library. */
int srg_set_tuning(const char *key, int64_t value);

/* ---- a3: adjacency normalisation  (SSRG/operators/utils.py:81-93) ------------------------ */
/*
 * Stage 1.  Structure of A~ = A + I over a canonical CSR (sorted rows, no duplicates).
 *   out_indptr[n+1] : row pointer of A~ (explicit zeros and entries cancelled by +I are dropped,
 *                     as scipy's csr_plus_csr does)
 *   out_count[n]    : nullable; entries per row of A~ (= integer degree incl. the self loop for
 *                     an unweighted graph)
 *   out_flags       : device int32, OR of SRG_FLAG_* (caller zeroes it)
 *   nnz             : upper bound of indptr[n] (sizes the hub-row segment list without a readback)
 * Rows longer than 1024 entries (power-law hubs) are processed as 1024-entry segments by separate
 * warps.  The fast kernels require a canonical CSR without explicit zeros and raise
 * SRG_FLAG_UNSORTED / SRG_FLAG_EXPLICIT_ZERO otherwise (retry after srg_csr_canonicalize / with
 * val_dtype | SRG_VAL_HAS_ZEROS).  Allocates scratch with cudaMallocAsync on `stream`.
 */
int srg_degree_selfloop_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                            int val_dtype, int64_t n, int64_t nnz, int32_t *out_indptr,
                            int32_t *out_count, int32_t *out_flags, void *stream);

/*
 * Stage 2.  R = D^(r-1) A~^T D^(-r)   (R[a,b] = (A~[b,a] * d_a^(r-1)) * d_b^(-r), fp64, this
 * multiply order) and optionally the PPR blend (1-alpha)*R + alpha*I
 * (SSRG/operators/graph_operator/symmetrical_simgraph_ppr_operator.py:19-20) when ppr_alpha >= 0.
 *   nnz                        : stored entries of A (capacity of the outputs is nnz + n)
 *   out_indices[out_indptr[n]] : column indices of R, sorted per row
 *   out_degree[n]              : nullable; d = A~.sum(1) in fp64, summed in numpy's add.reduceat
 *                                order (first element + 8-lane pairwise sum of the rest)
 *   out_val_f64 / out_val_f32  : either may be NULL; f32 = round-to-nearest of the fp64 value,
 *                                i.e. the `adj.data.astype(np.float32)` of utils.py:39
 * When stage 1 flagged the input (UNSORTED / BAD_INDEX / EXPLICIT_ZERO) nothing is computed and
 * out_indptr is overwritten with zeros (an empty matrix), so hops launched before the flags are
 * read gather nothing.
 * Symmetric-pattern path: sets SRG_FLAG_ASYMMETRIC (outputs then undefined) when A~ has an entry
 * (a,b) without (b,a); SRG_FLAG_ZERO_PRODUCT when a value is exactly 0 (scipy drops those).
 * Allocates scratch with cudaMallocAsync on `stream`.
 */
int srg_sym_norm_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                     int val_dtype, int64_t n, int64_t nnz, int32_t *out_indptr, double r,
                     double ppr_alpha, int32_t *out_indices, double *out_degree,
                     double *out_val_f64, float *out_val_f32, int32_t *out_flags, void *stream);

/*
 * The same two stages on a ROW SLICE [row0, row0 + n_rows) of an n_cols-column adjacency (column ids
 * stay global): the building blocks of the row-partitioned multi-GPU normalisation.
 *   srg_selfloop_rows_csr      stage 1 on the slice (diagonal = global row id)
 *   srg_selfloop_fill_rows_csr pattern of A~ for the slice, its values (written only when the slice
 *                              is weighted, see SRG_FLAG_WEIGHTED) and the slice's degrees
 *   srg_pow_tables_f64         d^(r-1), d^(-r) with inf -> 0 over a (gathered) degree vector
 *   srg_norm_values_rows_csr   R values of the slice from the GLOBAL power tables.  check_symmetry
 *                              (row0 must be 0: every row local) verifies A~ == A~^T exactly (mirror
 *                              lookup of every upper entry + upper/lower counts) and raises
 *                              SRG_FLAG_ASYMMETRIC otherwise; without it symmetry is the caller's promise.
 *                              nnz = allocated capacity (entries) of at_indices / at_val / outputs.
 *                              If stage 1 flagged the input, at_indptr is zeroed (empty matrix).
 */
int srg_selfloop_rows_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                          int val_dtype, int64_t n_rows, int64_t nnz, int64_t row0, int64_t n_cols,
                          int32_t *out_indptr, int32_t *out_count, int32_t *out_flags, void *stream);
int srg_selfloop_fill_rows_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                               int val_dtype, int64_t n_rows, int64_t nnz, int64_t row0,
                               int64_t n_cols, const int32_t *at_indptr, int32_t *at_indices,
                               double *at_val, double *out_degree, const int32_t *flags, void *stream);
int srg_pow_tables_f64(const double *degree, int64_t n, double r, double *out_left,
                       double *out_right, void *stream);
int srg_norm_values_rows_csr(int32_t *at_indptr, const int32_t *at_indices,
                             const double *at_val, const double *degree_rows, int64_t n_rows,
                             int64_t nnz, int64_t row0, int64_t n_cols, const double *pow_left,
                             const double *pow_right,
                             double ppr_alpha, int check_symmetry, double *out_val_f64,
                             float *out_val_f32, int32_t *flags, void *stream);

/*
 * General path of stage 2 for a NON-symmetric pattern (directed input): explicit transpose by a
 * stable key sort.  Same outputs as srg_sym_norm_csr plus out_indptr (the row pointer of R, which
 * differs from that of A~); at_indptr is the stage-1 row pointer of A~.  Synchronises `stream`
 * once (4-byte readback of the entry count).
 */
int srg_sym_norm_csr_general(const int32_t *indptr, const int32_t *indices, const void *data,
                             int val_dtype, int64_t n, int64_t nnz, const int32_t *at_indptr,
                             double r, double ppr_alpha, int32_t *out_indptr, int32_t *out_indices,
                             double *out_degree, double *out_val_f64, float *out_val_f32,
                             int32_t *out_flags, void *stream);

/* ---- 8f-2: magnetic-Laplacian normalisation of a directed graph ----------------------------------------
 * adj_to_directed_symmetric_mag_norm (SSRG/operators/utils.py:95-138), the normaliser of
 * SymDirMagLaplacianGraphOp / SymDirMagComPprGraphOp: symmetrised weights (A + A^T) / 2 with self loops,
 * degrees, D^(r-1) A_s D^(-r), Hadamard exp(i * q_angle * (A - A^T)); q_angle = 2 pi q as the host
 * evaluates it.  Outputs one CSR pattern (sorted, union of A, A^T and the diagonal; capacity 2 nnz + n)
 * with the real and the imaginary values in fp64 (what the reference's scipy matrices hold) and/or
 * rounded to fp32 (what the hops read, utils.py:39); any value output may be NULL.
 * ppr_alpha >= 0 applies real = (1 - alpha) real + alpha I, imag = (1 - alpha) imag
 * (symmetrical_directed_magnetic_comppr_operator.py:33-38); pass a negative value for none.
 * nnz must equal indptr[n].  out_flags: SRG_FLAG_BAD_INDEX, SRG_FLAG_ZERO_PRODUCT (a blended real entry
 * became exactly 0, which scipy would drop). */
int srg_mag_norm_csr(const int32_t *indptr, const int32_t *indices, const void *data, int val_dtype,
                     int64_t n, int64_t nnz, double r, double q_angle, double ppr_alpha, int32_t *out_indptr,
                     int32_t *out_indices, double *out_degree, double *out_real_f64, double *out_imag_f64,
                     float *out_real_f32, float *out_imag_f32, int32_t *out_flags, void *stream);

/*
 * Sparse x sparse product C = A B (float32 CSR; A: n_a x k_dim, B: k_dim x n_b) by expand / sort / compress.
 * Replaces the DENSE N x N products of adj_to_un_in_out_dir_symmetric_norm (in_L = P^T P, out_L = P P^T,
 * SSRG/operators/utils.py:216-227) and torch_sparse.spspmm(Psi, Psi^-1)
 * (SSRG/models/base_scalable/base_model.py:208-214, wavelet/src/gwnn_layer.py:59-75).  c_ij is the sequential
 * fp32 sum of a_ik * b_kj in ascending stored order of k; NaN results become 0 (utils.py:221); drop_zeros != 0
 * removes exact zeros from the pattern (torch.nonzero, utils.py:223).  Rows of C come out sorted.
 * a_vals / b_vals NULL = all ones.  out_* capacity `cap` entries (SRG_ERR_RANGE when the product holds more or
 * when more than 2^31-1 intermediate products arise); *out_nnz (host) = entries.  Synchronises the stream.
 */
int srg_spgemm_csr_f32(const int32_t *a_indptr, const int32_t *a_indices, const float *a_vals, int64_t n_a,
                       int64_t k_dim, const int32_t *b_indptr, const int32_t *b_indices, const float *b_vals,
                       int64_t n_b, int32_t drop_zeros, int32_t *out_indptr, int32_t *out_indices,
                       float *out_vals, int64_t cap, int64_t *out_nnz, int32_t *out_flags, void *stream);
/* row i of the output = row i of the input + the entry (i, i) appended (torch_geometric add_self_loops as used
 * at SSRG/operators/utils.py:199-201: one loop per node, duplicates are NOT merged here).  out_indices
 * capacity indptr[n] + n. */
int srg_csr_append_diagonal(const int32_t *indptr, const int32_t *indices, int64_t n, int32_t *out_indptr,
                            int32_t *out_indices, void *stream);
/* out_vals[p] = (deg[i]^(r-1) * vals[p]) * deg[col_p]^(-r) in float32, deg = row sums in stored order, inf -> 0:
 * the float32 normalisation of SSRG/operators/utils.py:204-210, :229-237, :249-257.  vals NULL = all ones;
 * out_degree optional. */
int srg_csr_sym_scale_f32(const int32_t *indptr, const int32_t *indices, const float *vals, int64_t n, float r,
                          float *out_vals, float *out_degree, void *stream);

/* ---- 8f-2: fast PPR-approximation normaliser of a directed graph (SSRG/operators/utils.py:262-335) ---------
 * Device stages; the host loop (convergence test, <= 100 sweeps) and the float32 degree normalisation
 * (srg_csr_sym_scale_f32) are driven by the caller (operators/utils.py:adj_to_fast_ppr_approx_symmetric_norm).
 *   srg_ppr_iterate_f64  y = (1 - a) A1^T (x / r) + s (z . x) over the CSR of A1^T (float32 duplicate counts), fp64;
 *                        stats3 (device, 3 doubles) = { z . x, |y - x|^2, sum(y) }               (utils.py:284-293)
 *   srg_ppr_symmetrize   L = (Pi^1/2 P Pi^-1/2 + Pi^-1/2 P^T Pi^1/2) / 2 with pi = x / stats3[2], P = D1 A1, NaN -> 0,
 *                        values rounded to float32; outputs: sorted CSR, capacity 2 nnz              (utils.py:297-309)
 * Not yet validated on hardware (round 1): exercised by an opt-in test only. */
int srg_ppr_iterate_f64(const int32_t *t_indptr, const int32_t *t_indices, const float *t_counts,
                        const double *degree, int64_t n, double ppr_alpha, const double *x, double *y,
                        double *stats3, void *stream);
int srg_ppr_symmetrize(const int32_t *indptr, const int32_t *indices, const float *counts, const double *degree,
                       const double *x, const double *stats3, int64_t n, int64_t nnz, int32_t *out_indptr,
                       int32_t *out_indices, float *out_vals, void *stream);

/* Two-order PPR approximation (adj_to_slow_first_second_ppr_approx_symmetric_norm, SSRG/operators/utils.py:337-424):
 *   srg_teleport_iterate_f64    one power-iteration sweep for the left eigenvector of the (n + 1) x (n + 1) teleport
 *                               matrix [[(1 - a) P, a], [1/n, 0]] (the reference: dense LAPACK eig, :353-369) over the CSR
 *                               of P^T; x, y: n + 1 doubles; stats3 = { sum_{i<n} x_i, |y - x|_1 over the first n, sum_{j<n} y_j }
 *   srg_csr_intersect_mean_f32  (A + B) / 2 on the entries where both sorted float32 CSR operands are non-zero: the
 *                               in-place masking L_in[L_out == 0] = 0, L_out[L_in == 0] = 0 and the mean of :405-410;
 *                               out_indices / out_vals capacity a_nnz, out_indptr[n] = entries.
 * Not yet validated on hardware (round 1): exercised by an opt-in test only. */
int srg_teleport_iterate_f64(const int32_t *t_indptr, const int32_t *t_indices, const float *t_vals, int64_t n,
                             double ppr_alpha, const double *x, double *y, double *stats3, void *stream);
int srg_csr_intersect_mean_f32(const int32_t *a_indptr, const int32_t *a_indices, const float *a_vals,
                               const int32_t *b_indptr, const int32_t *b_indices, const float *b_vals, int64_t n,
                               int64_t a_nnz, int32_t *out_indptr, int32_t *out_indices, float *out_vals, void *stream);

/* ---- synthetic inputs of the named shapes, generated on the device --------------------
 * Not a reference interface: BASELINE.json's configs 4 (power-law variant) and 5 are synthetic R-MAT
 * graphs too large to build on the host per rank.  Rows [row0, row1) of the symmetrised, duplicate-free,
 * loop-free adjacency (pattern only; every value is 1) of an R-MAT(a, b, c, 1-a-b-c) graph with
 * 2^scale ids scrambled by a fixed bijection, ids >= n rejected, m_draw edge ids drawn.  Counter-based:
 * edge id e alone determines the pair, so every rank can produce its own shard and the union over
 * the ranks is one well-defined graph.  out_indptr: row1-row0+1 entries (local rows, GLOBAL column ids),
 * out_indices capacity `cap` (SRG_ERR_RANGE when the shard needs more); *out_nnz (host) = entries.
 * Synchronises the stream (set-up path).  numpy restatement: scalable_roubust_gnn_b200/synth.py. */
int srg_synth_rmat_shard_csr(uint64_t seed, int32_t scale, int64_t m_draw, double a, double b, double c,
                             int64_t n, int64_t row0, int64_t row1, int64_t cap, int32_t *out_indptr,
                             int32_t *out_indices, int64_t *out_nnz, void *stream);
/* out[i, c] (n_rows x ld, pad columns zero) = U[0,1) float32 from hash(seed, row0 + i, col0 + c) */
int srg_synth_hash_features_f32(uint64_t seed, int64_t row0, int64_t n_rows, int32_t col0, int32_t F,
                                int32_t f_total, float *out, int64_t ld, void *stream);

/*
 * Transpose of a float32 CSR (square, n x n): the matrix the backward pass of the per-epoch sparse
 * product needs (torch.mm(self.adj, x) in Layer2GraphConvolution.forward,
 * SSRG/models/base_scalable/simple_models.py:228,233: grad_x = adj^T grad_y).  Stable key sort, so a
 * canonical input gives a canonical output; `nnz` is the capacity of the arrays, the live entry count
 * is indptr[n] on the device.  vals NULL = all ones.  Outputs have capacity nnz (tail zero-filled).
 */
int srg_csr_transpose_f32(const int32_t *indptr, const int32_t *indices, const float *vals, int64_t n,
                          int64_t nnz, int32_t *out_indptr, int32_t *out_indices, float *out_vals,
                          int32_t *out_flags, void *stream);

/*
 * Canonical form of a CSR with unsorted rows and/or duplicate entries (what scipy's
 * `adj.tocoo() + eye` does first, SSRG/operators/utils.py:82): rows sorted by column, duplicates
 * summed in stored order.  Outputs have capacity nnz; *out_nnz_dev (device int32) = entries kept.
 */
int srg_csr_canonicalize(const int32_t *indptr, const int32_t *indices, const void *data,
                         int val_dtype, int64_t n, int64_t nnz, int32_t *out_indptr,
                         int32_t *out_indices, double *out_vals, int32_t *out_nnz_dev,
                         int32_t *out_flags, void *stream);

/* ---- a8: sparsity masks (SSRG/data_process.py:35-67, SSRG/data_augument.py:28,99-102) -------- */
/* out[:, i] = edge_index[:, keep[i]]  (int64 2 x E row-major in, 2 x E_keep out): data_process.py:66 */
int srg_edge_gather_i64(const int64_t *edge_index, int64_t E, const int64_t *keep, int64_t E_keep,
                        int64_t *out, int32_t *out_flags, void *stream);
/* undirected duplicate-free adjacency (pattern CSR, every value 1.0) of an edge list:
 * data_augument.py:99-102 (cat both directions, unique).  out_indices capacity 2E;
 * *out_nnz_dev (device int32) = stored entries. */
int srg_edges_to_sym_csr(const int64_t *edge_index, int64_t E, int64_t n, int32_t *out_indptr,
                         int32_t *out_indices, int32_t *out_nnz_dev, int32_t *out_flags,
                         void *stream);
/* x * feature_mask (data_augument.py:28) into a (possibly padded) device layout; alias of
 * srg_pack_features_f32 with a non-NULL mask */
int srg_apply_feature_mask_f32(const float *x, int64_t ld_x, const int32_t *mask, float *out,
                               int64_t ld_out, int64_t n, int32_t F, void *stream);

/* ---- a5/a6: one propagation hop  (SSRG/operators/csrc/matmul.c:23-40) --------------------- */
/*
 * Y[i, 0:F] = sum_j vals[j] * X[indices[j], 0:F]  for j in indptr[i]..indptr[i+1], evaluated
 * per output element as a sequential fp32 FMA chain in CSR order starting from 0.0f, which is
 * the arithmetic of the reference's vfmadd loop, so results are bit-identical to it.
 * Y is overwritten (the reference accumulates into a pre-zeroed buffer).
 * n_rows rows of the CSR are processed; column indices address rows of X (global ids).
 * nnz: an upper bound of indptr[n_rows] (sizes the long-row scratch without a device readback);
 * rows longer than 1024 entries (power-law hubs) are evaluated as fixed 1024-entry segments whose
 * sums are added in order: deterministic, bit-identical for every shorter row.  Pass nnz <= 0 to
 * force the strict in-order chain for every row.
 */
int srg_spmm_csr_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                     int64_t n_rows, int64_t nnz, const float *X, int64_t ldx, float *Y,
                     int64_t ldy, int32_t F, void *stream);

/* ---- (a9, config 3) heat-kernel wavelets of the IDENTITY impulse as sparse matrices ----------------------------
 * Psi_s = sum_k c_sk T_k(L~) I, thresholded (r < tol -> dropped; tol = NaN: no threshold), float32 CSR — what
 * WaveletSparsifier.calculate_wavelet (wavelet/src/utils.py:89-104) and SpectralModel.calculate_wavelet
 * (SSRG/models/base_scalable/base_model.py:236-265) obtain from pygsp's cheby_op on dense N x N / N x 1000 impulse
 * blocks.  The recurrence runs on stored entries only (sparse products L T_{k-1} by expand / stable sort / ordered
 * fp64 segment sums); bit-identical to srg_cheby_filter_f64 on identity blocks.  L: the CSR of srg_laplacian_csr
 * (fp64, device) and must store every diagonal entry (SRG_ERR_UNSUPPORTED otherwise: isolated nodes).
 * coeffs: HOST [n_scales][order + 1].  Set-up path: synchronises the stream.  Results live behind the handle:
 *   srg_cheby_sparse_info   nnz of one scale, total products expanded, nnz of the order-M pattern
 *   srg_cheby_sparse_fetch  device-to-device copy of one scale's CSR into caller arrays
 *   srg_cheby_sparse_free   releases the handle */
int srg_cheby_sparse_run(const int32_t *l_indptr, const int32_t *l_indices, const double *l_vals, int64_t n,
                         int64_t l_nnz, double lmax, const double *coeffs, int32_t n_scales, int32_t order,
                         double tol, void **out_handle, void *stream);
int srg_cheby_sparse_info(void *handle, int32_t scale, int64_t *out_nnz, int64_t *out_products,
                          int64_t *out_pattern_nnz);
int srg_cheby_sparse_fetch(void *handle, int32_t scale, int32_t *indptr, int32_t *indices, float *vals,
                           void *stream);
int srg_cheby_sparse_free(void *handle);

/* Largest eigenvalue of a symmetric CSR matrix (the Laplacian) by Lanczos with full re-orthogonalisation, fp64, on
 * the device: what pygsp's Graph.estimate_lmax obtains from ARPACK (eigsh k=1, tol 5e-3; wavelet/src/utils.py:83,
 * SSRG/models/base_scalable/base_model.py:184) before multiplying by 1.01.  Stops when the top Ritz value moved by
 * less than tol / 10 over two steps, or after max_steps (<= 256).  Set-up path: synchronises the stream. */
int srg_lanczos_lambda_max_f64(const int32_t *indptr, const int32_t *indices, const double *vals, int64_t n,
                               double tol, int32_t max_steps, double *out_lambda, int32_t *out_steps, void *stream);

/* ---- (f-4) dataset augmentation before the path: edge_augument (SSRG/data_augument.py:73-103) ----------------
 * srg_endpoint_counts_i64   Counter(cat(edge_row, edge_col)) (:75-77): counts[v] and the first position of v in the
 *                           concatenation (INT64_MAX when v never occurs) — the Counter's insertion order, which
 *                           breaks the ties of `sorted(counts.items(), key=degree)` (:81)
 * srg_candidate_topk_f32    for low-degree node i (nodes[i]): L2 distances of the soft labels to cand[i, 0..cand_cnt[i])
 *                           (compute_distance, SSRG/utils.py:35-38) and the k_sel[i] closest in ascending order,
 *                           ties by candidate position, as pairs (node, candidate) at out_off[i]  (:88-95)
 * srg_csr_to_edge_index_i64 2 x nnz int64 edge list of a CSR pattern in (row, col) order = torch.unique(dim=1) of the
 *                           symmetrised list (:97-102) when the CSR comes from srg_edges_to_sym_csr */
int srg_endpoint_counts_i64(const int64_t *row, const int64_t *col, int64_t m, int64_t n, int32_t *counts,
                            int64_t *first_pos, int32_t *flags, void *stream);
int srg_candidate_topk_f32(const float *soft, int64_t ld, int64_t n, int32_t n_classes, const int32_t *nodes,
                           const int32_t *cand, const int32_t *cand_cnt, const int32_t *k_sel, const int32_t *out_off,
                           int32_t c_max, int32_t n_low, int64_t *out_src, int64_t *out_dst, void *stream);
int srg_csr_to_edge_index_i64(const int32_t *indptr, const int32_t *indices, int64_t n, int64_t nnz,
                              int64_t *out_edge_index, void *stream);

/* ---- (e) row-partitioned multi-GPU hop ------------------------------------------------------- */
/*
 * One hop over this rank's row slice with the exchange fused into the epilogue: every finished row
 * (global row dest_row0 + i) is stored into ALL n_dests (<= 8) next-hop buffers — this rank's own
 * and its peers', mapped over NVLink through srg_ipc_open — so no separate all-gather runs.  `dests`
 * is a HOST array of device pointers to the full (all rows) n_total x ldy buffers.  The caller
 * orders hops across ranks (a stream-ordered collective / barrier between hops).
 */
int srg_spmm_csr_f32_push(const int32_t *indptr, const int32_t *indices, const float *vals,
                          int64_t n_rows, int64_t nnz, const float *X, int64_t ldx,
                          float *const *dests, int32_t n_dests, int64_t dest_row0, int64_t ldy,
                          int32_t F, void *stream);
/* The same with one row offset per destination (dest_row0s, HOST array): the rank's own n_rows x ldy copy of the
 * hop — element k of the K+1 matrices GraphOp.propagate returns (SSRG/operators/base_operator.py:31-36) — is one
 * more destination with offset 0, so keeping every hop costs no extra pass.  n_dests <= 9. */
int srg_spmm_csr_f32_push2(const int32_t *indptr, const int32_t *indices, const float *vals,
                           int64_t n_rows, int64_t nnz, const float *X, int64_t ldx,
                           float *const *dests, const int64_t *dest_row0s, int32_t n_dests, int64_t ldy,
                           int32_t F, void *stream);
/* Orders consecutive push hops across ranks without a collective: publishes `epoch` into slot `my_slot` of every
 * peer's flag array (peer_flags: HOST array of n_peers device pointers to >= 32 uint32 each, peer-mapped) and waits
 * until every peer's epoch has arrived in local_flags[0..n_peers).  Stream-ordered; a peer that does not show up
 * within 2 s sets *timeout_flag (device int32, may be NULL) instead of hanging the GPU. */
int srg_peer_barrier(void *local_flags, void *const *peer_flags, int32_t n_peers, int32_t my_slot,
                     uint32_t epoch, int32_t *timeout_flag, void *stream);
/* input exchange without a collective: copy n_rows x ld floats into rows dest_row0.. of every
 * destination buffer (own + peers) */
int srg_push_rows_f32(const float *src, int64_t n_rows, int64_t ld, float *const *dests,
                      int32_t n_dests, int64_t dest_row0, void *stream);
/* ---- e (7): native multi-GPU handle — row partition + one NCCL all-gather per hop -------
 * The reference has no multi-GPU code.  One process per GPU; rank p owns rows [p * rows_per, (p + 1) * rows_per)
 * with rows_per = ceil(n / world); the degree vector is all-gathered once for the normalisation
 * (SSRG/operators/utils.py:81-93 applied to the local rows), every hop is an all-gather of the previous hop's
 * slices followed by the local SpMM (SSRG/operators/base_operator.py:31-36).  Bitwise equal to 1 GPU.
 * NCCL is bound at run time (dlopen libnccl.so.2); SRG_ERR_UNSUPPORTED when it cannot be loaded.
 *   srg_dist_unique_id   rank 0: ncclGetUniqueId into 128 bytes, to be handed to every rank by the caller
 *   srg_dist_init        collective: ncclCommInitRank + the two full n_pad x ld feature buffers (ld = roundup(F, 8))
 *   srg_dist_init_comm   the same around an existing ncclComm_t (not destroyed by srg_dist_destroy)
 *   srg_dist_partition   this rank's row0 / n_local / rows_per / ld
 *   srg_dist_propagate   indptr / indices / data: the rank's rows of the RAW adjacency (device, global column
 *                        ids, symmetric overall — the caller's promise, the mirror rows live elsewhere);
 *                        x_local: device n_local x ld_x; out_hops: NULL or a HOST array of K + 1 device pointers
 *                        (n_local x ld_out each, NULL entries skipped; [0] receives the input);
 *                        flags: device int32 (caller zeroes it), SRG_FLAG_*.  Stream-ordered, no host sync. */
int srg_dist_unique_id(void *id128);
int srg_dist_init(const void *id128, int32_t world, int32_t rank, int64_t n, int32_t F, void **out_handle);
int srg_dist_init_comm(void *nccl_comm, int32_t world, int32_t rank, int64_t n, int32_t F, void **out_handle);
int srg_dist_partition(const void *handle, int64_t *row0, int64_t *n_local, int64_t *rows_per, int64_t *ld);
int srg_dist_propagate(void *handle, const int32_t *indptr, const int32_t *indices, const void *data,
                       int val_dtype, int64_t nnz, const float *x_local, int64_t ld_x, int32_t K, double r,
                       double ppr_alpha, float *const *out_hops, int64_t ld_out, int32_t *flags, void *stream);
int srg_dist_destroy(void *handle);

/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `stream`: copy-engine transfer between local and
 * peer-mapped buffers (the per-hop exchange of the "copy" multi-GPU mode). */
int srg_copy_async(void *dst, const void *src, int64_t bytes, void *stream);
/* peer-mappable device buffers: plain cudaMalloc + CUDA IPC handles (64 bytes) */
int srg_ipc_alloc(void **ptr, int64_t bytes);
int srg_ipc_free(void *ptr);
int srg_ipc_get_handle(void *ptr, void *handle64);
int srg_ipc_open(const void *handle64, void **ptr);
int srg_ipc_close(void *ptr);

/* ---- a1: K hops, device resident  (SSRG/operators/base_operator.py:31-35) ------------------ */
/* hops[0] = input features, hops[k] = A^ * hops[k-1], k = 1..K; all n x ld fp32 device buffers. */
int srg_propagate_khop_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                           int64_t n, int64_t nnz, float *const *hops, int64_t ld, int32_t F,
                           int32_t K, void *stream);

/* ---- a9: Chebyshev heat-wavelet filter  (wavelet/src/utils.py:89-104,125-138; pygsp cheby_op) --- */
/*
 * Combinatorial Laplacian L = diag(W 1) - W of a canonical CSR W (what pygsp.graphs.Graph(W).L is
 * for lap_type='combinatorial').  out_indptr n+1, out_indices / out_vals capacity nnz + n; exact
 * zeros dropped; degree (nullable) = W.sum(1) in numpy's summation order.
 */
int srg_laplacian_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                      int val_dtype, int64_t n, int32_t *out_indptr, int32_t *out_indices,
                      double *out_vals, double *out_degree, int32_t *out_flags, void *stream);

/*
 * r_s = sum_k c[s][k] T_k(L~) X for n_scales (<= 4) coefficient vectors sharing the T_k, fp64:
 *   T0 = X, T1 = (L X - a X)/a, T_k = (2/a)(L T_{k-1} - a T_{k-1}) - T_{k-2},  a = lmax/2,
 *   r_s = 0.5 c[s][0] T0 + sum_{k>=1} c[s][k] T_k      (coeffs: HOST array [n_scales][order+1])
 * The recurrence and the accumulation are the epilogue of the SpMM of each order (one launch per
 * order).  tol: values < tol are zeroed in the last step (wavelet/src/utils.py:98); pass NaN for no
 * threshold.  out_r[s]: device n x ld fp64; out_r32 (nullable array, nullable entries): float32
 * copy of the thresholded result (leading dimension ld32).  work0/work1: n x ld fp64 scratch.
 * ld (in doubles) must be even, pointers 16-byte aligned.
 */
int srg_cheby_filter_f64(const int32_t *lap_indptr, const int32_t *lap_indices,
                         const double *lap_vals, int64_t n, const double *X, int64_t ld, int32_t B,
                         double lmax, const double *coeffs, int32_t n_scales, int32_t order,
                         double tol, double *const *out_r, float *const *out_r32, int64_t ld32,
                         double *work0, double *work1, void *stream);

/* ---- 8f-3: wavelet post-processing on the device ------------------------------------------------ */
/* thresholded dense float32 block (n x B, the out_r32 of srg_cheby_filter_f64) -> block CSR with global
 * column ids col0 + j (wavelet/src/utils.py:99-103, base_model.py:247-251).  Call it first with
 * out_cols == NULL (count only: out_indptr[n] is the entry count) to size out_cols / out_vals, then
 * again to fill them.  row_totals (nullable, n ints) is incremented by the per-row counts. */
int srg_dense_block_to_csr_f32(const float *dense, int64_t ld, int64_t n, int32_t B, int32_t col0,
                               int32_t *out_indptr, int32_t *out_cols, float *out_vals,
                               int64_t capacity, int32_t *row_totals, void *stream);
/* append the rows of one block CSR to the merged CSR at cursor[row] and advance the cursor; calling it
 * block after block in column order is sparse.hstack (base_model.py:265) without a sort */
int srg_csr_block_scatter_f32(int64_t n, const int32_t *block_indptr, const int32_t *block_cols,
                              const float *block_vals, int32_t *cursor, int32_t *out_cols,
                              float *out_vals, void *stream);
/* sklearn.preprocessing.normalize(X, norm='l1', axis=1) on a float32 CSR, in place
 * (wavelet/src/utils.py:106-112): sequential double sum of |v| per row, v = float32(v / sum) when sum != 0 */
int srg_csr_row_normalize_l1_f32(int64_t n, const int32_t *indptr, float *vals, void *stream);
/* int32 exclusive prefix sum, out has n + 1 entries */
int srg_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out, void *stream);

/* ---- 8f-1: message-operator aggregation of the hop list, on the device ------------------------- */
/* the non-learnable aggregators of SSRG/operators/message_operator/: last_message_op.py:9,
 * sum_message_op.py:9, mean_message_op.py:9, max_message_op.py:11, min_message_op.py:11,
 * concat_message_op.py:9, simple_weighted_message_op.py:44-58 (+ utils.py:426-437) */
#define SRG_AGG_NONE 0
#define SRG_AGG_LAST 1
#define SRG_AGG_SUM 2
#define SRG_AGG_MEAN 3
#define SRG_AGG_MAX 4
#define SRG_AGG_MIN 5
#define SRG_AGG_CONCAT 6
#define SRG_AGG_WEIGHTED 7
#define SRG_AGG_NAFS 8 /* over_smooth_distance_op.py:11-33; srg_propagate_aggregate_host only */
/* fold one hop matrix x (n x ld_x) into acc[:, col0:col0+F] (n x ld_acc).  first != 0 initialises acc
 * from x; mode -1 finalises a mean (acc / weight).  The arithmetic is the reference's torch
 * expression evaluated in hop order (sequential fp32 adds; weighted: product first). */
int srg_aggregate_update_f32(float *acc, int64_t ld_acc, int32_t col0, const float *x, int64_t ld_x,
                             int64_t n, int32_t F, int32_t mode, float weight, int32_t first,
                             void *stream);
/*
 * NAFS over-smoothing-distance aggregation (OverSmoothDistanceWeightedOp.combine,
 * SSRG/operators/message_operator/over_smooth_distance_op.py:11-33; aggregator of SSRG/models/nafs.py:12):
 *   score[i][j] = ((x0[i] . xj[i]) / (||xj[i]|| + 1e-10)) / (||x0[i]|| + 1e-10),
 *   weight[i]   = softmax_j(score[i]),   out[i] = sum_j weight[i][j] * xj[i]   (fp32, hop order).
 * hops: HOST array of n_hops (<= 64) DEVICE pointers to n x ld matrices (hops[0] = the input
 * features); out: device n x ld_out; weights_out: optional device n x n_hops (the softmax weights).
 * The reference walks the rows in a Python loop; one warp per row here.
 */
int srg_nafs_combine_f32(const float *const *hops, int32_t n_hops, int64_t ld, int64_t n, int32_t F,
                         float *out, int64_t ld_out, float *weights_out, void *stream);
/*
 * srg_propagate_host + the aggregation, with ONLY the aggregate coming back:
 *   out_agg (host, n x F_out, F_out = F or (agg_end-agg_start)*F for CONCAT) =
 *       msg_op.aggregate(graph_op.propagate(adj, feature))  over the hop slice [agg_start, agg_end)
 *   agg_weights: host array of agg_end-agg_start floats (WEIGHTED only).
 * Two ping-pong hop buffers + the accumulator stay on the device; K-1 of the K device->host copies
 * of srg_propagate_host disappear.  SRG_AGG_NAFS keeps all K+1 hop buffers (its weights need every
 * hop) and takes the whole list (agg_start / agg_end are ignored).
 */
int srg_propagate_aggregate_host(const int32_t *indptr, const int32_t *indices, const void *data,
                                 int val_dtype, int64_t n, int64_t nnz, const float *features,
                                 int32_t F, const int32_t *feature_mask, int32_t K, double r,
                                 double ppr_alpha, int32_t agg_mode, int32_t agg_start,
                                 int32_t agg_end, const float *agg_weights, float *out_agg,
                                 int device);

/* Host utility: 1 when every stored value of a host array is exactly 1 (or val_dtype is SRG_VAL_ONES), else 0;
 * `threads` worker threads scan it.  scipy holds an unweighted adjacency as float64 ones (`A.data`, what
 * SSRG/operators/utils.py:82 receives); with SRG_ONES_SHORTCUT=1 in the environment srg_propagate_host starts
 * without uploading that array while this check runs in the shadow of the transfers, and falls back to the
 * regular path when the check fails.  Off by default. */
int srg_host_all_ones(const void *data, int val_dtype, int64_t nnz, int32_t threads);

/* layout helpers: host layout (ld == F) <-> padded device layout (ld % 8 == 0, pad = 0).
 * mask (optional, int32 n x F, SSRG/data_process.py:38-39) is applied as x * mask
 * (SSRG/data_augument.py:28) while repacking. */
int srg_pack_features_f32(const float *src, int64_t ld_src, float *dst, int64_t ld_dst,
                          int64_t n, int32_t F, const int32_t *mask, void *stream);
int srg_unpack_features_f32(const float *src, int64_t ld_src, float *dst, int64_t ld_dst,
                            int64_t n, int32_t F, void *stream);

/* ---- a1..a7 in one call, host buffers  (GraphOp.propagate) ---------------------------------- */
/*
 * Host CSR of the RAW adjacency + host features -> K host output matrices (hop 1..K), each
 * n x F fp32 C-contiguous.  Does normalisation (r, ppr_alpha as above) and the hops on
 * `device`; out_norm_* (all nullable) receive the normalised CSR (scipy layout: int32 indptr
 * n+1, int32 indices, fp64 data; capacity nnz + n entries); *out_nnz receives its nnz.
 * feature_mask: optional host int32 n x F.
 */
int srg_propagate_host(const int32_t *indptr, const int32_t *indices, const void *data,
                       int val_dtype, int64_t n, int64_t nnz, const float *features, int32_t F,
                       const int32_t *feature_mask, int32_t K, double r, double ppr_alpha,
                       float *const *out_hops, int32_t *out_norm_indptr,
                       int32_t *out_norm_indices, double *out_norm_data, int64_t *out_nnz,
                       int device);

/* construct_adj alone on host buffers (same outputs as above; capacity nnz + n). */
int srg_construct_adj_host(const int32_t *indptr, const int32_t *indices, const void *data,
                           int val_dtype, int64_t n, int64_t nnz, double r, double ppr_alpha,
                           int32_t *out_indptr, int32_t *out_indices, double *out_data,
                           int64_t *out_nnz, int device);

/* release cached device workspaces / pinned staging owned by the host entry points */
int srg_release_workspace(void);

/* ---- literal ABI shims: same symbol names and signatures as the reference's libraries, so
 * SSRG/operators/utils.py can load this library in place of ./csrc/libmatmul.so or
 * ./csrc/libcudamatmul.so unmodified.  Host pointers; answer is accumulated into
 * (answer += A*mat) exactly as matmul.c does. ------------------------------------------------ */
void FloatCSRMulDenseOMP(float answer[], float data[], int indices[], int indptr[], float mat[],
                         int mat_row, int mat_col);
int FloatCSRMulDense(float answer[], int data_nnz, float data[], int indices[], int indptr[],
                     float mat[], int mat_row, int mat_col);

#ifdef __cplusplus
}
#endif
#endif /* SRGNN_B200_H */
