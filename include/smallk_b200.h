/* smallk_b200 — C ABI of the B200-native NMF iteration hot path.
 *
 * This is the drop-in boundary. The reference has no FFI of its
 * own for this path; its seams are the C-style library functions of
 * common/include/nmf.hpp:77-92 (Nmf, NmfSparse) and the solver-functor concept
 * NmfSolve is templated on (common/include/nmf_solve_generic.hpp:30-40). Each
 * entry point below names the reference interface it stands in for. Host code
 * (the headers under smallk_b200/host/, the C++ mirror of nmf.hpp / smallk.hpp, and the
 * Python ctypes binding used by tests and bench.py) sits ABOVE this header.
 *
 * Conventions: plain pointers and sizes, no exceptions across the boundary,
 * every function returns an smk_result (the reference's `Result` values,
 * common/include/nmf.hpp:17-26, plus SMK_CUDA_ERROR) and smk_last_error() gives
 * the text. All matrices are column-major doubles. One host thread per context.
 * There is no CPU fallback: without a CUDA device smk_create() fails.
 */
#ifndef SMALLK_B200_H
#define SMALLK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smk_ctx smk_ctx;

/* common/include/nmf.hpp:17-26 */
typedef enum
{
    SMK_OK = 0,
    SMK_NOTINITIALIZED = -1,
    SMK_INITIALIZED = -2,
    SMK_BAD_PARAM = -3,
    SMK_FAILURE = -4,
    SMK_SIZE_TOO_LARGE = -5,
    SMK_FLATCLUST_FAILURE = -6,
    SMK_CUDA_ERROR = -100
} smk_result;

/* common/include/nmf.hpp:28-41 (same numeric values) */
typedef enum { SMK_MU = 0, SMK_HALS = 1, SMK_RANK2 = 2, SMK_BPP = 3 } smk_algorithm;
typedef enum { SMK_PG_RATIO = 0, SMK_DELTA_FNORM = 1 } smk_progress;

/* NmfOptions, common/include/nmf.hpp:56-70 (max_threads has no meaning on the GPU and is ignored) */
typedef struct
{
    double tol;
    int algorithm;        /* smk_algorithm */
    int prog_est_algorithm; /* smk_progress */
    int height, width, k;
    int min_iter, max_iter, tolcount;
    int max_threads;
    int verbose;
    int normalize;
} smk_nmf_options;

/* NmfStats, common/include/nmf.hpp:43-53 */
typedef struct
{
    unsigned long long elapsed_us;
    int iteration_count;
} smk_nmf_stats;

/* ---- lifecycle: NmfInitialize / NmfIsInitialized / NmfFinalize (common/src/nmf.cpp:36-52) ----
 * A context is used by one host thread at a time; different contexts (own stream, own scratch) may be used from different host
 * threads concurrently (the hierclust driver sorts on two worker contexts while the main one factors). ---- */
int smk_create(smk_ctx** ctx, int device);
void smk_destroy(smk_ctx* ctx);
const char* smk_last_error(const smk_ctx* ctx);
int smk_device_sm_count(const smk_ctx* ctx);
int smk_device_index(const smk_ctx* ctx);     /* the CUDA device smk_create was given (-1 for a null context) */
/* Run on a caller-owned CUDA stream (e.g. torch's current stream); 0 restores the context's own. */
int smk_set_stream(smk_ctx* ctx, void* cuda_stream);
int smk_synchronize(smk_ctx* ctx);

/* ---- multi-GPU: one process per GPU; A and H are sharded by column block.
 * The 128-byte id comes from smk_comm_unique_id() on rank 0 and is broadcast by the host
 * program (torch.distributed / MPI). After smk_comm_init the solver sums H*H' (k x k) and H*A' (k x m) over the
 * ranks each outer iteration and redistributes the row blocks of W; everything else stays local. The exchanges are
 * the library's own kernels over NVLink peer memory (csrc/peer.cu; all ranks on one NVSwitch node, <= 8); NCCL
 * bootstraps them and is the selectable alternative (environment SMK_PEER=0). ---- */
int smk_comm_unique_id(void* id128);
int smk_comm_init(smk_ctx* ctx, int rank, int nranks, const void* id128);

/* ---- the input matrix A (this rank's column block when sharded) ----
 * DenseMatrix<R> A(m, n, buf_a, ldim_a): common/src/nmf.cpp:224. */
int smk_load_dense(smk_ctx* ctx, const double* A_host, long long ldA, int m, int n);
/* Same, but A already lives in device memory and is borrowed, not copied. */
int smk_load_dense_device(smk_ctx* ctx, const double* A_dev, long long ldA, int m, int n);
/* SparseMatrix<double> A(height, width, nz, col_offsets, row_indices, data): common/src/nmf.cpp:288.
 * CSC, 32-bit indices, rows need not be sorted, duplicates are kept. */
int smk_load_csc(smk_ctx* ctx, int m, int n, unsigned int nnz,
                 const unsigned int* col_offsets, const unsigned int* row_indices, const double* data);

/* ---- column subsets of the loaded matrix (hierclust) ----
 * A.SubMatrixColsCompact(Asubset, col_indices, old_to_new_rows, new_to_old_rows):
 * common/include/sparse_matrix_impl.hpp:479-591 (sparse: the listed columns in list order; rows left without
 * entries are dropped and the rest renumbered in ascending order) and dense_matrix_impl.hpp:224-285 (dense: all
 * rows kept). The subset becomes the ACTIVE matrix of the context: smk_nmf / smk_solver_* then factor it, with
 * opts.height = *new_height and opts.width = count. new_to_old_rows is a host buffer of at least m entries;
 * *new_height of them are written. The loaded matrix itself stays resident; smk_select_all re-activates it.
 * Errors as the reference's std::logic_error cases: empty list, index out of range, all-zero submatrix -> SMK_BAD_PARAM. */
int smk_select_columns(smk_ctx* ctx, const unsigned int* col_indices, int count, int* new_height,
                       unsigned int* new_to_old_rows);
int smk_select_all(smk_ctx* ctx);

/* ---- Result Nmf(opts, A, W, H, stats) / NmfSparse(...): common/src/nmf.cpp:173,232 ----
 * W (m x k) and H (k x n) are host buffers: initial guess in, factors out. The matrix must have
 * been loaded with one of the calls above; opts.height/width must match it. */
int smk_nmf(smk_ctx* ctx, const smk_nmf_options* opts,
            double* W_host, int ldW, double* H_host, int ldH, smk_nmf_stats* stats);

/* ---- the solver-functor seam (Init / operator() / progress estimator,
 * common/include/nmf_solver_bpp.hpp:310,342; progress_estimator_generic.hpp:114-148) ----
 * begin : upload W0, H0 and run Solver::Init + ProgressEst::Init.
 * step  : run `count` outer iterations (solver() only, no metric). Returns SMK_FAILURE where the
 *         reference's solver() returns false.
 * progress: ProgressEst::Update(iter, ...) for the current state -> *metric.
 * get   : copy the current W (m x k), H (k x n) and optionally gradW, gradH to host buffers. */
int smk_solver_begin(smk_ctx* ctx, const smk_nmf_options* opts,
                     const double* W0_host, int ldW, const double* H0_host, int ldH);
int smk_solver_step(smk_ctx* ctx, int count);
int smk_solver_progress(smk_ctx* ctx, double* metric);
/* run   : `count` times { operator(); ProgressEst::Update } — the body of NmfSolve's loop (nmf_solve_generic.hpp:67-123) with
 *         the metric evaluated every iteration — enqueued back to back with NO host synchronisation in between: the metric
 *         of every iteration is computed on the device and all of them come back in one copy at the end (metrics may be
 *         NULL). The first progress evaluation after begin stores pg0, exactly as ProgEstGenericPgRatio::Update does. */
int smk_solver_run(smk_ctx* ctx, int count, double* metrics);
int smk_solver_get(smk_ctx* ctx, double* W_host, int ldW, double* H_host, int ldH,
                   double* gradW_host, int ldgW, double* gradH_host, int ldgH);
int smk_solver_normalize(smk_ctx* ctx);     /* NormalizeAndScale(W, H): common/include/normalize.hpp:118-138 */
/* Device time of the last smk_solver_step call (CUDA events on the context's stream), and the
 * number of kernels it launched. */
int smk_solver_last_step_ms(smk_ctx* ctx, float* ms, long long* kernel_launches);

/* Measurement hook (environment SMK_PHASES=1): device time per phase of the solver steps since the last call, as
 * "name=ms;name=ms;..." (CUDA events between the phases; includes the exchange kernels of the multi-GPU path). */
int smk_phase_report(smk_ctx* ctx, char* buf, int len);

/* Measurement hook for the roofline line of bench.py: re-runs one of the two big contractions of the
 * current solver state `reps` times (which: 0 = W'A, 1 = H A') and reports the mean device time per
 * launch from CUDA events on the context's stream. State is left as a solver step would leave it. */
int smk_solver_time_product(smk_ctx* ctx, int which, int reps, float* mean_ms);

/* bool NnlsHals(A, W, H, tol, verbose, max_iter): common/include/nnls.hpp:249-316 — the flat-clustering step of
 * HierNmf2WithFlat (hierclust/include/clust_flat_generic.hpp:33-74). W (m x k) is fixed, H (k x n) carries the
 * initial guess in and the solution out; on success (pg < tol * pg0) W's columns are normalised and H's rows
 * scaled (normalize.hpp:118-138) and both are copied back. SMK_FAILURE = iteration limit reached. */
int smk_nnls_hals(smk_ctx* ctx, int k, double* W_host, int ldW, double* H_host, int ldH, double tol, int max_iter,
                  int* iterations);

/* ---- bool preprocess_tf(TermFrequencyMatrix& A, term_indices, doc_indices, scores, MAX_ITER, DOCS_PER_TERM, TERMS_PER_DOC):
 * preprocessor/src/preprocess.cpp:81-250 — the tf-idf pipeline that produces the sparse matrices NmfSparse factors: rows sorted
 * inside every column, rounds of term pruning (total count < docs_per_term, or present in every document), document pruning
 * (< terms_per_doc distinct terms) and removal of duplicated documents (the copy with the largest index stays) until a round
 * removes nothing, then scores (1 + ln count) * ln(width / document frequency) scaled to unit column norm. In: m x n CSC of
 * term counts (host arrays, counts as doubles as the MatrixMarket loader delivers them). Out (host arrays with capacities of the
 * INPUT sizes: n + 1, nnz, nnz, nnz, m, n): the pruned matrix, its scores aligned with its entries, and for every surviving
 * row / column its original index. Index outputs are bit-identical to the reference's; scores agree to rounding level.
 * SMK_FAILURE = every document was pruned (the reference returns false). ---- */
int smk_preprocess_tf(smk_ctx* ctx, unsigned int m, unsigned int n, unsigned int nnz, const unsigned int* col_offsets,
                      const unsigned int* row_indices, const double* counts, unsigned int max_iter, unsigned int docs_per_term,
                      unsigned int terms_per_doc, unsigned int* out_m, unsigned int* out_n, unsigned int* out_nnz,
                      unsigned int* out_col_offsets, unsigned int* out_row_indices, unsigned int* out_counts, double* out_scores,
                      unsigned int* term_indices, unsigned int* doc_indices);

/* ---- ordering primitives for the host-side tree code of hierclust ----
 * desc_ordered(values) of hierclust/include/clust_hier_util.hpp:46-57: the permutation that lists the indices by
 * decreasing value, ties by increasing index (a stable descending radix sort on the device; -0.0 is treated as 0.0,
 * NaN is not supported). order_host receives n ints. */
int smk_argsort_desc(smk_ctx* ctx, const double* values_host, int n, int* order_host);
/* std::sort(v.begin(), v.end(), std::greater<double>()) on a host array (NDCG ideal scores, clust_hier_util.hpp:86). */
int smk_sort_desc(smk_ctx* ctx, double* values_host, int n);

/* ---- primitive-level entry points (the linear-algebra seam, SURVEY.md §8b ④), host buffers ----
 * Gemm(orientA, orientB, 1, A, B, 0, C) on DenseMatrix: common/include/dense_matrix_ops.hpp:255-270.
 * C (M x N) = op(A) * op(B); transA/transB are 0 (NORMAL) or 1 (TRANSPOSE). */
int smk_gemm(smk_ctx* ctx, int transA, int transB, int M, int N, int K,
             const double* A, int ldA, const double* B, int ldB, double* C, int ldC);
/* bool NnlsBlockpivot(LHS, RHS, X, Y): common/include/nnls.hpp:144. LHS k x k, RHS/X/Y k x q. */
int smk_nnls_bpp(smk_ctx* ctx, int k, int q, const double* LHS, const double* RHS, double* X, double* Y);
/* Diagnostic for the parity tests: how often UpdatePassiveSet's backup rule (common/src/nnls.cpp:64-72: toggle the row
 * BitMatrix::MaxRowIndex reports, defect for k > 32 included) has fired since the last smk_nnls_bpp / smk_solver_begin. */
int smk_nnls_backup_count(smk_ctx* ctx, int* count);
/* Gemm on SparseMatrix (common/include/sparse_gemm.hpp:26-74) against the loaded CSC matrix:
 * variant 0: C = alpha*A*B + beta*C, 1: alpha*A*B' + beta*C, 2: alpha*B*A + beta*C, 3: alpha*B'*A + beta*C.
 * B is Bh x Bw, C is Ch x Cw, tight leading dimensions. */
int smk_sparse_gemm(smk_ctx* ctx, int variant, double alpha, const double* B, int Bh, int Bw,
                    double beta, double* C, int Ch, int Cw);

#ifdef __cplusplus
}
#endif
#endif /* SMALLK_B200_H */
