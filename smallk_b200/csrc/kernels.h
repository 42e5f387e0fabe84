// smallk_b200 — host-side launchers of the sm_100a kernels (internal header).
#pragma once

#include <cuda_runtime.h>
#include <cstddef>
#include <string>
#include <algorithm>

#include "peer.h"

namespace smk {

// device status words shared by the kernels of one solver
enum { ST_ANY_NONOPT = 0, ST_FAIL_ITER = 1, ST_NORM_EPS = 2, ST_DEFER_COUNT = 3, ST_BAD_INDEX = 4, ST_PG_NAN = 5, ST_COMM_TIMEOUT = 6, ST_BACKUP_COUNT = 7, ST_COUNT = 8 };

// ---- gemm_f64.cu ----------------------------------------------------------
// Optional epilogue of gemm_f64 for the column-sharded H*A' (peer.cu): column c of C belongs to rank c / cols_per_rank and is
// stored into that rank's receive slot [rank] instead of C; the CTA finishing the last tile publishes `epoch` to all peers.
struct GemmScatter
{
    PeerTable table;
    size_t recv_off = 0;            // byte offset of the receive slots in every rank's exchange region
    long long piece = 0;            // doubles per (owner, sender) slot = M * cols_per_rank
    int cols_per_rank = 1;
    int rank = 0, nranks = 0;       // nranks == 0: epilogue off
    int ntiles = 0;                 // filled by gemm_f64
    unsigned long long epoch = 0;
    unsigned int* done = nullptr;   // device counter, zero on entry, left zero
};
// C (M x N) = A (M x R, col-major) * Bop - D, Bop = B (R x N col-major) if !nt, else B' with B (N x R col-major).
void gemm_f64(cudaStream_t stream, bool nt, int M, int N, int R,
              const double* A, long long lda, const double* B, long long ldb,
              double* C, long long ldc, const double* D, long long ldd,
              double* workspace, size_t workspace_bytes, int num_sms, int* partials_only = nullptr,
              const GemmScatter* scatter = nullptr);
// partials_only != null: the split-R partial tiles are left in the workspace ([*partials_only][M x N], ld = M, at least one)
// and NOT summed — the multi-GPU reduce-scatter adds them in split order while it forwards them (peer.cu).
int gemm_pick_splits(int M, int N, int R, int num_sms, size_t workspace_bytes, bool whole_tiles = false);
// Zeroes the arrival counters at the head of a split-R workspace: once after allocating it (launches leave them zero).
void gemm_workspace_prepare(cudaStream_t stream, double* workspace, size_t workspace_bytes);
// where gemm_f64(..., partials_only) leaves its partial tiles inside the workspace
const double* gemm_partials(const double* workspace);

// ---- nnls_bpp.cu ----------------------------------------------------------
void nnls_bpp(cudaStream_t stream, int k, int q, const double* LHS, long long ldl,
              const double* RHS, long long ldr, double* X, long long ldx, double* Y, long long ldy,
              int* status, unsigned int* counter, void* deferred, int outer_iter, int num_sms,
              const double* Ginv = nullptr, const int* ginv_flag = nullptr);
// LHS^-1 (k x k, tight) + success flag for nnls_bpp's complement path (k > 32), one launch; when the caller does not pass them
// nnls_bpp forms them itself on `stream`. The solvers run it on a side stream under the big product preceding the solve.
void nnls_prepare_inverse(cudaStream_t stream, int k, const double* LHS, long long ldl, double* Ginv, int* ok);
size_t nnls_deferred_bytes(int q, int k, int num_sms);
// does nnls_bpp use G^-1 at this k? (32 < k <= 256: the complement path of the register / shared-memory kernels)
inline bool nnls_uses_inverse(int k) { return k > 32 && k <= 256; }
void nnls_bpp_finish(cudaStream_t stream, int k, int q, double* X, long long ldx, double* Y, long long ldy, int* status, int num_sms);

// ---- elementwise.cu -------------------------------------------------------
// out (cols x rows, ld = ldo) = in' where in is rows x cols (ld = ldi)
void transpose_f64(cudaStream_t stream, int rows, int cols, const double* in, long long ldi, double* out, long long ldo);
// X .*= Num ./ (Den + 1e-13)      (nmf_solver_mu.hpp:27-71)
void mu_update(cudaStream_t stream, long long count, double* X, const double* Num, const double* Den);
// acc[slot] = sum over entries of G^2 where (G < 0 || X > 0)   (projected_gradient.hpp:125-171)
// partial: device scratch of at least 1024 doubles. Deterministic two-level reduction.
void pg_sumsq(cudaStream_t stream, long long count, const double* G, const double* X, double* partial, double* acc_slot, int num_sms);
// Both sums of ProjectedGradientNorm in one launch: acc[0] from (G1, X1), acc[1] from (G2, X2) (either count may be 0).
// partial: >= 2048 doubles; ticket: one self-resetting arrival counter. With prog != null the last block also runs
// ProgressEst::Update on the device: prog[0] = pg0, prog[1] = "pg0 captured"; the metric goes to *metric_out.
void pg_pair(cudaStream_t stream, long long c1, const double* G1, const double* X1, long long c2, const double* G2, const double* X2,
             double* partial, unsigned int* ticket, double* acc, double* prog, double* metric_out, int* status, int num_sms);
// ProgressEst::Update from sums already in acc[0..1]: mode 0 = PG ratio (progress_estimator_generic.hpp:87-104),
// mode 1 = relative delta-W (:58-69: acc[0] = ||W_prev - W||^2, acc[1] = ||W||^2).
void progress_metric_launch(cudaStream_t stream, int mode, const double* acc, double* prog, double* metric_out, int* status);
// acc[slot] = sum (A - B)^2 (B may be null -> sum A^2)
void diff_sumsq(cudaStream_t stream, long long count, const double* A, const double* B, double* partial, double* acc_slot, int num_sms);

// ---- factors.cu -----------------------------------------------------------
// One HALS sweep over the rows of X (k x q): for r = 0..k-1, for every column j
//   x(r,j) <- max(0, x(r,j) + (R(r,j) - sum_p G(r,p) x(p,j)) / G(r,r)),  NaN -> 0
// (nmf_solver_hals.hpp:26-61 for H with G = W'W, R = W'A; :64-117 for W' with G = HH', R = HA').
// normalize_rows = true adds the in-sweep unit-2-norm scaling of row r (= column r of W) and the
// all-zero -> epsilon rule of :103-115. partial: >= 4096 doubles of scratch.
// scratch (optional, hals_sweep_scratch_doubles(q) doubles) enables the blocked W-side sweep, which reads X k/16
// times per sweep instead of k times.
void hals_sweep(cudaStream_t stream, int k, int q, double* X, const double* G, const double* R,
                bool normalize_rows, double* norms, double* partial, int num_sms, double* scratch = nullptr);
size_t hals_sweep_scratch_doubles(int q);
// Rank-2 solve + optimal active set for X (2 x q): nmf_solver_rank2.hpp:25-318.
// w_side selects the SystemSolveW / OptimalActiveSetW formulas (same algebra, transposed roles).
void rank2_update(cudaStream_t stream, int q, double* X, const double* G, const double* B, bool w_side,
                  int* status, int outer_iter);
// NormalizeAndScale (normalize.hpp:118-161): rows of Wt scaled to unit norm, rows of H by the norm.
// If HHt/HAt are given (rank-2, nmf_solver_rank2.hpp:422-441) they are rescaled as well.
void normalize_and_scale(cudaStream_t stream, int k, int m, int n, double* Wt, double* H, double* norms,
                         int* status, double* partial, int num_sms, double* HHt = nullptr, double* HAt = nullptr);

// ---- spmm.cu --------------------------------------------------------------
struct SparseDev;
// out (k x ncols, ld = ldo) : out(:,j) = sum over the entries (idx, val) of compressed column j of val * B(:, idx)
// with B k x * column-major. Serves W'A (CSC of A, B = Wt) and H A' (CSR of A, B = H); entries are
// added in storage order, which is the order the reference adds them (sparse_gemm_ba_impl.hpp:26-140).
void spmm_gather(cudaStream_t stream, int ncols, const unsigned int* ptr, const unsigned int* idx, const double* val,
                 int k, const double* B, long long ldb, double alpha, double beta, double* out, long long ldo, int num_sms);
struct SegTable;
constexpr int kSpmmSeg = 512;
constexpr int kSpmmMaxK = 256;        // rows of the dense operand one kernel pass holds per warp; a wider operand goes in row blocks
// Segment table of a compressed matrix (ncols compressed columns, offsets ptr[ncols + 1]); synchronises the stream.
void build_segments(cudaStream_t stream, int ncols, const unsigned int* ptr, SegTable& T, int num_sms);
// Same product as spmm_gather, one work item per segment; columns cut into several segments are summed from
// per-segment partials in segment order (deterministic; differs from the one-pass order at rounding level only).
// partial: at least T.nslots * k doubles. ngather = number of vectors idx can address (m or n).
void spmm_gather_seg(cudaStream_t stream, int ncols, const SegTable& T, const unsigned int* idx, const double* val,
                     int k, const double* B, long long ldb, double alpha, double beta, double* out, long long ldo,
                     double* partial, int num_sms, int ngather);
// Builds rowptr/colidx/valr (CSR = stable transpose) from the CSC arrays already on the device.
void build_csr(cudaStream_t stream, SparseDev& S, bool keep_scratch = false);

} // namespace smk
