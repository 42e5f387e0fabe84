// smallk_b200 — the context behind the C ABI: device buffers, the loaded matrix, solver state.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>
#include <string>
#include <vector>
#include <climits>

#include "common.cuh"
#include "kernels.h"
#include "peer.h"
#include "../../include/smallk_b200.h"

namespace smk {

template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p && owned) cudaFree(p); p = nullptr; n = 0; owned = true; }
    // view of memory owned by somebody else (the NVLink exchange region, peer.cu)
    void alias(T* ptr, size_t count) { release(); p = ptr; n = count; owned = false; }
    // grow-only allocation; contents are not preserved across a growth
    void reserve(size_t count)
    {
        if (count <= n) return;
        release();
        SMK_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        n = count;
    }
    void zero(cudaStream_t s) { if (p) SMK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// Work decomposition of a compressed matrix for the gather SpMM: every compressed column is cut into segments of at
// most kSpmmSeg stored entries, so that one heavy column (a popular term, a hub node) cannot serialise the kernel.
struct SegTable
{
    int nseg = 0, nslots = 0, nmulti = 0;
    DevBuf<unsigned int> col, beg, end, slot;   // per segment: its column, entry range, partial slot (0xFFFFFFFF: writes the output directly)
    DevBuf<unsigned int> first_slot;            // per column (+1): first partial slot
    DevBuf<unsigned int> multi_col;             // the columns that have more than one segment
    DevBuf<unsigned int> t_cnt, t_first, t_mpos; // build temporaries (kept: the column-subset matrix is rebuilt many times)
    DevBuf<unsigned char> t_scan;
};

struct SparseDev
{
    int m = 0, n = 0;
    unsigned int nnz = 0;
    // CSC as given (duplicates and unsorted rows preserved)
    DevBuf<unsigned int> colptr, rowidx;
    DevBuf<double> val;
    // CSR = CSC of A', built by a stable counting sort (sparse_matrix_ops.hpp:37-126)
    DevBuf<unsigned int> rowptr, colidx;
    DevBuf<double> valr;
    // build_csr temporaries; kept between calls only for the column-subset matrix, which is rebuilt many times
    DevBuf<unsigned int> t_colof, t_ids, t_ids_sorted, t_rows_sorted;
    DevBuf<unsigned char> t_sort;
    SegTable seg_cols, seg_rows;                // segments of the CSC columns / of the CSR rows
};

} // namespace smk

struct smk_ctx
{
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    // ---- the matrix
    bool has_dense = false, has_sparse = false;
    const double* dA = nullptr;     // dense, column-major
    long long ldA = 0;
    int m = 0, n = 0;
    smk::DevBuf<double> A_store;
    smk::SparseDev S;               // the loaded sparse matrix
    smk::SparseDev* Sa = &S;        // the ACTIVE sparse matrix: &S, or &Ssub after smk_select_columns

    // ---- column subset (hierclust: SubMatrixColsCompact), see submatrix.cu
    bool subset_active = false;
    const double* full_dA = nullptr;
    long long full_ldA = 0;
    int full_m = 0, full_n = 0;
    smk::SparseDev Ssub;
    smk::DevBuf<double> A_sub;
    smk::DevBuf<unsigned int> sub_cols, sub_flags, sub_o2n, sub_n2o;
    smk::DevBuf<unsigned char> sub_scan_tmp;

    // ---- scratch shared by kernels
    smk::DevBuf<double> ws;         // split-R partial tiles
    smk::DevBuf<int> status;        // ST_COUNT ints
    smk::DevBuf<unsigned int> counter;
    smk::DevBuf<unsigned int> ticket;   // arrival tickets of the fused rank-2 kernels (self-resetting; zeroed once)
    smk::DevBuf<unsigned char> deferred; // BPP columns handed from the fast to the slow NNLS kernel
    smk::DevBuf<double> partial;    // 1024 block partials
    smk::DevBuf<double> acc;        // 8 scalars
    smk::DevBuf<double> io;         // staging for host<->device transposes
    smk::DevBuf<double> sort_keys, sort_keys_out;       // smk_argsort_desc / smk_sort_desc
    smk::DevBuf<int> sort_vals, sort_vals_out;
    smk::DevBuf<unsigned char> sort_tmp;
    smk::DevBuf<double> spmm_partial; // partial sums of the segmented SpMM (columns with more than one segment)

    // ---- solver state (both factors are kept "k x big", column-major: H is k x n, Wt = W' is k x m)
    bool active = false;
    smk_nmf_options opts;
    int steps_done = 0;
    double pg0 = 0.0;
    smk::DevBuf<double> H, Wt, gradH, gradWt, WtW, HHt, WtA, HAt, T1, T2, Wprev, norms;
    smk::DevBuf<double> prog;       // ProgressEst state on the device: [0] = pg0, [1] = "pg0 captured", [2] = metric of the last update
    smk::DevBuf<double> trace;      // metric per iteration of smk_solver_run / of the device-side loop (nmf_loop.cu)
    smk::DevBuf<int> loop_state;    // device-side loop of smk_nmf: iteration, success count, outcome
    bool pg_ready = false;          // the last solver_step left both projected-gradient sums in acc[0..1] (fused rank-2)
    bool status_cached = false;     // status_host holds the status words as of the last solver_progress
    int status_host[smk::ST_COUNT] = {0, INT_MAX, 0, 0, 0};
    double* pinned = nullptr;       // 16 doubles of page-locked host memory: per-iteration readbacks (PG sums, status words)
    float last_ms = 0.f;
    long long last_launches = 0;

    // ---- multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // W-side row sharding (BPP, MU): rank g owns rows [g*m_loc, min(m, (g+1)*m_loc)) of W; k x m buffers are padded to
    // k x (m_loc * nranks) so that reduce-scatter / all-gather move equal pieces. m_loc == m when nranks == 1.
    int m_loc = 0;
    bool w_sharded = false;
    // NVLink peer-memory exchange (peer.cu): on by default when nranks > 1, SMK_PEER=0 selects the NCCL collectives.
    // x_loc = rows per exchanged block of a k x m buffer (= m_loc when the W update is row-sharded).
    bool use_peer = false;
    int x_loc = 0;
    smk::PeerComm peer;
    smk::DevBuf<unsigned int> peer_ticket;
    // G^-1 for the NNLS solves of BPP (nnls_prepare_inverse), computed on the side stream under the big product that precedes
    // each solve: invH = (W'W)^-1 for the H side, invW = (H H')^-1 for the W side
    struct InvBuf
    {
        smk::DevBuf<double> Ginv;
        smk::DevBuf<int> ok;
        cudaEvent_t fork = nullptr, join = nullptr;
        bool pending = false;       // side-stream work recorded in `join` has not been waited for by the main stream yet
        bool valid = false;         // Ginv / ok belong to the current Gram matrix
    } invH, invW;
    smk::DevBuf<double> ws_side;    // split-R workspace of the Gram matrices computed on the side stream
    cudaStream_t side = nullptr;        // helper stream (highest priority) for work overlapped with the big products
    // per-phase device times of solver_step (SMK_PHASES=1): cudaEvent pairs, summed on demand
    bool phases_on = false;
    std::vector<std::pair<const char*, cudaEvent_t>> phase_marks;
    std::vector<cudaEvent_t> phase_pool;
    size_t phase_pool_used = 0;
    // HALS on several GPUs: the W sweep is replicated (its in-sweep norms couple the rows), but W'W and gradW are formed from this
    // rank's block of x_loc rows only (Gram partials summed over the ranks, PG sum over own rows)
    bool grad_sharded = false;
    int g_row0() const { return rank * x_loc; }
    int g_rows() const { const int r = m - rank * x_loc; return r < 0 ? 0 : (r < x_loc ? r : x_loc); }
    int w_row0() const { return rank * m_loc; }
    int w_rows() const { const int r = m - rank * m_loc; return r < 0 ? 0 : (r < m_loc ? r : m_loc); }
};
