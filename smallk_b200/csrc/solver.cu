// smallk_b200 — the NMF solvers as sequences of kernels on one stream.
//
// Mirrors the solver functors of the reference (Init + operator()):
//   Solver_Generic_BPP      common/include/nmf_solver_bpp.hpp:301-383
//   Solver_Generic_MU       common/include/nmf_solver_mu.hpp:74-169
//   Solver_Generic_HALS_Da  common/include/nmf_solver_hals.hpp:120-207
//   Solver_Generic_Rank2    common/include/nmf_solver_rank2.hpp:321-461
// and the progress estimators of progress_estimator_generic.hpp:30-108.
//
// Layout decision: both factors live on the device as "k x big" column-major
// matrices — H (k x n) as in the reference and Wt = W' (k x m). With that, the
// W side of every solver is the H side with (A, W, H) -> (A', H', W'): the same
// GEMM, NNLS and update kernels serve both, the reference's explicit transposed
// copy of A (nmf_solver_bpp.hpp:319; +8mn bytes) and its per-iteration
// W <-> Wt transposes (:363-364) disappear, and W is transposed only when it
// crosses the host boundary.
#include <cstdlib>
#include "context.h"
#include "solver.h"

namespace smk {

namespace {

// SMK_RANK2_FUSED=0 keeps the generic kernel sequence for sparse rank-2 (A/B measurements, parity diagnostics)
bool rank2_fused_enabled()
{
    const char* e = getenv("SMK_RANK2_FUSED");
    return !(e && atoi(e) == 0);
}

void nccl_check(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess) throw std::string(what) + ": " + ncclGetErrorString(r);
}

// sum over the ranks, in place: our own one-shot kernel over peer memory, or NCCL (SMK_PEER=0)
void allreduce_sum(smk_ctx* c, double* buf, size_t count)
{
    if (c->nranks <= 1) return;
    if (c->use_peer) peer_allreduce(c, buf, static_cast<int>(count), nullptr, nullptr, 0, nullptr, nullptr);
    else nccl_check(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, c->comm, c->stream), "ncclAllReduce");
}

// SMK_PHASES=1: a cudaEvent between the phases of solver_step; solver_phase_report() sums the gaps per name
struct Phase
{
    smk_ctx* c;
    explicit Phase(smk_ctx* ctx) : c(ctx) { mark("begin"); }
    void mark(const char* name)
    {
        if (!c->phases_on) return;
        if (c->phase_pool_used == c->phase_pool.size())
        {
            cudaEvent_t e;
            SMK_CUDA(cudaEventCreate(&e));
            c->phase_pool.push_back(e);
        }
        cudaEvent_t e = c->phase_pool[c->phase_pool_used++];
        SMK_CUDA(cudaEventRecord(e, c->stream));
        c->phase_marks.emplace_back(name, e);
    }
};

// ---- the products of one outer iteration ---------------------------------
// WtA (k x n) = Wt * A
void prod_WtA(smk_ctx* c)
{
    const int k = c->opts.k;
    if (c->has_dense)
        gemm_f64(c->stream, false, k, c->n, c->m, c->Wt.p, k, c->dA, c->ldA, c->WtA.p, k, nullptr, 0,
                 c->ws.p, c->ws.n * sizeof(double), c->num_sms);
    else
        spmm_gather_seg(c->stream, c->n, c->Sa->seg_cols, c->Sa->rowidx.p, c->Sa->val.p, k, c->Wt.p, k, 1.0, 0.0, c->WtA.p, k,
                        c->spmm_partial.p, c->num_sms, c->m);
}

// HAt (k x m) = H * A'   (summed over the column shards when running multi-GPU)
void prod_HAt(smk_ctx* c)
{
    const int k = c->opts.k;
    const long long piece = static_cast<long long>(k) * c->x_loc;
    if (c->use_peer && c->has_dense)
    {
        // the GEMM leaves its split-R partial tiles in the workspace; the reduce-scatter kernel adds them in split order
        // WHILE it stores each row block into its owner's receive slot (peer.cu): reduction pass = NVLink transfer
        if (peer_fused_by_env())
        {
            // ONE kernel: product, split-R reduction by the last-arriving CTA of each tile, and that CTA stores the finished tile
            // straight into the receive slot of the rank owning those rows of W; the last tile publishes the epoch (gemm_f64.cu)
            const GemmScatter sc = peer_scatter_begin(c, k, c->x_loc);
            gemm_f64(c->stream, true, k, c->m, c->n, c->H.p, k, c->dA, c->ldA, c->HAt.p, k, nullptr, 0,
                     c->ws.p, c->ws.n * sizeof(double), c->num_sms, nullptr, &sc);
            peer_reduce_scatter_finish(c, piece, c->HAt.p + c->rank * piece);
        }
        else
        {
            int splits = 0;
            gemm_f64(c->stream, true, k, c->m, c->n, c->H.p, k, c->dA, c->ldA, c->HAt.p, k, nullptr, 0,
                     c->ws.p, c->ws.n * sizeof(double), c->num_sms, &splits);
            peer_reduce_scatter(c, gemm_partials(c->ws.p), splits, static_cast<long long>(k) * c->m, piece, c->HAt.p + c->rank * piece);
        }
        if (!c->w_sharded) peer_allgather(c, 1, piece);
        return;
    }
    if (c->has_dense)
        gemm_f64(c->stream, true, k, c->m, c->n, c->H.p, k, c->dA, c->ldA, c->HAt.p, k, nullptr, 0,
                 c->ws.p, c->ws.n * sizeof(double), c->num_sms);
    else
        spmm_gather_seg(c->stream, c->m, c->Sa->seg_rows, c->Sa->colidx.p, c->Sa->valr.p, k, c->H.p, k, 1.0, 0.0, c->HAt.p, k,
                        c->spmm_partial.p, c->num_sms, c->n);
    if (c->nranks <= 1) return;
    if (c->use_peer)
    {
        peer_reduce_scatter(c, c->HAt.p, 1, static_cast<long long>(k) * c->m, piece, c->HAt.p + c->rank * piece);
        if (!c->w_sharded) peer_allgather(c, 1, piece);
    }
    else if (c->w_sharded)
    {
        // row-sharded W update: every rank receives the sum of its own k x m_loc slice only
        nccl_check(ncclReduceScatter(c->HAt.p, c->HAt.p + c->rank * piece, piece, ncclDouble, ncclSum, c->comm, c->stream), "ncclReduceScatter");
    }
    else nccl_check(ncclAllReduce(c->HAt.p, c->HAt.p, static_cast<size_t>(k) * c->m, ncclDouble, ncclSum, c->comm, c->stream), "ncclAllReduce");
}

// after a row-sharded W update: every rank's slice of Wt to every rank
void gather_Wt(smk_ctx* c)
{
    if (c->nranks <= 1 || !c->w_sharded) return;
    const long long piece = static_cast<long long>(c->opts.k) * c->m_loc;
    if (c->use_peer) peer_allgather(c, 0, piece);
    else nccl_check(ncclAllGather(c->Wt.p + c->rank * piece, c->Wt.p, piece, ncclDouble, c->comm, c->stream), "ncclAllGather");
}

// G (k x k) = X * X' for X k x q
void gram(smk_ctx* c, const double* X, int q, double* G)
{
    const int k = c->opts.k;
    gemm_f64(c->stream, true, k, k, q, X, k, X, k, G, k, nullptr, 0, c->ws.p, c->ws.n * sizeof(double), c->num_sms);
}
// the same on the side stream, with the side stream's own split-R workspace (the main stream's is busy with a big product)
void gram_side(smk_ctx* c, const double* X, int q, double* G)
{
    const int k = c->opts.k;
    gemm_f64(c->side, true, k, k, q, X, k, X, k, G, k, nullptr, 0, c->ws_side.p, c->ws_side.n * sizeof(double), c->num_sms);
}

// out (k x q) = G (k x k) * X (k x q) - R   (R may be null)
void gram_times(smk_ctx* c, const double* G, const double* X, int q, const double* R, double* out)
{
    const int k = c->opts.k;
    gemm_f64(c->stream, false, k, q, k, G, k, X, k, out, k, R, k, nullptr, 0, c->num_sms);
}

// ---- work overlapped with the big products: a side stream (highest priority) forked from / joined to the main stream -------
void side_begin(smk_ctx* c, smk_ctx::InvBuf& b)
{
    if (!c->side)
    {
        int lo = 0, hi = 0;
        SMK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SMK_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
    }
    if (!b.fork)
    {
        SMK_CUDA(cudaEventCreateWithFlags(&b.fork, cudaEventDisableTiming));
        SMK_CUDA(cudaEventCreateWithFlags(&b.join, cudaEventDisableTiming));
    }
    SMK_CUDA(cudaEventRecord(b.fork, c->stream));
    SMK_CUDA(cudaStreamWaitEvent(c->side, b.fork, 0));
}
void side_end(smk_ctx* c, smk_ctx::InvBuf& b)
{
    SMK_CUDA(cudaEventRecord(b.join, c->side));
    b.pending = true;
}
void side_join(smk_ctx* c, smk_ctx::InvBuf& b)
{
    if (!b.pending) return;
    SMK_CUDA(cudaStreamWaitEvent(c->stream, b.join, 0));
    b.pending = false;
}
// G^-1 for the NNLS solves against G (BPP, k > 32), on the side stream, between side_begin and side_end
void inverse_side(smk_ctx* c, const double* G, smk_ctx::InvBuf& b)
{
    const int k = c->opts.k;
    b.valid = false;
    if (c->opts.algorithm != SMK_BPP || !nnls_uses_inverse(k)) return;
    b.Ginv.reserve(static_cast<size_t>(k) * k);
    b.ok.reserve(1);
    nnls_prepare_inverse(c->side, k, G, k, b.Ginv.p, b.ok.p);
    b.valid = true;
}
// BPP, k > 32: G^-1 for the next NNLS solve against G — it runs under the big product that the main stream launches next
// (the inverse was a fixed ~0.19 ms inside every NNLS launch when each CTA formed it for itself)
void prepare_inverse(smk_ctx* c, const double* G, smk_ctx::InvBuf& b)
{
    if (c->opts.algorithm != SMK_BPP || !nnls_uses_inverse(c->opts.k)) return;
    side_begin(c, b);
    inverse_side(c, G, b);
    side_end(c, b);
}
// BPP: may the Gram matrices (+ their rank sums + inverses) run on the side stream under the big products? Not with the NCCL
// exchange (collectives must be issued in one order on one stream); SMK_SIDE_GRAM=0 turns it off (measurements).
bool side_gram_enabled(const smk_ctx* c)
{
    if (c->opts.algorithm != SMK_BPP || (c->nranks > 1 && !c->use_peer)) return false;
    const char* e = getenv("SMK_SIDE_GRAM");
    return !(e && atoi(e) == 0);
}

void compute_HHt(smk_ctx* c)
{
    gram(c, c->H.p, c->n, c->HHt.p);
    allreduce_sum(c, c->HHt.p, static_cast<size_t>(c->opts.k) * c->opts.k);
    prepare_inverse(c, c->HHt.p, c->invW);
}
void compute_WtW(smk_ctx* c)
{
    if (c->w_sharded && c->use_peer)
    {
        // every rank Grams the rows of W it has just updated; the k x k partials are summed in rank order (peer.cu)
        gram(c, c->Wt.p + static_cast<size_t>(c->opts.k) * c->w_row0(), c->w_rows(), c->WtW.p);
        allreduce_sum(c, c->WtW.p, static_cast<size_t>(c->opts.k) * c->opts.k);
    }
    else if (c->grad_sharded)
    {
        // replicated W (HALS): every rank Grams its block of rows, the partials are summed in rank order — one eighth of the
        // k x k x m product per rank at 8 GPUs instead of all of it on every rank
        gram(c, c->Wt.p + static_cast<size_t>(c->opts.k) * c->g_row0(), c->g_rows(), c->WtW.p);
        allreduce_sum(c, c->WtW.p, static_cast<size_t>(c->opts.k) * c->opts.k);
    }
    else gram(c, c->Wt.p, c->m, c->WtW.p);
    prepare_inverse(c, c->WtW.p, c->invH);
}

void run_nnls(smk_ctx* c, const double* LHS, const double* RHS, double* X, double* Y, int q, smk_ctx::InvBuf& inv)
{
    const int k = c->opts.k;
    const double* Ginv = nullptr;
    const int* ginv_ok = nullptr;
    side_join(c, inv);
    if (inv.valid) { Ginv = inv.Ginv.p; ginv_ok = inv.ok.p; }
    nnls_bpp(c->stream, k, q, LHS, k, RHS, k, X, k, Y, k, c->status.p, c->counter.p, c->deferred.p, c->steps_done, c->num_sms,
             Ginv, ginv_ok);
    // "zeroize everything iff any column was non-optimal" couples the column shards: one int, OR-reduced
    if (c->nranks > 1)
    {
        if (c->use_peer) peer_allreduce(c, c->acc.p + 6, 0, c->status.p + ST_ANY_NONOPT, nullptr, 0, nullptr, nullptr);
        else nccl_check(ncclAllReduce(c->status.p + ST_ANY_NONOPT, c->status.p + ST_ANY_NONOPT, 1, ncclInt, ncclMax, c->comm, c->stream), "ncclAllReduce");
    }
    nnls_bpp_finish(c->stream, k, q, X, k, Y, k, c->status.p, c->num_sms);
}

} // namespace

void solver_alloc(smk_ctx* c)
{
    const size_t k = c->opts.k, n = c->n;
    c->w_sharded = c->nranks > 1 && (c->opts.algorithm == SMK_BPP || c->opts.algorithm == SMK_MU);
    c->grad_sharded = c->nranks > 1 && c->opts.algorithm == SMK_HALS;
    // the one-shot exchange slots of peer.cu hold a k x k Gram matrix for k <= 256; larger ranks exchange through NCCL
    c->use_peer = c->nranks > 1 && peer_enabled_by_env() && static_cast<long long>(k) * static_cast<long long>(k) + 64 <= kPeerSmallCap;
    c->x_loc = c->nranks > 1 ? (c->m + c->nranks - 1) / c->nranks : c->m;       // rows per exchanged block
    c->m_loc = c->w_sharded ? c->x_loc : c->m;
    const bool padded = c->w_sharded || c->use_peer;
    const size_t m = padded ? static_cast<size_t>(c->x_loc) * c->nranks : c->m;     // padded row count of the k x m buffers
    c->H.reserve(k * n);
    c->gradH.reserve(k * n); c->gradWt.reserve(k * m);
    c->WtW.reserve(k * k); c->HHt.reserve(k * k);
    c->WtA.reserve(k * n);
    if (c->use_peer)
    {
        // Wt, HAt and the reduce-scatter receive slots live in the exchange region every peer maps (peer.cu)
        peer_setup(c, k * m);
        c->Wt.alias(peer_big_buffer(c, 0), k * m);
        c->HAt.alias(peer_big_buffer(c, 1), k * m);
    }
    else
    {
        if (!c->Wt.owned) c->Wt.release();
        if (!c->HAt.owned) c->HAt.release();
        c->Wt.reserve(k * m); c->HAt.reserve(k * m);
    }
    if (padded)
    {
        SMK_CUDA(cudaMemsetAsync(c->Wt.p, 0, k * m * sizeof(double), c->stream));
        SMK_CUDA(cudaMemsetAsync(c->HAt.p, 0, k * m * sizeof(double), c->stream));
        SMK_CUDA(cudaMemsetAsync(c->gradWt.p, 0, k * m * sizeof(double), c->stream));
    }
    c->norms.reserve(k);
    if (c->has_sparse) c->spmm_partial.reserve(static_cast<size_t>(std::max(c->Sa->seg_cols.nslots, c->Sa->seg_rows.nslots)) * k + 1);
    c->deferred.reserve(nnls_deferred_bytes(static_cast<int>(std::max(m, n)), c->opts.k, c->num_sms));
    if (c->opts.algorithm == SMK_MU) { c->T1.reserve(k * n); c->T2.reserve(k * m); }    // m is the padded row count here
    if (c->opts.algorithm == SMK_RANK2 && c->has_sparse && c->nranks <= 1) c->T2.reserve(k * m);   // unnormalised W of the fused iteration
    if (c->opts.algorithm == SMK_HALS) c->T2.reserve(std::max(k * m, hals_sweep_scratch_doubles(static_cast<int>(m))));
    if (c->opts.prog_est_algorithm == SMK_DELTA_FNORM) c->Wprev.reserve(k * m);
    // split-R workspace: enough for the gram matrices at 4*SMs splits and for the big products at a few splits
    size_t want = std::max<size_t>(static_cast<size_t>(4 * c->num_sms) * k * k,
                                   std::min<size_t>(static_cast<size_t>(32) * k * std::max(m, n), (size_t(768) << 20) / sizeof(double)));
    c->ws.reserve(want + 8192);
    gemm_workspace_prepare(c->stream, c->ws.p, c->ws.n * sizeof(double));
    if (c->opts.algorithm == SMK_BPP)
    {
        c->ws_side.reserve(static_cast<size_t>(4 * c->num_sms) * k * k + 8192);
        gemm_workspace_prepare(c->stream, c->ws_side.p, c->ws_side.n * sizeof(double));
    }
    static const int init[ST_COUNT] = {0, INT_MAX, 0, 0, 0, 0, 0, 0};
    SMK_CUDA(cudaMemcpyAsync(c->status.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    c->prog.reserve(4);
    SMK_CUDA(cudaMemsetAsync(c->prog.p, 0, 4 * sizeof(double), c->stream));       // "pg0 not captured yet"
    const char* ph = getenv("SMK_PHASES");
    c->phases_on = ph && atoi(ph) != 0;
}

void solver_init(smk_ctx* c)
{
    switch (c->opts.algorithm)
    {
    case SMK_BPP:      // nmf_solver_bpp.hpp:310-335
    case SMK_MU:       // nmf_solver_mu.hpp:93-110
    case SMK_RANK2:    // nmf_solver_rank2.hpp:333-347
        compute_WtW(c);
        prod_WtA(c);
        break;
    case SMK_HALS:     // nmf_solver_hals.hpp:141-159
        compute_HHt(c);
        prod_HAt(c);
        break;
    }
    if (c->opts.prog_est_algorithm == SMK_DELTA_FNORM)   // ProgEstGenericDeltaW::Init
        SMK_CUDA(cudaMemcpyAsync(c->Wprev.p, c->Wt.p, sizeof(double) * c->opts.k * c->m, cudaMemcpyDeviceToDevice, c->stream));
}

void solver_step(smk_ctx* c)
{
    const int k = c->opts.k, m = c->m, n = c->n;
    c->pg_ready = false;
    c->status_cached = false;
    Phase ph(c);
    switch (c->opts.algorithm)
    {
    case SMK_BPP:      // nmf_solver_bpp.hpp:342-377
        run_nnls(c, c->WtW.p, c->WtA.p, c->H.p, c->gradH.p, n, c->invH);
        ph.mark("nnls_H");
        if (side_gram_enabled(c))
        {
            // H H' (+ its sum over the ranks + its inverse) on the side stream, under the H A' product
            side_begin(c, c->invW);
            gram_side(c, c->H.p, n, c->HHt.p);
            if (c->nranks > 1) peer_allreduce(c, c->HHt.p, k * k, nullptr, nullptr, 0, nullptr, nullptr, c->side);
            inverse_side(c, c->HHt.p, c->invW);
            side_end(c, c->invW);
        }
        else compute_HHt(c);
        ph.mark("HHt");
        prod_HAt(c);
        ph.mark("HAt");
        {
            const size_t off = static_cast<size_t>(k) * c->w_row0();      // 0 unless the W update is row-sharded
            run_nnls(c, c->HHt.p, c->HAt.p + off, c->Wt.p + off, c->gradWt.p + off, c->w_rows(), c->invW);
            ph.mark("nnls_W");
            if (side_gram_enabled(c))
            {
                // W'W from the rows this rank has just updated (+ rank sum + inverse) on the side stream, under the
                // all-gather of W and the W'A product
                side_begin(c, c->invH);
                gram_side(c, c->Wt.p + off, c->w_rows(), c->WtW.p);
                if (c->nranks > 1) peer_allreduce(c, c->WtW.p, k * k, nullptr, nullptr, 0, nullptr, nullptr, c->side);
                inverse_side(c, c->WtW.p, c->invH);
                side_end(c, c->invH);
                gather_Wt(c);
                ph.mark("gather_Wt");
            }
            else
            {
                gather_Wt(c);
                ph.mark("gather_Wt");
                compute_WtW(c);
            }
        }
        ph.mark("WtW");
        prod_WtA(c);
        ph.mark("WtA");
        side_join(c, c->invH);          // gradH needs W'W (a no-op when it was computed on the main stream)
        gram_times(c, c->WtW.p, c->H.p, n, c->WtA.p, c->gradH.p);
        ph.mark("gradH");
        break;
    case SMK_MU:       // nmf_solver_mu.hpp:118-163
        gram_times(c, c->WtW.p, c->H.p, n, nullptr, c->T1.p);
        mu_update(c->stream, static_cast<long long>(k) * n, c->H.p, c->WtA.p, c->T1.p);
        compute_HHt(c);
        prod_HAt(c);
        {
            const size_t off = static_cast<size_t>(k) * c->w_row0();
            const int rows = c->w_rows();
            if (rows > 0)
            {
                gram_times(c, c->HHt.p, c->Wt.p + off, rows, nullptr, c->T2.p + off);
                mu_update(c->stream, static_cast<long long>(k) * rows, c->Wt.p + off, c->HAt.p + off, c->T2.p + off);
            }
            gather_Wt(c);
            prod_WtA(c);
            compute_WtW(c);
            if (rows > 0) gram_times(c, c->HHt.p, c->Wt.p + off, rows, c->HAt.p + off, c->gradWt.p + off);
        }
        gram_times(c, c->WtW.p, c->H.p, n, c->WtA.p, c->gradH.p);
        break;
    case SMK_HALS:     // nmf_solver_hals.hpp:166-199
        hals_sweep(c->stream, k, m, c->Wt.p, c->HHt.p, c->HAt.p, /*normalize=*/true, c->norms.p, c->partial.p, c->num_sms, c->T2.p);
        ph.mark("hals_W");
        compute_WtW(c);
        ph.mark("WtW");
        prod_WtA(c);
        ph.mark("WtA");
        hals_sweep(c->stream, k, n, c->H.p, c->WtW.p, c->WtA.p, /*normalize=*/false, c->norms.p, c->partial.p, c->num_sms);
        ph.mark("hals_H");
        gram_times(c, c->WtW.p, c->H.p, n, c->WtA.p, c->gradH.p);
        ph.mark("gradH");
        compute_HHt(c);
        ph.mark("HHt");
        prod_HAt(c);
        ph.mark("HAt");
        if (c->grad_sharded)
        {
            // gradW only feeds the projected-gradient sum: this rank's row block is enough (solver_progress_enqueue sums over the ranks)
            const size_t off = static_cast<size_t>(k) * c->g_row0();
            if (c->g_rows() > 0) gram_times(c, c->HHt.p, c->Wt.p + off, c->g_rows(), c->HAt.p + off, c->gradWt.p + off);
        }
        else gram_times(c, c->HHt.p, c->Wt.p, m, c->HAt.p, c->gradWt.p);
        ph.mark("gradW");
        break;
    case SMK_RANK2:    // nmf_solver_rank2.hpp:353-455
        if (c->has_sparse && c->nranks <= 1 && rank2_fused_enabled())
        {
            rank2_fused_step(c, c->T2.p);     // the same iteration in three kernels, projected-gradient sums included
            ph.mark("rank2_fused");
            c->steps_done += 1;
            c->pg_ready = true;
            return;
        }
        rank2_update(c->stream, n, c->H.p, c->WtW.p, c->WtA.p, false, c->status.p, c->steps_done);
        compute_HHt(c);
        prod_HAt(c);
        rank2_update(c->stream, m, c->Wt.p, c->HHt.p, c->HAt.p, true, c->status.p, c->steps_done);
        // NormalizeAndScale(W, H, s); HHt, AHt rescaled analytically (:420-441)
        normalize_and_scale(c->stream, 2, m, n, c->Wt.p, c->H.p, c->norms.p, c->status.p, c->partial.p, c->num_sms,
                            c->HHt.p, c->HAt.p);
        gram_times(c, c->HHt.p, c->Wt.p, m, c->HAt.p, c->gradWt.p);
        compute_WtW(c);
        prod_WtA(c);
        gram_times(c, c->WtW.p, c->H.p, n, c->WtA.p, c->gradH.p);
        break;
    }
    ph.mark("rest");
    c->steps_done += 1;
}

// ProgressEst::Update for the state after `steps_done` iterations, entirely on the device: the metric is written to
// *metric_dev (device memory); no host synchronisation. pg0 lives in c->prog (captured by the first PG evaluation after
// solver_begin, whenever that is).
void solver_progress_enqueue(smk_ctx* c, double* metric_dev)
{
    const long long k = c->opts.k;
    Phase ph(c);
    if (c->opts.prog_est_algorithm == SMK_PG_RATIO)
    {
        // projected_gradient.hpp:125-171. The H part is a sum over this rank's columns; the W part is a sum over this rank's
        // rows when the W update is row-sharded, else every rank holds all of gradW and rank 0 alone contributes it.
        const long long woff = c->w_sharded ? k * c->w_row0() : (c->grad_sharded ? k * c->g_row0() : 0);
        long long wcount = c->w_sharded ? k * c->w_rows() : (c->grad_sharded ? k * c->g_rows() : k * c->m);
        if (c->nranks > 1 && !c->w_sharded && !c->grad_sharded && c->rank != 0) wcount = 0;
        double* prog = c->nranks <= 1 ? c->prog.p : nullptr;        // one rank: the reduction kernel finishes the metric itself
        if (!c->pg_ready)
            pg_pair(c->stream, wcount, c->gradWt.p + woff, c->Wt.p + woff, k * c->n, c->gradH.p, c->H.p, c->partial.p,
                    c->ticket.p + 4, c->acc.p, prog, metric_dev, c->status.p, c->num_sms);
        else if (c->nranks <= 1)
            progress_metric_launch(c->stream, 0, c->acc.p, c->prog.p, metric_dev, c->status.p);
        if (c->nranks > 1)
        {
            if (c->use_peer)    // sums, the failure flag and the metric in one exchange kernel
                peer_allreduce(c, c->acc.p, 2, nullptr, c->status.p + ST_FAIL_ITER, 0, c->prog.p, metric_dev);
            else
            {
                nccl_check(ncclAllReduce(c->acc.p, c->acc.p, 2, ncclDouble, ncclSum, c->comm, c->stream), "ncclAllReduce");
                nccl_check(ncclAllReduce(c->status.p + ST_FAIL_ITER, c->status.p + ST_FAIL_ITER, 1, ncclInt, ncclMin, c->comm, c->stream), "ncclAllReduce");
                progress_metric_launch(c->stream, 0, c->acc.p, c->prog.p, metric_dev, c->status.p);
            }
        }
    }
    else
    {
        // progress_estimator_generic.hpp:58-69 (W is replicated or all-gathered: every rank computes the same numbers)
        diff_sumsq(c->stream, k * c->m, c->Wprev.p, c->Wt.p, c->partial.p, c->acc.p + 0, c->num_sms);
        diff_sumsq(c->stream, k * c->m, c->Wt.p, nullptr, c->partial.p, c->acc.p + 1, c->num_sms);
        SMK_CUDA(cudaMemcpyAsync(c->Wprev.p, c->Wt.p, sizeof(double) * k * c->m, cudaMemcpyDeviceToDevice, c->stream));
        progress_metric_launch(c->stream, 1, c->acc.p, c->prog.p, metric_dev, c->status.p);
    }
    ph.mark("progress");
}

namespace {
// the status words as of now, through page-locked memory; `extra` doubles from extra_dev ride on the same synchronisation
void read_status(smk_ctx* c, const double* extra_dev, int extra)
{
    int* st_pinned = reinterpret_cast<int*>(c->pinned + 8);
    if (extra > 0) SMK_CUDA(cudaMemcpyAsync(c->pinned, extra_dev, extra * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaMemcpyAsync(st_pinned, c->status.p, sizeof(c->status_host), cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < ST_COUNT; ++i) c->status_host[i] = st_pinned[i];
    c->status_cached = true;
}

int status_verdict(smk_ctx* c)
{
    if (c->status_host[ST_COMM_TIMEOUT]) { c->err = "peer exchange timed out (a rank died or fell more than 60 s behind)"; return SMK_FAILURE; }
    if (c->status_host[ST_PG_NAN]) { c->err = "ProjectedGradientNorm: NaN"; return SMK_FAILURE; }
    return SMK_OK;
}
} // namespace

// ProgressEst::Update + the metric on the host. Synchronises the stream.
int solver_progress(smk_ctx* c, double* metric)
{
    solver_progress_enqueue(c, c->prog.p + 2);
    read_status(c, c->prog.p + 2, 1);
    *metric = c->pinned[0];
    return status_verdict(c);
}

// `count` outer iterations, each followed by its progress update, with no host synchronisation in between; the metrics
// come back in one copy at the end (metrics_host may be null). Returns SMK_FAILURE if a solver step failed.
int solver_run(smk_ctx* c, int count, double* metrics_host)
{
    c->trace.reserve(static_cast<size_t>(std::max(count, 1)));
    for (int i = 0; i < count; ++i)
    {
        solver_step(c);
        solver_progress_enqueue(c, c->trace.p + i);
    }
    if (c->nranks > 1 && c->opts.prog_est_algorithm != SMK_PG_RATIO)
    {
        // the failure flag did not ride on a progress exchange: all ranks must take the same exit
        if (c->use_peer) peer_allreduce(c, c->acc.p + 6, 0, nullptr, c->status.p + ST_FAIL_ITER, 0, nullptr, nullptr);
        else nccl_check(ncclAllReduce(c->status.p + ST_FAIL_ITER, c->status.p + ST_FAIL_ITER, 1, ncclInt, ncclMin, c->comm, c->stream), "ncclAllReduce");
    }
    if (metrics_host && count > 0)
        SMK_CUDA(cudaMemcpyAsync(metrics_host, c->trace.p, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    read_status(c, nullptr, 0);
    if (c->status_host[ST_FAIL_ITER] != INT_MAX) { c->err = "NMF solver failure on iteration " + std::to_string(c->status_host[ST_FAIL_ITER] + 1); return SMK_FAILURE; }
    if (c->opts.algorithm == SMK_RANK2 && c->status_host[ST_NORM_EPS]) { c->err = "Normalize: column norm < machine epsilon"; return SMK_FAILURE; }
    return status_verdict(c);
}

// "name=ms;name=ms;..." summed over the solver steps since the last report (SMK_PHASES=1), then reset. Synchronises.
std::string solver_phase_report(smk_ctx* c)
{
    std::string out;
    if (!c->phases_on || c->phase_marks.empty()) return out;
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<std::pair<std::string, double>> acc;
    for (size_t i = 1; i < c->phase_marks.size(); ++i)
    {
        const char* name = c->phase_marks[i].first;
        if (std::string(name) == "begin") continue;
        float ms = 0.f;
        SMK_CUDA(cudaEventElapsedTime(&ms, c->phase_marks[i - 1].second, c->phase_marks[i].second));
        bool found = false;
        for (auto& a : acc) if (a.first == name) { a.second += ms; found = true; break; }
        if (!found) acc.emplace_back(name, ms);
    }
    for (auto& a : acc) out += a.first + "=" + std::to_string(a.second) + ";";
    c->phase_marks.clear();
    c->phase_pool_used = 0;
    return out;
}

// NormalizeAndScale(W, H): normalize.hpp:118-138. Returns SMK_FAILURE if a column norm < eps.
int solver_normalize(smk_ctx* c)
{
    const int k = c->opts.k;
    normalize_and_scale(c->stream, k, c->m, c->n, c->Wt.p, c->H.p, c->norms.p, c->status.p, c->partial.p, c->num_sms);
    int st[ST_COUNT];
    SMK_CUDA(cudaMemcpyAsync(st, c->status.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    if (st[ST_NORM_EPS]) { c->err = "Normalize: column norm < machine epsilon"; return SMK_FAILURE; }
    return SMK_OK;
}

// NnlsHals (nnls.hpp:249-316): H-only HALS against a fixed W. Expects solver_alloc done with algorithm HALS and
// Wt, H uploaded. Returns SMK_OK on convergence (W, H normalised), SMK_FAILURE at the iteration limit.
int solver_nnls_hals(smk_ctx* c, double tol, int max_iter, int* iterations)
{
    const int k = c->opts.k, n = c->n;
    compute_WtW(c);
    prod_WtA(c);
    double pg0 = 0.0;
    int it = 0;
    for (int i = 0; i < max_iter; ++i)
    {
        it = i + 1;
        hals_sweep(c->stream, k, n, c->H.p, c->WtW.p, c->WtA.p, /*normalize=*/false, c->norms.p, c->partial.p, c->num_sms);
        gram_times(c, c->WtW.p, c->H.p, n, c->WtA.p, c->gradH.p);
        pg_sumsq(c->stream, static_cast<long long>(k) * n, c->gradH.p, c->H.p, c->partial.p, c->acc.p, c->num_sms);
        double h = 0.0;
        SMK_CUDA(cudaMemcpyAsync(&h, c->acc.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        const double pg = sqrt(h);
        if (pg != pg) { c->err = "ProjectedGradientNorm: NaN"; if (iterations) *iterations = it; return SMK_FAILURE; }
        if (i == 0) { pg0 = pg; continue; }
        if (pg < tol * pg0)
        {
            if (iterations) *iterations = it;
            return solver_normalize(c);
        }
    }
    if (iterations) *iterations = it;
    c->err = "NNLS solver reached iteration limit.";
    return SMK_FAILURE;
}

void solver_product(smk_ctx* c, int which) { if (which == 0) prod_WtA(c); else prod_HAt(c); }

// First outer iteration (0-based) in which a kernel reported solver failure, or INT_MAX. Synchronises unless the last
// solver_progress already brought the (rank-reduced) status words to the host.
int solver_fail_iter(smk_ctx* c)
{
    const bool reduced_by_progress = c->nranks <= 1 || c->opts.prog_est_algorithm == SMK_PG_RATIO;
    if (c->status_cached && reduced_by_progress) return c->status_host[ST_FAIL_ITER];
    if (c->nranks > 1)      // all ranks must take the same exit
    {
        if (c->use_peer) peer_allreduce(c, c->acc.p + 6, 0, nullptr, c->status.p + ST_FAIL_ITER, 0, nullptr, nullptr);
        else nccl_check(ncclAllReduce(c->status.p + ST_FAIL_ITER, c->status.p + ST_FAIL_ITER, 1, ncclInt, ncclMin, c->comm, c->stream), "ncclAllReduce");
    }
    read_status(c, nullptr, 0);
    return c->status_host[ST_FAIL_ITER];
}

} // namespace smk
