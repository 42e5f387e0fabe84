// smallk_b200 — NmfSolve's stop-tested iterations as ONE CUDA graph with a device-side WHILE loop.
//
// The loop of NmfSolve (common/include/nmf_solve_generic.hpp:67-123) tests the progress metric after every iteration from
// min_iter on: with a host loop that is one stream synchronisation + ~20 launches per iteration, which is what an iteration
// COSTS when the matrices are small (the 202 rank-2 factorizations of a 64-leaf hierclust run: 13 800 iterations of ~40 us of
// kernels each; the 256 x 256 command-line case). Here one outer iteration (solver step + progress update) is captured once into
// the body of a conditional WHILE node (CUDA 12.4+); a one-thread kernel at the end of the body applies the reference's stop rule
// (metric <= tol for tolcount consecutive evaluations; solver failure; NaN; max_iter) and sets the loop condition on the device.
// The host launches the graph once and reads the outcome once. Iteration counts, factors and metrics are those of the host loop.
//
// Used when there is one rank (the peer exchange kernels carry host-side epoch counters in their arguments, which a replayed graph
// cannot advance), the caller did not ask for live progress lines, and phase timing is off. SMK_GRAPH=0 turns it off.
#include <cstdlib>
#include "context.h"
#include "solver.h"

namespace smk {

namespace {

struct LoopState { int iter, success_count, result, fail_iter; };      // result: 0 running / ran out, 1 converged, 2 solver failure, 3 NaN or exchange failure, 4 norm < eps

__global__ void nmf_loop_control_kernel(cudaGraphConditionalHandle handle, LoopState* st, const int* __restrict__ status, const double* __restrict__ metric,
                                        double* __restrict__ trace, double tol, int tolcount, int max_iter, int is_rank2)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned int cont = 1u;
    const int iter = st->iter;
    const double mv = *metric;
    trace[iter] = mv;
    if (status[ST_PG_NAN] || status[ST_COMM_TIMEOUT]) { st->result = 3; cont = 0u; }
    else if (status[ST_FAIL_ITER] != INT_MAX) { st->result = 2; st->fail_iter = iter; cont = 0u; }
    else if (is_rank2 && status[ST_NORM_EPS]) { st->result = 4; cont = 0u; }
    else if (mv <= tol) { if (++st->success_count >= tolcount) { st->result = 1; cont = 0u; } }
    else st->success_count = 0;
    if (cont)
    {
        st->iter = iter + 1;
        if (iter + 1 >= max_iter) cont = 0u;
    }
    cudaGraphSetConditional(handle, cont);
}

bool graph_enabled()
{
    const char* e = getenv("SMK_GRAPH");
    return !(e && atoi(e) == 0);
}

} // namespace

// Runs iterations [first_iter, max_iter) of NmfSolve's loop with the stop test on the device. On return *iter / *success are what
// the host loop would have left, *rc is SMK_OK or the failure code (c->err set). Returns false if the graph could not be built
// (nothing has been executed then: the caller falls back to the host loop from first_iter).
bool nmf_loop_graph(smk_ctx* c, int first_iter, int* iter, bool* success, int* rc)
{
    const smk_nmf_options& o = c->opts;
    if (!graph_enabled() || c->nranks > 1 || c->phases_on || o.verbose || o.max_iter - first_iter < 4) return false;
    c->trace.reserve(static_cast<size_t>(o.max_iter));
    c->loop_state.reserve(8);
    LoopState init = {first_iter, 0, 0, INT_MAX};
    SMK_CUDA(cudaMemcpyAsync(c->loop_state.p, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));             // `init` lives on this stack frame; also: no capture while work is pending elsewhere

    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const int steps_before = c->steps_done;
    bool capturing = false;
    auto cleanup = [&]() {
        if (capturing) { cudaGraph_t dummy = nullptr; cudaStreamEndCapture(c->stream, &dummy); capturing = false; }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        c->steps_done = steps_before;
    };
    try
    {
        SMK_CUDA(cudaGraphCreate(&graph, 0));
        cudaGraphConditionalHandle handle;
        SMK_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams params = {};
        params.type = cudaGraphNodeTypeConditional;
        params.conditional.handle = handle;
        params.conditional.type = cudaGraphCondTypeWhile;
        params.conditional.size = 1;
        cudaGraphNode_t node;
        SMK_CUDA(cudaGraphAddNode(&node, graph, nullptr, 0, &params));
        cudaGraph_t body = params.conditional.phGraph_out[0];
        SMK_CUDA(cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
        capturing = true;
        solver_step(c);
        solver_progress_enqueue(c, c->prog.p + 2);
        nmf_loop_control_kernel<<<1, 32, 0, c->stream>>>(handle, reinterpret_cast<LoopState*>(c->loop_state.p), c->status.p, c->prog.p + 2, c->trace.p,
                                                       o.tol, o.tolcount, o.max_iter, o.algorithm == SMK_RANK2 ? 1 : 0);
        SMK_LAUNCH_CHECK();
        SMK_CUDA(cudaStreamEndCapture(c->stream, nullptr));
        capturing = false;
        SMK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    }
    catch (...)
    {
        cleanup();
        return false;
    }
    // from here on work is executed: errors are errors
    SMK_CUDA(cudaGraphLaunch(exec, c->stream));
    LoopState st;
    int status[ST_COUNT];
    SMK_CUDA(cudaMemcpyAsync(&st, c->loop_state.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaMemcpyAsync(status, c->status.p, sizeof(status), cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    for (int i = 0; i < ST_COUNT; ++i) c->status_host[i] = status[i];
    c->status_cached = true;
    *iter = st.iter;
    c->steps_done = steps_before + (st.iter - first_iter) + ((st.result != 0) ? 1 : 0);
    *success = (st.result == 1);
    *rc = SMK_OK;
    if (st.result == 2)
    {
        *iter = st.fail_iter;
        c->status_host[ST_FAIL_ITER] = st.fail_iter;
        c->err = "NMF solver failure on iteration " + std::to_string(st.fail_iter + 1);
        *rc = SMK_FAILURE;
    }
    else if (st.result == 3)
    {
        c->err = status[ST_COMM_TIMEOUT] ? "peer exchange timed out" : "ProjectedGradientNorm: NaN";
        *rc = SMK_FAILURE;
    }
    else if (st.result == 4) { c->err = "Normalize: column norm < machine epsilon"; *rc = SMK_FAILURE; }
    return true;
}

} // namespace smk
