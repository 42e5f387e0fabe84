// smallk_b200 — the exchange steps of the column-sharded NMF iteration as OUR OWN kernels over NVLink peer memory.
//
// Multi-GPU layout: A and H are split by column block, one process per GPU; per outer iteration
// the ranks exchange (1) k x k Gram matrices and a handful of scalars / flags, (2) the k x m product H*A' (summed over
// the ranks, each rank keeping a row block of it) and (3) the updated row blocks of W. NCCL does these as seven separate
// collectives of 4 bytes ... 10 MB per iteration, each paying its own launch + protocol latency (r01: 0.8 ms of a
// 1.38 ms step at 8 GPUs). Here every rank maps every other rank's exchange region (cudaIpc handles, NVSwitch gives
// all-to-all load/store) and the exchanges are plain kernels on the solver's stream:
//
//   peer_allreduce_kernel   one-shot all-reduce of <= 64K doubles (+ an OR flag and a failure flag riding along): every
//                           rank stores its vector into its slot on every peer, signals, waits for the others' signals and
//                           adds the slots in RANK ORDER — the result is bitwise identical on all ranks and independent of
//                           timing; optionally finishes ProgressEst::Update (the PG ratio) in the same launch.
//   peer_scatter_kernel     reduce-scatter, producer half: takes the split-R partial tiles of the H*A' GEMM (or the SpMM
//                           output), adds them in split order and stores each row block straight into its OWNER's receive
//                           slot — the GEMM's own reduction pass doubles as the NVLink transfer, no staging copy.
//   peer_gather_sum_kernel  reduce-scatter, consumer half: waits for all producers, adds the receive slots in rank order.
//   peer_allgather_kernel   every rank stores its block of a k x m buffer into the same place on every peer, signals, and
//                           returns when all blocks have arrived.
//
// Synchronisation: per exchange class one 64-bit flag per (receiver, sender), carrying a monotonically increasing epoch;
// data stores -> __threadfence_system() -> (last CTA) st.release.sys of the epoch; receivers spin with ld.acquire.sys.
// Every spin has a wall-clock bound (60 s): a rank that dies cannot hang the others, they raise ST_COMM_TIMEOUT instead.
// NCCL is still used to bootstrap (exchange of the IPC handles) and remains selectable (SMK_PEER=0) for A/B measurements.
#include <cstdlib>
#include <cstring>
#include "context.h"
#include "peer.h"
#include "peer_device.cuh"

namespace smk {

namespace {

// ---------------------------------------------------------------------------
// one-shot all-reduce (sum) of `count` doubles, in place in `data`, + optional flags:
//   or_flag   : int, becomes 1 on every rank if it was non-zero on any rank (BPP's "some column was non-optimal")
//   fail_iter : int, first outer iteration with a solver failure (INT_MAX = none): becomes the minimum over the ranks
// slots: kPeerSmallCap doubles per (parity, sender) in every rank's region.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
peer_allreduce_kernel(PeerTable t, int rank, int nranks, unsigned long long epoch, double* __restrict__ data, int count,
                      int* or_flag, int* fail_iter, unsigned int* ticket, int* status,
                      int metric_mode, double* prog, double* metric_out)
{
    __shared__ bool s_last;
    __shared__ int s_ok;
    const size_t slot_base = kPeerSmallOffset + (static_cast<size_t>(epoch & 1ull) * kPeerMaxRanks) * kPeerSmallCap * sizeof(double);
    const int total = count + 2;              // [count] = or flag, [count + 1] = failure flag
    const int gsz = gridDim.x * blockDim.x, gid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = gid; i < total; i += gsz)
    {
        double v;
        if (i < count) v = data[i];
        else if (i == count) v = (or_flag && *or_flag != 0) ? 1.0 : 0.0;
        else v = fail_iter ? static_cast<double>(*fail_iter) : static_cast<double>(INT_MAX);
        for (int p = 0; p < nranks; ++p)
        {
            const int peer = (rank + p) % nranks;         // spread the traffic over the links
            reinterpret_cast<double*>(t.base[peer] + slot_base)[static_cast<size_t>(rank) * kPeerSmallCap + i] = v;
        }
    }
    if (stores_done_last_cta(ticket, &s_last))
    {
        if (threadIdx.x < nranks) st_release_sys(flag_ptr(t, threadIdx.x, kFlagSmall, rank), epoch);
    }
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < nranks && !wait_flag(flag_ptr(t, rank, kFlagSmall, threadIdx.x), epoch)) s_ok = 0;
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) atomicExch(&status[ST_COMM_TIMEOUT], 1); return; }
    const double* slots = reinterpret_cast<const double*>(t.base[rank] + slot_base);
    for (int i = gid; i < total; i += gsz)
    {
        double s = __ldcg(slots + i);
        if (i <= count) { for (int r = 1; r < nranks; ++r) s += __ldcg(slots + static_cast<size_t>(r) * kPeerSmallCap + i); }
        else { for (int r = 1; r < nranks; ++r) s = fmin(s, __ldcg(slots + static_cast<size_t>(r) * kPeerSmallCap + i)); }
        if (i < count) data[i] = s;
        else if (i == count) { if (or_flag) *or_flag = (s > 0.0) ? 1 : 0; }
        else if (fail_iter) *fail_iter = static_cast<int>(s);
    }
    // ProgressEst::Update on the reduced sums (single-CTA launches only: data[0..1] were written by this CTA)
    if (prog && gridDim.x == 1)
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            if (metric_mode == 0)
            {
                const double pg = sqrt(data[0] + data[1]);
                if (pg != pg) atomicExch(&status[ST_PG_NAN], 1);
                if (prog[1] == 0.0) { prog[0] = pg; prog[1] = 1.0; *metric_out = 1.0; }
                else *metric_out = pg / prog[0];
            }
            else *metric_out = sqrt(data[0]) / sqrt(data[1]);
        }
    }
}

// ---------------------------------------------------------------------------
// reduce-scatter, producer: out element e of the k x m_pad matrix = sum over the splits (ascending) of partial[s][e]
// (entries beyond k*m do not exist: zero); block g (piece doubles) goes to rank g's receive slot [rank].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
peer_scatter_kernel(PeerTable t, int rank, int nranks, unsigned long long epoch, const double* __restrict__ partial, int splits,
                    long long valid, long long piece, size_t recv_off, unsigned int* ticket)
{
    __shared__ bool s_last;
    const long long total = piece * nranks;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if (((piece | valid) & 1) == 0)
    {
        // two doubles per thread and step: 16-byte peer stores (a pair never straddles two blocks when piece is even)
        for (long long e = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 2; e < total; e += stride * 2)
        {
            double2 v = make_double2(0.0, 0.0);
            if (e < valid)
            {
                v = *reinterpret_cast<const double2*>(partial + e);
                for (int z = 1; z < splits; ++z)
                {
                    const double2 b = *reinterpret_cast<const double2*>(partial + static_cast<long long>(z) * valid + e);
                    v.x += b.x; v.y += b.y;
                }
            }
            const int g = static_cast<int>(e / piece);
            *reinterpret_cast<double2*>(reinterpret_cast<double*>(t.base[g] + recv_off) + static_cast<long long>(rank) * piece + (e - g * piece)) = v;
        }
    }
    else
    {
        for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total; e += stride)
        {
            double v = 0.0;
            if (e < valid)
            {
                v = partial[e];
                for (int z = 1; z < splits; ++z) v += partial[static_cast<long long>(z) * valid + e];
            }
            const int g = static_cast<int>(e / piece);
            (reinterpret_cast<double*>(t.base[g] + recv_off) + static_cast<long long>(rank) * piece)[e - g * piece] = v;
        }
    }
    if (stores_done_last_cta(ticket, &s_last))
    {
        if (threadIdx.x < nranks) st_release_sys(flag_ptr(t, threadIdx.x, kFlagScatter, rank), epoch);
    }
}

// reduce-scatter, consumer: out[0..piece) = sum over ranks r (ascending) of recv[r][0..piece)
__global__ void __launch_bounds__(256)
peer_gather_sum_kernel(PeerTable t, int rank, int nranks, unsigned long long epoch, long long piece, size_t recv_off,
                       double* __restrict__ out, int* status)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < nranks && !wait_flag(flag_ptr(t, rank, kFlagScatter, threadIdx.x), epoch)) s_ok = 0;
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) atomicExch(&status[ST_COMM_TIMEOUT], 1); return; }
    const double* recv = reinterpret_cast<const double*>(t.base[rank] + recv_off);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < piece; e += stride)
    {
        double s = __ldcg(recv + e);
        for (int r = 1; r < nranks; ++r) s += __ldcg(recv + static_cast<long long>(r) * piece + e);
        out[e] = s;
    }
}

// ---------------------------------------------------------------------------
// all-gather of a k x m_pad buffer that lives at the same offset in every rank's region: my block -> every peer
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
peer_allgather_kernel(PeerTable t, int rank, int nranks, unsigned long long epoch, size_t buf_off, long long piece,
                      unsigned int* ticket, int* status)
{
    __shared__ bool s_last;
    __shared__ int s_ok;
    const double* src = reinterpret_cast<const double*>(t.base[rank] + buf_off) + static_cast<long long>(rank) * piece;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if ((piece & 1) == 0)
    {
        for (long long e = i0 * 2; e < piece; e += stride * 2)
        {
            const double2 v = *reinterpret_cast<const double2*>(src + e);
            for (int p = 1; p < nranks; ++p)
            {
                const int peer = (rank + p) % nranks;
                *reinterpret_cast<double2*>(reinterpret_cast<double*>(t.base[peer] + buf_off) + static_cast<long long>(rank) * piece + e) = v;
            }
        }
    }
    else
    {
        for (long long e = i0; e < piece; e += stride)
        {
            const double v = src[e];
            for (int p = 1; p < nranks; ++p)
            {
                const int peer = (rank + p) % nranks;
                (reinterpret_cast<double*>(t.base[peer] + buf_off) + static_cast<long long>(rank) * piece)[e] = v;
            }
        }
    }
    if (stores_done_last_cta(ticket, &s_last))
    {
        if (threadIdx.x < nranks) st_release_sys(flag_ptr(t, threadIdx.x, kFlagGather, rank), epoch);
    }
    // every CTA waits: when the kernel has completed, all blocks of all ranks are in place
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < nranks && !wait_flag(flag_ptr(t, rank, kFlagGather, threadIdx.x), epoch)) s_ok = 0;
    __syncthreads();
    if (!s_ok && threadIdx.x == 0) atomicExch(&status[ST_COMM_TIMEOUT], 1);
}

void nccl_ok(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess) throw std::string(what) + ": " + ncclGetErrorString(r);
}

} // namespace

bool peer_enabled_by_env()
{
    const char* e = getenv("SMK_PEER");
    return !(e && atoi(e) == 0);
}

// SMK_PEER_FUSED=0 keeps the GEMM and the reduce-scatter producer as two kernels (A/B measurements)
bool peer_fused_by_env()
{
    const char* e = getenv("SMK_PEER_FUSED");
    return !(e && atoi(e) == 0);
}

void peer_release(smk_ctx* c)
{
    PeerComm& P = c->peer;
    if (!P.local) return;
    cudaStreamSynchronize(c->stream);
    for (int r = 0; r < P.nranks; ++r)
        if (r != P.rank && P.table.base[r]) cudaIpcCloseMemHandle(P.table.base[r]);
    // every rank has unmapped its peers before anybody frees (a barrier through NCCL; harmless when the comm is gone)
    if (c->comm && P.nranks > 1)
    {
        if (ncclAllReduce(c->acc.p, c->acc.p, 1, ncclDouble, ncclSum, c->comm, c->stream) == ncclSuccess)
            cudaStreamSynchronize(c->stream);
    }
    cudaFree(P.local);
    P = PeerComm();
}

// (Re)creates the exchange region with room for `big_doubles` doubles per k x m_pad buffer (three of them: Wt, HAt and the
// receive slots) and maps every peer's region. Collective: all ranks call it with the same arguments.
void peer_setup(smk_ctx* c, size_t big_doubles)
{
    PeerComm& P = c->peer;
    if (c->nranks > kPeerMaxRanks) throw std::string("peer exchange supports at most 8 ranks (one node)");
    const size_t big_bytes = (big_doubles * sizeof(double) + 255) & ~static_cast<size_t>(255);
    if (P.local && P.big_bytes >= big_bytes && P.nranks == c->nranks) return;
    peer_release(c);
    P.rank = c->rank; P.nranks = c->nranks;
    P.big_bytes = big_bytes;
    P.bytes = kPeerBigOffset + 3 * big_bytes;
    SMK_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.local), P.bytes));
    SMK_CUDA(cudaMemsetAsync(P.local, 0, P.bytes, c->stream));
    // exchange the IPC handles through NCCL (bootstrap only)
    cudaIpcMemHandle_t mine;
    SMK_CUDA(cudaIpcGetMemHandle(&mine, P.local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    DevBuf<unsigned char> hbuf;
    hbuf.reserve(static_cast<size_t>(64) * (c->nranks + 1));
    SMK_CUDA(cudaMemcpyAsync(hbuf.p, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    nccl_ok(ncclAllGather(hbuf.p, hbuf.p + 64, 64, ncclChar, c->comm, c->stream), "ncclAllGather (IPC handles)");
    std::vector<cudaIpcMemHandle_t> all(c->nranks);
    SMK_CUDA(cudaMemcpyAsync(all.data(), hbuf.p + 64, static_cast<size_t>(64) * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < c->nranks; ++r)
    {
        if (r == c->rank) { P.table.base[r] = P.local; continue; }
        void* ptr = nullptr;
        SMK_CUDA(cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess));
        P.table.base[r] = static_cast<unsigned char*>(ptr);
    }
    // second barrier: nobody starts storing into a peer before that peer's memset has completed (it was enqueued before
    // the first all-gather on the peer's stream, and this all-reduce completes only after every rank has reached it)
    nccl_ok(ncclAllReduce(c->acc.p, c->acc.p, 1, ncclDouble, ncclSum, c->comm, c->stream), "ncclAllReduce (barrier)");
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    P.epoch[0] = P.epoch[1] = P.epoch[2] = 0;
    c->peer_ticket.reserve(4);
    SMK_CUDA(cudaMemsetAsync(c->peer_ticket.p, 0, 4 * sizeof(unsigned int), c->stream));
}

double* peer_big_buffer(smk_ctx* c, int which)
{
    return reinterpret_cast<double*>(c->peer.local + kPeerBigOffset + static_cast<size_t>(which) * c->peer.big_bytes);
}

void peer_allreduce(smk_ctx* c, double* data, int count, int* or_flag, int* fail_iter, int metric_mode, double* prog, double* metric_out,
                    cudaStream_t stream)
{
    if (!stream) stream = c->stream;
    PeerComm& P = c->peer;
    if (count + 2 > kPeerSmallCap) throw std::string("peer_allreduce: vector too long");
    const unsigned long long epoch = ++P.epoch[kFlagSmall];
    const int grid = (prog || count <= 2048) ? 1 : std::min(32, ceil_div(count, 2048));
    peer_allreduce_kernel<<<grid, 512, 0, stream>>>(P.table, P.rank, P.nranks, epoch, data, count, or_flag, fail_iter,
                                                     c->peer_ticket.p + 0, c->status.p, metric_mode, prog, metric_out);
    SMK_LAUNCH_CHECK();
}

// partial: [splits][valid] doubles on this rank (valid = k * m existing entries); the sum's row block g lands in rank g's
// receive slots, and out (this rank's k x m_loc block) is their sum in rank order.
void peer_reduce_scatter(smk_ctx* c, const double* partial, int splits, long long valid, long long piece, double* out)
{
    PeerComm& P = c->peer;
    const unsigned long long epoch = ++P.epoch[kFlagScatter];
    const size_t recv_off = kPeerBigOffset + 2 * P.big_bytes;
    const long long total = piece * P.nranks;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((total / 2 + 255) / 256, 4LL * c->num_sms)));
    peer_scatter_kernel<<<grid, 256, 0, c->stream>>>(P.table, P.rank, P.nranks, epoch, partial, splits, valid, piece, recv_off,
                                                   c->peer_ticket.p + 1);
    SMK_LAUNCH_CHECK();
    const int grid2 = static_cast<int>(std::max<long long>(1, std::min<long long>((piece + 255) / 256, 2LL * c->num_sms)));
    peer_gather_sum_kernel<<<grid2, 256, 0, c->stream>>>(P.table, P.rank, P.nranks, epoch, piece, recv_off, out, c->status.p);
    SMK_LAUNCH_CHECK();
}

// Fused variant: the GEMM's own epilogue is the producer half (GemmScatter); this returns the descriptor for the next
// launch and peer_reduce_scatter_finish() runs the consumer half.
GemmScatter peer_scatter_begin(smk_ctx* c, int rows_k, int cols_per_rank)
{
    PeerComm& P = c->peer;
    GemmScatter sc;
    sc.table = P.table;
    sc.recv_off = kPeerBigOffset + 2 * P.big_bytes;
    sc.piece = static_cast<long long>(rows_k) * cols_per_rank;
    sc.cols_per_rank = cols_per_rank;
    sc.rank = P.rank; sc.nranks = P.nranks;
    sc.epoch = ++P.epoch[kFlagScatter];
    sc.done = c->peer_ticket.p + 3;
    return sc;
}

void peer_reduce_scatter_finish(smk_ctx* c, long long piece, double* out)
{
    PeerComm& P = c->peer;
    const size_t recv_off = kPeerBigOffset + 2 * P.big_bytes;
    const int grid2 = static_cast<int>(std::max<long long>(1, std::min<long long>((piece + 255) / 256, 2LL * c->num_sms)));
    peer_gather_sum_kernel<<<grid2, 256, 0, c->stream>>>(P.table, P.rank, P.nranks, P.epoch[kFlagScatter], piece, recv_off, out, c->status.p);
    SMK_LAUNCH_CHECK();
}

void peer_allgather(smk_ctx* c, int which_buffer, long long piece)
{
    PeerComm& P = c->peer;
    const unsigned long long epoch = ++P.epoch[kFlagGather];
    const size_t off = kPeerBigOffset + static_cast<size_t>(which_buffer) * P.big_bytes;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((piece / 2 + 255) / 256, static_cast<long long>(c->num_sms))));
    peer_allgather_kernel<<<grid, 256, 0, c->stream>>>(P.table, P.rank, P.nranks, epoch, off, piece, c->peer_ticket.p + 2, c->status.p);
    SMK_LAUNCH_CHECK();
}

} // namespace smk
