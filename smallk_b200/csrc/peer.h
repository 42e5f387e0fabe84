// smallk_b200 — NVLink peer-memory exchange region of one rank (peer.cu). Internal header.
#pragma once

#include <cstddef>
#include <cuda_runtime.h>

struct smk_ctx;

namespace smk {

constexpr int kPeerMaxRanks = 8;                      // one NVSwitch node
constexpr int kPeerSmallCap = 65536 + 64;             // doubles per (parity, sender) slot of the one-shot all-reduce: k*k for k <= 256
enum { kFlagSmall = 0, kFlagScatter = 1, kFlagGather = 2, kFlagClasses = 3 };

// Region layout (identical on every rank, so one offset addresses the same object everywhere):
//   [kPeerFlagOffset)   flags[class][sender]            64-bit epochs, written by the sender, polled by the owner
//   [kPeerSmallOffset)  small[parity][sender][cap]      all-reduce slots, double-buffered by epoch parity
//   [kPeerBigOffset)    three k x m_pad buffers: 0 = Wt, 1 = HAt, 2 = receive slots of the reduce-scatter [sender][k x m_loc]
constexpr size_t kPeerFlagOffset = 0;
constexpr size_t kPeerSmallOffset = 4096;
constexpr size_t kPeerBigOffset = kPeerSmallOffset + static_cast<size_t>(2) * kPeerMaxRanks * kPeerSmallCap * sizeof(double);
static_assert(kPeerBigOffset % 256 == 0, "big buffers must stay 256-byte aligned");

struct PeerTable { unsigned char* base[kPeerMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; };

struct PeerComm
{
    int rank = 0, nranks = 1;
    unsigned char* local = nullptr;
    size_t bytes = 0, big_bytes = 0;
    PeerTable table;
    unsigned long long epoch[kFlagClasses] = {0, 0, 0};
};

bool peer_enabled_by_env();
void peer_setup(smk_ctx* c, size_t big_doubles);
void peer_release(smk_ctx* c);
double* peer_big_buffer(smk_ctx* c, int which);
// in-place sum over the ranks of data[0..count) (+ OR of *or_flag, + "a failure anywhere fails everywhere" for *fail_iter);
// with prog != null the same launch finishes ProgressEst::Update on data[0..1] (metric_mode as progress_metric_launch)
// (stream: 0 = the context's main stream)
void peer_allreduce(smk_ctx* c, double* data, int count, int* or_flag, int* fail_iter, int metric_mode, double* prog, double* metric_out,
                    cudaStream_t stream = nullptr);
void peer_reduce_scatter(smk_ctx* c, const double* partial, int splits, long long valid, long long piece, double* out);
void peer_allgather(smk_ctx* c, int which_buffer, long long piece);
struct GemmScatter;
GemmScatter peer_scatter_begin(smk_ctx* c, int rows_k, int cols_per_rank);
void peer_reduce_scatter_finish(smk_ctx* c, long long piece, double* out);
bool peer_fused_by_env();

} // namespace smk
