// smallk_b200 — bandwidth-bound helpers: transposes, MU update, progress-metric reductions.
#include "common.cuh"
#include "kernels.h"

namespace smk {

namespace {

// 32 x 32 tile transpose through shared memory (padded: no bank conflicts).
__global__ void transpose_kernel(int rows, int cols, const double* __restrict__ in, long long ldi,
                                 double* __restrict__ out, long long ldo)
{
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
    {
        const int r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[static_cast<long long>(c) * ldi + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
    {
        const int c = c0 + threadIdx.x, r = r0 + j;
        if (r < rows && c < cols) out[static_cast<long long>(r) * ldo + c] = tile[threadIdx.x][j];
    }
}

__global__ void mu_update_kernel(long long count, double* __restrict__ X, const double* __restrict__ Num,
                                 const double* __restrict__ Den)
{
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        X[i] *= (Num[i] / (Den[i] + 1.0e-13));
}

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (warp == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

__global__ void pg_partial_kernel(long long count, const double* __restrict__ G, const double* __restrict__ X,
                                  double* __restrict__ partial)
{
    double s = 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const double g = G[i];
        if (g < 0.0 || X[i] > 0.0) s += g * g;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void diff_partial_kernel(long long count, const double* __restrict__ A, const double* __restrict__ B,
                                    double* __restrict__ partial)
{
    double s = 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const double d = B ? (A[i] - B[i]) : A[i];
        s += d * d;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void final_sum_kernel(int n, const double* __restrict__ partial, double* __restrict__ out)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    s = block_sum(s);
    if (threadIdx.x == 0) *out = s;
}

// Both projected-gradient sums of one outer iteration (projected_gradient.hpp:125-171: the W part and the H part) in ONE
// launch: block partials in a fixed grid-stride order, the block that arrives last adds them in block order (deterministic)
// and, when `prog` is given, finishes ProgressEst::Update on the device (progress_metric below).
__device__ __forceinline__ void progress_metric(int mode, const double* acc, double* prog, double* metric_out, int* status)
{
    if (mode == 0)
    {
        // ProgEstGenericPgRatio, progress_estimator_generic.hpp:87-104: the first evaluation stores pg0 and reports 1
        const double pg = sqrt(acc[0] + acc[1]);
        if (pg != pg) atomicExch(&status[ST_PG_NAN], 1);
        if (prog[1] == 0.0) { prog[0] = pg; prog[1] = 1.0; *metric_out = 1.0; }
        else *metric_out = pg / prog[0];
    }
    else *metric_out = sqrt(acc[0]) / sqrt(acc[1]);     // ProgEstGenericDeltaW, :58-69
}

__global__ void pg_pair_kernel(long long c1, const double* __restrict__ G1, const double* __restrict__ X1,
                               long long c2, const double* __restrict__ G2, const double* __restrict__ X2,
                               double* __restrict__ partial, unsigned int* __restrict__ ticket, double* __restrict__ acc,
                               double* prog, double* metric_out, int* status)
{
    __shared__ bool s_last;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    // 16-byte loads, two independent pairs in flight per thread and trip (the four arrays are cudaMalloc'd: 256-byte aligned)
    auto part = [&](const double* __restrict__ G, const double* __restrict__ X, const long long cnt) {
        double s = 0.0;
        const long long pairs = cnt >> 1;
        const double2* G2v = reinterpret_cast<const double2*>(G);
        const double2* X2v = reinterpret_cast<const double2*>(X);
        const bool vec = ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(X)) & 15u) == 0;
        if (vec)
        {
            long long i = i0;
            for (; i + stride < pairs; i += 2 * stride)
            {
                const double2 ga = G2v[i], xa = X2v[i], gb = G2v[i + stride], xb = X2v[i + stride];
                if (ga.x < 0.0 || xa.x > 0.0) s += ga.x * ga.x;
                if (ga.y < 0.0 || xa.y > 0.0) s += ga.y * ga.y;
                if (gb.x < 0.0 || xb.x > 0.0) s += gb.x * gb.x;
                if (gb.y < 0.0 || xb.y > 0.0) s += gb.y * gb.y;
            }
            for (; i < pairs; i += stride)
            {
                const double2 ga = G2v[i], xa = X2v[i];
                if (ga.x < 0.0 || xa.x > 0.0) s += ga.x * ga.x;
                if (ga.y < 0.0 || xa.y > 0.0) s += ga.y * ga.y;
            }
            if ((cnt & 1) && i0 == 0) { const double g = G[cnt - 1]; if (g < 0.0 || X[cnt - 1] > 0.0) s += g * g; }
        }
        else
            for (long long i = i0; i < cnt; i += stride) { const double g = G[i]; if (g < 0.0 || X[i] > 0.0) s += g * g; }
        return s;
    };
    double s1 = part(G1, X1, c1), s2 = part(G2, X2, c2);
    s1 = block_sum(s1);
    s2 = block_sum(s2);
    if (threadIdx.x == 0)
    {
        partial[blockIdx.x] = s1;
        partial[1024 + blockIdx.x] = s2;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    s1 = 0.0; s2 = 0.0;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x) { s1 += __ldcg(partial + i); s2 += __ldcg(partial + 1024 + i); }
    s1 = block_sum(s1);
    s2 = block_sum(s2);
    if (threadIdx.x == 0)
    {
        acc[0] = s1; acc[1] = s2;
        *ticket = 0u;
        if (prog) progress_metric(0, acc, prog, metric_out, status);
    }
}

__global__ void progress_metric_kernel(int mode, const double* acc, double* prog, double* metric_out, int* status)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) progress_metric(mode, acc, prog, metric_out, status);
}

int reduce_blocks(long long count, int num_sms)
{
    long long b = (count + 255) / 256;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(b, std::min(1024, 4 * num_sms))));
}

} // namespace

void transpose_f64(cudaStream_t stream, int rows, int cols, const double* in, long long ldi, double* out, long long ldo)
{
    if (rows <= 0 || cols <= 0) return;
    dim3 grid(ceil_div(rows, 32), ceil_div(cols, 32));
    dim3 block(32, 8);
    transpose_kernel<<<grid, block, 0, stream>>>(rows, cols, in, ldi, out, ldo);
    SMK_LAUNCH_CHECK();
}

void mu_update(cudaStream_t stream, long long count, double* X, const double* Num, const double* Den)
{
    if (count <= 0) return;
    int blocks = static_cast<int>(std::min<long long>((count + 255) / 256, 4096));
    mu_update_kernel<<<blocks, 256, 0, stream>>>(count, X, Num, Den);
    SMK_LAUNCH_CHECK();
}

void pg_sumsq(cudaStream_t stream, long long count, const double* G, const double* X, double* partial, double* acc_slot, int num_sms)
{
    int blocks = reduce_blocks(count, num_sms);
    pg_partial_kernel<<<blocks, 256, 0, stream>>>(count, G, X, partial);
    SMK_LAUNCH_CHECK();
    final_sum_kernel<<<1, 256, 0, stream>>>(blocks, partial, acc_slot);
    SMK_LAUNCH_CHECK();
}

void pg_pair(cudaStream_t stream, long long c1, const double* G1, const double* X1, long long c2, const double* G2, const double* X2,
             double* partial, unsigned int* ticket, double* acc, double* prog, double* metric_out, int* status, int num_sms)
{
    const int blocks = reduce_blocks(std::max(c1, c2), num_sms);
    pg_pair_kernel<<<blocks, 256, 0, stream>>>(c1, G1, X1, c2, G2, X2, partial, ticket, acc, prog, metric_out, status);
    SMK_LAUNCH_CHECK();
}

void progress_metric_launch(cudaStream_t stream, int mode, const double* acc, double* prog, double* metric_out, int* status)
{
    progress_metric_kernel<<<1, 32, 0, stream>>>(mode, acc, prog, metric_out, status);
    SMK_LAUNCH_CHECK();
}

void diff_sumsq(cudaStream_t stream, long long count, const double* A, const double* B, double* partial, double* acc_slot, int num_sms)
{
    int blocks = reduce_blocks(count, num_sms);
    diff_partial_kernel<<<blocks, 256, 0, stream>>>(count, A, B, partial);
    SMK_LAUNCH_CHECK();
    final_sum_kernel<<<1, 256, 0, stream>>>(blocks, partial, acc_slot);
    SMK_LAUNCH_CHECK();
}

} // namespace smk
