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

void diff_sumsq(cudaStream_t stream, long long count, const double* A, const double* B, double* partial, double* acc_slot, int num_sms)
{
    int blocks = reduce_blocks(count, num_sms);
    diff_partial_kernel<<<blocks, 256, 0, stream>>>(count, A, B, partial);
    SMK_LAUNCH_CHECK();
    final_sum_kernel<<<1, 256, 0, stream>>>(blocks, partial, acc_slot);
    SMK_LAUNCH_CHECK();
}

} // namespace smk
