// smallk_b200 — device sorts behind smk_argsort_desc / smk_sort_desc: the orderings the hierclust priority score needs
// (hierclust/include/clust_hier_util.hpp:30-57,86). The reference calls std::sort on m-entry index vectors three to
// nine times per tree node; here each is one stable radix sort (cub) of at most m keys.
#include <cub/cub.cuh>
#include "context.h"
#include "solver.h"

namespace smk {

namespace {
__global__ void prep_keys_kernel(int n, double* __restrict__ keys, int* __restrict__ vals)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        keys[i] = keys[i] + 0.0;      // -0.0 -> +0.0: the comparator of the reference does not distinguish them
        if (vals) vals[i] = i;
    }
}
} // namespace

void device_sort_desc(smk_ctx* c, const double* host_in, int n, int* order_host, double* sorted_host)
{
    cudaStream_t s = c->stream;
    c->sort_keys.reserve(n); c->sort_keys_out.reserve(n);
    SMK_CUDA(cudaMemcpyAsync(c->sort_keys.p, host_in, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    const int blocks = std::max(1, std::min(ceil_div(n, 256), 4 * c->num_sms));
    size_t bytes = 0;
    if (order_host)
    {
        c->sort_vals.reserve(n); c->sort_vals_out.reserve(n);
        prep_keys_kernel<<<blocks, 256, 0, s>>>(n, c->sort_keys.p, c->sort_vals.p);
        SMK_LAUNCH_CHECK();
        SMK_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, c->sort_keys.p, c->sort_keys_out.p, c->sort_vals.p,
                                                           c->sort_vals_out.p, n, 0, 64, s));
        c->sort_tmp.reserve(bytes);
        SMK_CUDA(cub::DeviceRadixSort::SortPairsDescending(c->sort_tmp.p, bytes, c->sort_keys.p, c->sort_keys_out.p, c->sort_vals.p,
                                                           c->sort_vals_out.p, n, 0, 64, s));
        SMK_CUDA(cudaMemcpyAsync(order_host, c->sort_vals_out.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    }
    else
    {
        prep_keys_kernel<<<blocks, 256, 0, s>>>(n, c->sort_keys.p, nullptr);
        SMK_LAUNCH_CHECK();
        SMK_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, bytes, c->sort_keys.p, c->sort_keys_out.p, n, 0, 64, s));
        c->sort_tmp.reserve(bytes);
        SMK_CUDA(cub::DeviceRadixSort::SortKeysDescending(c->sort_tmp.p, bytes, c->sort_keys.p, c->sort_keys_out.p, n, 0, 64, s));
        SMK_CUDA(cudaMemcpyAsync(sorted_host, c->sort_keys_out.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    }
    launch_counter() += 8;
    SMK_CUDA(cudaStreamSynchronize(s));
}

} // namespace smk
