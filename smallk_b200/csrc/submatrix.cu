// smallk_b200 — column subsets of the loaded matrix, kept on the device (hierclust).
//
// HierNMF2 factors, at every tree node, the columns of A that belong to the node:
//   A.SubMatrixColsCompact(Asubset, subset, old_to_new_rows, new_to_old_rows)
//     sparse: common/include/sparse_matrix_impl.hpp:479-591 — copy the listed columns in list order, drop every
//             row that lost all its entries, renumber the surviving rows in ascending order;
//     dense : common/include/dense_matrix_impl.hpp:224-285 — copy the listed columns, all rows kept.
// The reference does this on the host with three sequential passes over the subset's nonzeros plus two over all
// m rows. Here A stays in HBM: lengths -> exclusive scan -> one warp per column copies its entries and flags the
// rows it touches -> exclusive scan of the flags = old_to_new -> renumber. The CSR twin of the subset (needed by
// H*A') is rebuilt by the same stable radix sort as for the full matrix, so the summation order inside every
// row stays the reference's (column-ascending).
#include <cub/cub.cuh>
#include "context.h"
#include "solver.h"

namespace smk {

namespace {

__global__ void col_lengths_kernel(int count, int n_full, const unsigned int* __restrict__ cols,
                                   const unsigned int* __restrict__ colptr, unsigned int* __restrict__ len, int* __restrict__ bad)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= count; j += gridDim.x * blockDim.x)
    {
        unsigned int l = 0;
        if (j < count)
        {
            const unsigned int c = cols[j];
            if (c >= static_cast<unsigned int>(n_full)) atomicExch(bad, 1);
            else l = colptr[c + 1] - colptr[c];
        }
        len[j] = l;
    }
}

// one warp per subset column: copy (row, value) pairs in storage order, flag the rows
__global__ void copy_cols_kernel(int count, const unsigned int* __restrict__ cols, const unsigned int* __restrict__ colptr,
                                 const unsigned int* __restrict__ rowidx, const double* __restrict__ val,
                                 const unsigned int* __restrict__ sub_colptr, unsigned int* __restrict__ sub_rowidx,
                                 double* __restrict__ sub_val, unsigned int* __restrict__ flags)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < count; j += warps)
    {
        const unsigned int c = cols[j];
        const unsigned int src = colptr[c], len = colptr[c + 1] - src, dst = sub_colptr[j];
        for (unsigned int e = lane; e < len; e += 32)
        {
            const unsigned int r = rowidx[src + e];
            sub_rowidx[dst + e] = r;
            sub_val[dst + e] = val[src + e];
            flags[r] = 1u;
        }
    }
}

__global__ void new_to_old_kernel(int m, const unsigned int* __restrict__ flags, const unsigned int* __restrict__ o2n,
                                  unsigned int* __restrict__ n2o)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x)
        if (flags[r]) n2o[o2n[r]] = r;
}

__global__ void renumber_kernel(unsigned int nnz, const unsigned int* __restrict__ o2n, unsigned int* __restrict__ rowidx)
{
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x)
        rowidx[e] = o2n[rowidx[e]];
}

__global__ void gather_dense_cols_kernel(int m, int count, const unsigned int* __restrict__ cols, const double* __restrict__ A,
                                         long long ldA, double* __restrict__ out)
{
    for (int j = blockIdx.y; j < count; j += gridDim.y)
    {
        const double* src = A + static_cast<long long>(cols[j]) * ldA;
        double* dst = out + static_cast<long long>(j) * m;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) dst[r] = src[r];
    }
}

void exclusive_scan(smk_ctx* c, const unsigned int* in, unsigned int* out, int count)
{
    size_t bytes = 0;
    SMK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, c->stream));
    c->sub_scan_tmp.reserve(bytes);
    SMK_CUDA(cub::DeviceScan::ExclusiveSum(c->sub_scan_tmp.p, bytes, in, out, count, c->stream));
    launch_counter() += 2;
}

} // namespace

void select_all(smk_ctx* c)
{
    if (!c->subset_active) return;
    c->dA = c->full_dA; c->ldA = c->full_ldA; c->m = c->full_m; c->n = c->full_n;
    c->Sa = &c->S;
    c->subset_active = false;
    c->active = false;
}

int select_columns(smk_ctx* c, const unsigned int* cols_host, int count, unsigned int* new_to_old_host)
{
    if (!c->subset_active) { c->full_dA = c->dA; c->full_ldA = c->ldA; c->full_m = c->m; c->full_n = c->n; }
    const int m = c->full_m, n_full = c->full_n;
    cudaStream_t s = c->stream;
    c->sub_cols.reserve(static_cast<size_t>(std::max(count, n_full)));
    SMK_CUDA(cudaMemcpyAsync(c->sub_cols.p, cols_host, sizeof(unsigned int) * count, cudaMemcpyHostToDevice, s));
    int* bad = c->status.p + ST_BAD_INDEX;
    SMK_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));

    if (c->has_dense)
    {
        // range check on the host: the list is already here
        for (int j = 0; j < count; ++j)
            if (cols_host[j] >= static_cast<unsigned int>(n_full)) throw std::string("SubMatrixColsCompact: column index out of range");
        c->A_sub.reserve(static_cast<size_t>(m) * std::max(count, 1));
        dim3 grid(std::max(1, std::min(ceil_div(m, 256), 64)), std::min(count, 65535));
        gather_dense_cols_kernel<<<grid, 256, 0, s>>>(m, count, c->sub_cols.p, c->full_dA, c->full_ldA, c->A_sub.p);
        SMK_LAUNCH_CHECK();
        for (int r = 0; r < m; ++r) new_to_old_host[r] = static_cast<unsigned int>(r);
        c->dA = c->A_sub.p; c->ldA = m; c->m = m; c->n = count;
        c->subset_active = true; c->active = false;
        return m;
    }

    const SparseDev& F = c->S;
    SparseDev& U = c->Ssub;
    U.colptr.reserve(static_cast<size_t>(n_full) + 2);
    c->sub_flags.reserve(static_cast<size_t>(m) + 1);
    c->sub_o2n.reserve(static_cast<size_t>(m) + 1);
    c->sub_n2o.reserve(static_cast<size_t>(m));
    // lengths (count + 1 entries, the last one 0) -> exclusive scan in place = column offsets
    col_lengths_kernel<<<std::max(1, std::min(ceil_div(count + 1, 256), 4 * c->num_sms)), 256, 0, s>>>(
        count, n_full, c->sub_cols.p, F.colptr.p, U.colptr.p, bad);
    SMK_LAUNCH_CHECK();
    exclusive_scan(c, U.colptr.p, U.colptr.p, count + 1);
    unsigned int h_nnz = 0;
    int h_bad = 0;
    SMK_CUDA(cudaMemcpyAsync(&h_nnz, U.colptr.p + count, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    SMK_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    SMK_CUDA(cudaStreamSynchronize(s));
    if (h_bad) throw std::string("SparseMatrix::SubMatrixColsCompact: column index out of range");
    if (h_nnz == 0) throw std::string("SparseMatrix::SubMatrixColsCompact: submatrix is the zero matrix");
    U.rowidx.reserve(std::max<size_t>(h_nnz, F.nnz)); U.val.reserve(std::max<size_t>(h_nnz, F.nnz));   // sized once for the whole run
    SMK_CUDA(cudaMemsetAsync(c->sub_flags.p, 0, sizeof(unsigned int) * (static_cast<size_t>(m) + 1), s));
    copy_cols_kernel<<<std::max(1, std::min(ceil_div(count, 8), 8 * c->num_sms)), 256, 0, s>>>(
        count, c->sub_cols.p, F.colptr.p, F.rowidx.p, F.val.p, U.colptr.p, U.rowidx.p, U.val.p, c->sub_flags.p);
    SMK_LAUNCH_CHECK();
    exclusive_scan(c, c->sub_flags.p, c->sub_o2n.p, m + 1);
    const int eb = std::max(1, std::min(ceil_div(m, 256), 8 * c->num_sms));
    new_to_old_kernel<<<eb, 256, 0, s>>>(m, c->sub_flags.p, c->sub_o2n.p, c->sub_n2o.p);
    SMK_LAUNCH_CHECK();
    renumber_kernel<<<std::max(1u, std::min<unsigned int>((h_nnz + 255) / 256, 8u * c->num_sms)), 256, 0, s>>>(h_nnz, c->sub_o2n.p, U.rowidx.p);
    SMK_LAUNCH_CHECK();
    unsigned int h_new_m = 0;
    SMK_CUDA(cudaMemcpyAsync(&h_new_m, c->sub_o2n.p + m, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    SMK_CUDA(cudaStreamSynchronize(s));
    SMK_CUDA(cudaMemcpyAsync(new_to_old_host, c->sub_n2o.p, sizeof(unsigned int) * h_new_m, cudaMemcpyDeviceToHost, s));
    U.m = static_cast<int>(h_new_m); U.n = count; U.nnz = h_nnz;
    // scratch of the CSR build sized for the full matrix once, so no reallocation happens down the tree
    U.t_colof.reserve(F.nnz); U.t_ids.reserve(F.nnz); U.t_ids_sorted.reserve(F.nnz); U.t_rows_sorted.reserve(F.nnz);
    U.colidx.reserve(F.nnz); U.valr.reserve(F.nnz); U.rowptr.reserve(static_cast<size_t>(m) + 1);
    build_csr(s, U, /*keep_scratch=*/true);
    build_segments(s, U.n, U.colptr.p, U.seg_cols, c->num_sms);
    build_segments(s, U.m, U.rowptr.p, U.seg_rows, c->num_sms);
    c->Sa = &U; c->m = U.m; c->n = count;
    c->subset_active = true; c->active = false;
    return U.m;
}

} // namespace smk
