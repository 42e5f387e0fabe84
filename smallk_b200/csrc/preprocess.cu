// smallk_b200 — the tf-idf preprocessing of a term-frequency matrix on the device.
//
// Replaces preprocess_tf (preprocessor/src/preprocess.cpp:81-250), the upstream producer of the sparse matrices the NMF path
// factors: rows sorted inside every column (SortRows, :120), then rounds of
//     PruneRows     :278-369   a term stays iff its total count >= docs_per_term AND it does not occur in every document
//     PrunableCols  :372-401   a document stays iff it has >= terms_per_doc distinct terms
//     UniqueCols    :665-760   of every group of identical columns the one with the LARGEST index stays (:561-565, :641-655)
//     PruneCols     :404-445   compaction, order kept
// until a round removes nothing (or max_iter rounds), then the scores (1 + ln count) * ln(width / document frequency) with every
// column scaled to unit 2-norm (:193-230). All of it is HBM-bound integer work over (row, count) pairs: histograms by integer
// atomics (exact, order-free), stream compaction by flag -> exclusive scan -> scatter (order kept), duplicate detection by a
// 64-bit content hash per column, a stable sort of the hashes and an EXACT comparison inside every run of equal hashes (the
// reference hashes with SpookyHash and compares exactly too: which hash finds the candidates is an implementation detail of
// "identical"). Index outputs are bit-identical to the reference's; the scores differ from it at rounding level only (the
// device's log() and a warp-ordered sum of squares instead of glibc's log and a sequential sum).
// CUB supplies the scans and the two sorts; the kernels below are the passes over the matrix.
#include <cub/cub.cuh>
#include <vector>

#include "context.h"

namespace smk {

namespace {

constexpr int kT = 256;

inline int blocks_for(long long n, int cap = 8 * 148 * 4) { return static_cast<int>(std::max<long long>(1, std::min<long long>((n + kT - 1) / kT, cap))); }

__global__ void pp_counts_to_u32(unsigned int nnz, const double* __restrict__ in, unsigned int* __restrict__ out)
{
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) out[e] = static_cast<unsigned int>(in[e]);
}

// hist[r] = sum of the counts of row r, hist_nz[r] = number of entries of row r (integer atomics: exact)
__global__ void pp_row_hist(unsigned int nnz, const unsigned int* __restrict__ rows, const unsigned int* __restrict__ counts,
                            unsigned int* __restrict__ hist, unsigned int* __restrict__ hist_nz)
{
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x)
    {
        const unsigned int r = rows[e];
        atomicAdd(hist + r, counts[e]);
        atomicAdd(hist_nz + r, 1u);
    }
}

__global__ void pp_row_keep(unsigned int height, const unsigned int* __restrict__ hist, const unsigned int* __restrict__ hist_nz,
                            unsigned int docs_per_term, unsigned int width, unsigned int* __restrict__ keep)
{
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < height; r += gridDim.x * blockDim.x)
        keep[r] = (hist[r] >= docs_per_term && hist_nz[r] < width) ? 1u : 0u;
}

__global__ void pp_entry_flag(unsigned int nnz, const unsigned int* __restrict__ rows, const unsigned int* __restrict__ keep_r, unsigned int* __restrict__ flag)
{
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) flag[e] = keep_r[rows[e]];
}

// entries that survive the row pruning move to their scanned position with the renumbered row; column starts follow the scan
__global__ void pp_compact_entries(unsigned int nnz, const unsigned int* __restrict__ rows, const unsigned int* __restrict__ counts,
                                   const unsigned int* __restrict__ flag, const unsigned int* __restrict__ pos, const unsigned int* __restrict__ renum,
                                   unsigned int* __restrict__ rows2, unsigned int* __restrict__ counts2)
{
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x)
        if (flag[e]) { rows2[pos[e]] = renum[rows[e]]; counts2[pos[e]] = counts[e]; }
}

__global__ void pp_colptr_after_rows(unsigned int width, unsigned int nnz, unsigned int new_nnz, const unsigned int* __restrict__ colptr,
                                     const unsigned int* __restrict__ pos, unsigned int* __restrict__ colptr2)
{
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c <= width; c += gridDim.x * blockDim.x)
    {
        const unsigned int s = colptr[c];
        colptr2[c] = (s < nnz) ? pos[s] : new_nnz;
    }
}

__global__ void pp_compact_index(unsigned int count, const unsigned int* __restrict__ keep, const unsigned int* __restrict__ pos,
                                 const unsigned int* __restrict__ idx, unsigned int* __restrict__ idx2)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        if (keep[i]) idx2[pos[i]] = idx[i];
}

__global__ void pp_col_keep_len(unsigned int width, const unsigned int* __restrict__ colptr, unsigned int terms_per_doc, unsigned int* __restrict__ keep)
{
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < width; c += gridDim.x * blockDim.x)
        keep[c] = (colptr[c + 1] - colptr[c] >= terms_per_doc) ? 1u : 0u;
}

__global__ void pp_kept_len(unsigned int width, const unsigned int* __restrict__ colptr, const unsigned int* __restrict__ keep, unsigned int* __restrict__ len)
{
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < width; c += gridDim.x * blockDim.x)
        len[c] = keep[c] ? colptr[c + 1] - colptr[c] : 0u;
}

// PruneCols: a warp moves one kept column to its new place
__global__ void pp_move_columns(unsigned int width, const unsigned int* __restrict__ colptr, const unsigned int* __restrict__ keep,
                                const unsigned int* __restrict__ cpos, const unsigned int* __restrict__ newstart,
                                const unsigned int* __restrict__ rows, const unsigned int* __restrict__ counts,
                                unsigned int* __restrict__ colptr2, unsigned int* __restrict__ rows2, unsigned int* __restrict__ counts2)
{
    const int lane = threadIdx.x & 31;
    const unsigned int wpb = blockDim.x >> 5;
    for (unsigned int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < width; c += gridDim.x * wpb)
    {
        if (!keep[c]) continue;
        const unsigned int s = colptr[c], e = colptr[c + 1], d = newstart[c];
        if (lane == 0) colptr2[cpos[c]] = d;
        for (unsigned int i = s + lane; i < e; i += 32) { rows2[d + (i - s)] = rows[i]; counts2[d + (i - s)] = counts[i]; }
    }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// one warp per column: a 64-bit hash of the (position, row, count) triples and of the length
__global__ void pp_hash_columns(unsigned int width, const unsigned int* __restrict__ colptr, const unsigned int* __restrict__ rows,
                                const unsigned int* __restrict__ counts, unsigned long long* __restrict__ key, unsigned int* __restrict__ col)
{
    const int lane = threadIdx.x & 31;
    const unsigned int wpb = blockDim.x >> 5;
    for (unsigned int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < width; c += gridDim.x * wpb)
    {
        const unsigned int s = colptr[c], e = colptr[c + 1];
        unsigned long long h = 0ull;
        for (unsigned int i = s + lane; i < e; i += 32)
            h += mix64((static_cast<unsigned long long>(rows[i]) << 32 | counts[i]) ^ mix64(0x9e3779b97f4a7c15ull * (i - s + 1)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        if (lane == 0) { key[c] = mix64(h ^ (static_cast<unsigned long long>(e - s) << 1)); col[c] = c; }
    }
}

// sorted by hash (stable: equal hashes keep ascending column index). A column is a duplicate — and goes — iff a LATER column of
// the same run has exactly its content.
__global__ void pp_mark_duplicates(unsigned int width, const unsigned long long* __restrict__ key_sorted, const unsigned int* __restrict__ col_sorted,
                                   const unsigned int* __restrict__ colptr, const unsigned int* __restrict__ rows, const unsigned int* __restrict__ counts,
                                   unsigned int* __restrict__ mask)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < width; i += gridDim.x * blockDim.x)
    {
        const unsigned int c = col_sorted[i];
        const unsigned int s = colptr[c], len = colptr[c + 1] - s;
        unsigned int keep = 1u;
        for (unsigned int j = i + 1; j < width && key_sorted[j] == key_sorted[i]; ++j)
        {
            const unsigned int c2 = col_sorted[j], s2 = colptr[c2];
            if (colptr[c2 + 1] - s2 != len) continue;
            bool same = true;
            for (unsigned int t = 0; t < len; ++t)
                if (rows[s + t] != rows[s2 + t] || counts[s + t] != counts[s2 + t]) { same = false; break; }
            if (same) { keep = 0u; break; }
        }
        mask[c] = keep;
    }
}

// scores of one column by one warp: (1 + ln count) * ln(width / df[row]), scaled to unit 2-norm
__global__ void pp_scores(unsigned int width, const unsigned int* __restrict__ colptr, const unsigned int* __restrict__ rows,
                          const unsigned int* __restrict__ counts, const unsigned int* __restrict__ df, double* __restrict__ scores)
{
    const int lane = threadIdx.x & 31;
    const unsigned int wpb = blockDim.x >> 5;
    for (unsigned int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < width; c += gridDim.x * wpb)
    {
        const unsigned int s = colptr[c], e = colptr[c + 1];
        double ss = 0.0;
        for (unsigned int i = s + lane; i < e; i += 32)
        {
            const double v = (1.0 + log(static_cast<double>(counts[i]))) * log(static_cast<double>(width) / static_cast<double>(df[rows[i]]));
            scores[i] = v;
            ss += v * v;
        }
        ss = warp_sum(ss);
        const double scale = 1.0 / sqrt(ss);
        for (unsigned int i = s + lane; i < e; i += 32) scores[i] *= scale;
    }
}

__global__ void pp_count_nz(unsigned int nnz, const unsigned int* __restrict__ rows, unsigned int* __restrict__ df)
{
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) atomicAdd(df + rows[e], 1u);
}

__global__ void pp_iota(unsigned int n, unsigned int* __restrict__ v)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

struct Scratch
{
    DevBuf<unsigned char> tmp;
    void need(size_t bytes) { tmp.reserve(bytes); }
};

// exclusive scan of count flags into pos; returns the total (one 8-byte readback)
unsigned int scan_total(cudaStream_t st, Scratch& S, const unsigned int* flags, unsigned int* pos, unsigned int count)
{
    if (count == 0) return 0;
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, flags, pos, static_cast<int>(count), st);
    S.need(bytes);
    cub::DeviceScan::ExclusiveSum(S.tmp.p, bytes, flags, pos, static_cast<int>(count), st);
    unsigned int last[2];
    SMK_CUDA(cudaMemcpyAsync(&last[0], pos + count - 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SMK_CUDA(cudaMemcpyAsync(&last[1], flags + count - 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SMK_CUDA(cudaStreamSynchronize(st));
    return last[0] + last[1];
}

} // namespace

// Returns SMK_OK, or SMK_FAILURE when every document was pruned (the reference's preprocess_tf returns false).
int preprocess_tf_device(smk_ctx* c, unsigned int m, unsigned int n, unsigned int nnz, const unsigned int* col_offsets, const unsigned int* row_indices,
                         const double* counts_in, unsigned int max_iter, unsigned int docs_per_term, unsigned int terms_per_doc,
                         unsigned int* out_m, unsigned int* out_n, unsigned int* out_nnz, unsigned int* out_colptr, unsigned int* out_rows,
                         unsigned int* out_counts, double* out_scores, unsigned int* term_indices, unsigned int* doc_indices)
{
    cudaStream_t st = c->stream;
    Scratch S;
    const size_t cap_e = std::max<size_t>(nnz, 1), cap_c = static_cast<size_t>(n) + 1, cap_r = std::max<size_t>(m, 1);
    DevBuf<unsigned int> colptr[2], rows[2], cnts[2], term[2], doc[2], hist, hist_nz, keep, pos, flag, epos, len, newstart, colidx, colidx_s, ckeep;
    DevBuf<unsigned long long> key, key_s;
    DevBuf<double> dcounts, scores;
    for (int b = 0; b < 2; ++b) { colptr[b].reserve(cap_c); rows[b].reserve(cap_e); cnts[b].reserve(cap_e); term[b].reserve(cap_r); doc[b].reserve(cap_c); }
    const size_t cap_rc = std::max(cap_r, cap_c);
    hist.reserve(cap_r); hist_nz.reserve(cap_r); keep.reserve(cap_rc); pos.reserve(cap_rc); flag.reserve(cap_e); epos.reserve(cap_e);
    len.reserve(cap_c); newstart.reserve(cap_c); ckeep.reserve(cap_c); colidx.reserve(cap_c); colidx_s.reserve(cap_c); key.reserve(cap_c); key_s.reserve(cap_c);
    dcounts.reserve(cap_e); scores.reserve(cap_e);

    SMK_CUDA(cudaMemcpyAsync(colptr[0].p, col_offsets, sizeof(unsigned int) * (static_cast<size_t>(n) + 1), cudaMemcpyHostToDevice, st));
    if (nnz)
    {
        SMK_CUDA(cudaMemcpyAsync(rows[1].p, row_indices, sizeof(unsigned int) * nnz, cudaMemcpyHostToDevice, st));
        SMK_CUDA(cudaMemcpyAsync(dcounts.p, counts_in, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
        pp_counts_to_u32<<<blocks_for(nnz), kT, 0, st>>>(nnz, dcounts.p, cnts[1].p);
        SMK_LAUNCH_CHECK();
        // SortRows (:120): rows ascending inside every column, counts following (stable)
        size_t bytes = 0;
        cub::DeviceSegmentedSort::StableSortPairs(nullptr, bytes, rows[1].p, rows[0].p, cnts[1].p, cnts[0].p, static_cast<int>(nnz), static_cast<int>(n),
                                                  colptr[0].p, colptr[0].p + 1, st);
        S.need(bytes);
        cub::DeviceSegmentedSort::StableSortPairs(S.tmp.p, bytes, rows[1].p, rows[0].p, cnts[1].p, cnts[0].p, static_cast<int>(nnz), static_cast<int>(n),
                                                  colptr[0].p, colptr[0].p + 1, st);
    }
    pp_iota<<<blocks_for(m), kT, 0, st>>>(m, term[0].p);
    pp_iota<<<blocks_for(n), kT, 0, st>>>(n, doc[0].p);
    SMK_LAUNCH_CHECK();

    int cur = 0;                // which of the two buffers holds the matrix / the row index map / the column index map
    int tcur = 0, dcur = 0;
    unsigned int height = m, width = n, cnnz = nnz;

    // keeps the columns with mask[c] != 0 (order kept); returns the new width
    auto drop_columns = [&](const unsigned int* mask, unsigned int new_width) {
        scan_total(st, S, mask, pos.p, width);                                  // cpos
        pp_kept_len<<<blocks_for(width), kT, 0, st>>>(width, colptr[cur].p, mask, len.p);
        SMK_LAUNCH_CHECK();
        const unsigned int new_nnz = scan_total(st, S, len.p, newstart.p, width);
        const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((static_cast<long long>(width) + 7) / 8, 8LL * c->num_sms)));
        pp_move_columns<<<grid, kT, 0, st>>>(width, colptr[cur].p, mask, pos.p, newstart.p, rows[cur].p, cnts[cur].p, colptr[cur ^ 1].p, rows[cur ^ 1].p,
                                            cnts[cur ^ 1].p);
        SMK_LAUNCH_CHECK();
        SMK_CUDA(cudaMemcpyAsync(colptr[cur ^ 1].p + new_width, &new_nnz, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
        pp_compact_index<<<blocks_for(width), kT, 0, st>>>(width, mask, pos.p, doc[dcur].p, doc[dcur ^ 1].p);
        SMK_LAUNCH_CHECK();
        SMK_CUDA(cudaStreamSynchronize(st));                                    // new_nnz lives on the host stack
        cur ^= 1; dcur ^= 1;
        width = new_width; cnnz = new_nnz;
    };
    // mask of the columns that survive UniqueCols, in keep; returns their number
    auto unique_mask = [&]() -> unsigned int {
        if (width == 0) return 0;
        const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((static_cast<long long>(width) + 7) / 8, 8LL * c->num_sms)));
        pp_hash_columns<<<grid, kT, 0, st>>>(width, colptr[cur].p, rows[cur].p, cnts[cur].p, key.p, colidx.p);
        SMK_LAUNCH_CHECK();
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, key.p, key_s.p, colidx.p, colidx_s.p, static_cast<int>(width), 0, 64, st);
        S.need(bytes);
        cub::DeviceRadixSort::SortPairs(S.tmp.p, bytes, key.p, key_s.p, colidx.p, colidx_s.p, static_cast<int>(width), 0, 64, st);
        pp_mark_duplicates<<<blocks_for(width), kT, 0, st>>>(width, key_s.p, colidx_s.p, colptr[cur].p, rows[cur].p, cnts[cur].p, keep.p);
        SMK_LAUNCH_CHECK();
        return scan_total(st, S, keep.p, pos.p, width);
    };

    unsigned int it = 0;
    while (it < max_iter)
    {
        // ---- PruneRows
        SMK_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned int) * height, st));
        SMK_CUDA(cudaMemsetAsync(hist_nz.p, 0, sizeof(unsigned int) * height, st));
        if (cnnz) { pp_row_hist<<<blocks_for(cnnz), kT, 0, st>>>(cnnz, rows[cur].p, cnts[cur].p, hist.p, hist_nz.p); SMK_LAUNCH_CHECK(); }
        pp_row_keep<<<blocks_for(height), kT, 0, st>>>(height, hist.p, hist_nz.p, docs_per_term, width, keep.p);
        SMK_LAUNCH_CHECK();
        const unsigned int new_height = scan_total(st, S, keep.p, pos.p, height);       // pos = new row numbers
        if (new_height != height)
        {
            unsigned int new_nnz = 0;
            if (cnnz)
            {
                pp_entry_flag<<<blocks_for(cnnz), kT, 0, st>>>(cnnz, rows[cur].p, keep.p, flag.p);
                SMK_LAUNCH_CHECK();
                new_nnz = scan_total(st, S, flag.p, epos.p, cnnz);
                pp_compact_entries<<<blocks_for(cnnz), kT, 0, st>>>(cnnz, rows[cur].p, cnts[cur].p, flag.p, epos.p, pos.p, rows[cur ^ 1].p, cnts[cur ^ 1].p);
                SMK_LAUNCH_CHECK();
            }
            pp_colptr_after_rows<<<blocks_for(width + 1), kT, 0, st>>>(width, cnnz, new_nnz, colptr[cur].p, epos.p, colptr[cur ^ 1].p);
            SMK_LAUNCH_CHECK();
            pp_compact_index<<<blocks_for(height), kT, 0, st>>>(height, keep.p, pos.p, term[tcur].p, term[tcur ^ 1].p);
            SMK_LAUNCH_CHECK();
            cur ^= 1; tcur ^= 1;
            height = new_height; cnnz = new_nnz;
        }
        // ---- PrunableCols / UniqueCols / PruneCols (:126-180)
        pp_col_keep_len<<<blocks_for(width), kT, 0, st>>>(width, colptr[cur].p, terms_per_doc, ckeep.p);
        SMK_LAUNCH_CHECK();
        unsigned int new_width = scan_total(st, S, ckeep.p, pos.p, width);
        if (new_width == width)
        {
            new_width = unique_mask();
            if (new_width == width) break;
            drop_columns(keep.p, new_width);
        }
        else
        {
            if (new_width == 0) return SMK_FAILURE;
            drop_columns(ckeep.p, new_width);
            new_width = unique_mask();
            if (new_width != width) drop_columns(keep.p, new_width);
        }
        it += 1;
    }
    // ---- scores (:193-230)
    SMK_CUDA(cudaMemsetAsync(hist_nz.p, 0, sizeof(unsigned int) * std::max(height, 1u), st));
    if (cnnz)
    {
        pp_count_nz<<<blocks_for(cnnz), kT, 0, st>>>(cnnz, rows[cur].p, hist_nz.p);
        SMK_LAUNCH_CHECK();
        const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((static_cast<long long>(width) + 7) / 8, 8LL * c->num_sms)));
        pp_scores<<<grid, kT, 0, st>>>(width, colptr[cur].p, rows[cur].p, cnts[cur].p, hist_nz.p, scores.p);
        SMK_LAUNCH_CHECK();
    }
    *out_m = height; *out_n = width; *out_nnz = cnnz;
    SMK_CUDA(cudaMemcpyAsync(out_colptr, colptr[cur].p, sizeof(unsigned int) * (static_cast<size_t>(width) + 1), cudaMemcpyDeviceToHost, st));
    if (cnnz)
    {
        SMK_CUDA(cudaMemcpyAsync(out_rows, rows[cur].p, sizeof(unsigned int) * cnnz, cudaMemcpyDeviceToHost, st));
        SMK_CUDA(cudaMemcpyAsync(out_counts, cnts[cur].p, sizeof(unsigned int) * cnnz, cudaMemcpyDeviceToHost, st));
        SMK_CUDA(cudaMemcpyAsync(out_scores, scores.p, sizeof(double) * cnnz, cudaMemcpyDeviceToHost, st));
    }
    if (height) SMK_CUDA(cudaMemcpyAsync(term_indices, term[tcur].p, sizeof(unsigned int) * height, cudaMemcpyDeviceToHost, st));
    if (width) SMK_CUDA(cudaMemcpyAsync(doc_indices, doc[dcur].p, sizeof(unsigned int) * width, cudaMemcpyDeviceToHost, st));
    SMK_CUDA(cudaStreamSynchronize(st));
    return SMK_OK;
}

} // namespace smk
