// smallk_b200 — solver-level entry points used by the C ABI (internal header).
#pragma once

#include "context.h"

namespace smk {

void solver_alloc(smk_ctx* c);
void solver_init(smk_ctx* c);
void solver_step(smk_ctx* c);
int solver_progress(smk_ctx* c, double* metric);
void solver_progress_enqueue(smk_ctx* c, double* metric_dev);
int solver_run(smk_ctx* c, int count, double* metrics_host);
std::string solver_phase_report(smk_ctx* c);
// nmf_loop.cu: iterations [first_iter, max_iter) of NmfSolve's loop as one CUDA graph with a device-side WHILE (see the file)
bool nmf_loop_graph(smk_ctx* c, int first_iter, int* iter, bool* success, int* rc);
int solver_normalize(smk_ctx* c);
int solver_fail_iter(smk_ctx* c);
void solver_product(smk_ctx* c, int which);
// rank2_fused.cu: one Rank2 iteration on the active sparse matrix in three kernels (Wu: 2 * m doubles of scratch)
void rank2_fused_step(smk_ctx* c, double* Wu);
int solver_nnls_hals(smk_ctx* c, double tol, int max_iter, int* iterations);

// ---- submatrix.cu: SubMatrixColsCompact on the device (sparse_matrix_impl.hpp:479-591,
// dense_matrix_impl.hpp:224-285). Makes the listed columns of the loaded matrix the active matrix of the
// context; returns the new height and writes new_to_old_rows (host, >= m entries).
int select_columns(smk_ctx* c, const unsigned int* cols_host, int count, unsigned int* new_to_old_host);
void select_all(smk_ctx* c);

// ---- sort.cu: stable descending sort of a host array on the device; exactly one of order_host / sorted_host is set
void device_sort_desc(smk_ctx* c, const double* host_in, int n, int* order_host, double* sorted_host);

// ---- preprocess.cu: preprocess_tf (preprocessor/src/preprocess.cpp:81-250) on the device
int preprocess_tf_device(smk_ctx* c, unsigned int m, unsigned int n, unsigned int nnz, const unsigned int* col_offsets, const unsigned int* row_indices,
                         const double* counts_in, unsigned int max_iter, unsigned int docs_per_term, unsigned int terms_per_doc,
                         unsigned int* out_m, unsigned int* out_n, unsigned int* out_nnz, unsigned int* out_colptr, unsigned int* out_rows,
                         unsigned int* out_counts, double* out_scores, unsigned int* term_indices, unsigned int* doc_indices);

} // namespace smk
