// smallk_b200 host — the sparse input type that crosses the library boundary.
// Mirrors the public surface of the reference's SparseMatrix<T> (common/include/sparse_matrix_decl.hpp:21-132)
// that callers of ClustSparse / the file loaders use: compressed-column storage with 32-bit indices, triplet
// loading followed by a stable compression (rows keep load order inside a column, duplicates are kept:
// sparse_matrix_impl.hpp:184-258). All arithmetic on it happens on the GPU after smk_load_csc.
#pragma once

#include <stdexcept>
#include <vector>

template <typename T>
class SparseMatrix
{
public:
    SparseMatrix() : height_(0), width_(0), loading_(false) {}
    SparseMatrix(const unsigned int height, const unsigned int width, const unsigned int nzmax)
        : height_(height), width_(width), loading_(false)
    {
        col_offsets_.assign(static_cast<size_t>(width) + 1, 0u);
        row_indices_.reserve(nzmax); data_.reserve(nzmax);
    }
    SparseMatrix(const unsigned int height, const unsigned int width, const unsigned int nz,
                 const unsigned int* col_offsets, const unsigned int* row_indices, const T* data)
        : height_(height), width_(width), loading_(false),
          col_offsets_(col_offsets, col_offsets + width + 1), row_indices_(row_indices, row_indices + nz), data_(data, data + nz)
    {}

    unsigned int Height() const { return height_; }
    unsigned int Width() const { return width_; }
    unsigned int Size() const { return col_offsets_.empty() ? 0u : col_offsets_[width_]; }

    unsigned int* ColBuffer() { return &col_offsets_[0]; }
    unsigned int* RowBuffer() { return &row_indices_[0]; }
    T* DataBuffer() { return &data_[0]; }
    const unsigned int* LockedColBuffer() const { return &col_offsets_[0]; }
    const unsigned int* LockedRowBuffer() const { return &row_indices_[0]; }
    const T* LockedDataBuffer() const { return &data_[0]; }

    void Clear() { height_ = width_ = 0; col_offsets_.clear(); row_indices_.clear(); data_.clear(); tc_.clear(); }
    void Reserve(const unsigned int height, const unsigned int width, const unsigned int nzmax)
    {
        height_ = height; width_ = width;
        col_offsets_.assign(static_cast<size_t>(width) + 1, 0u);
        row_indices_.clear(); data_.clear(); tc_.clear();
        row_indices_.reserve(nzmax); data_.reserve(nzmax); tc_.reserve(nzmax);
    }

    // triplet loading: BeginLoad, Load(row, col, value)..., EndLoad
    void BeginLoad() { loading_ = true; row_indices_.clear(); data_.clear(); tc_.clear(); }
    void Load(const unsigned int row, const unsigned int col, const T& value)
    {
        if (!loading_) throw std::logic_error("SparseMatrix::Load: BeginLoad has not been called");
        if (row >= height_) height_ = row + 1;
        if (col >= width_) width_ = col + 1;
        row_indices_.push_back(row); tc_.push_back(col); data_.push_back(value);
    }
    void EndLoad()
    {
        // stable counting sort on the column index
        const size_t nz = data_.size();
        col_offsets_.assign(static_cast<size_t>(width_) + 1, 0u);
        for (size_t e = 0; e < nz; ++e) col_offsets_[tc_[e] + 1]++;
        for (size_t c = 0; c < width_; ++c) col_offsets_[c + 1] += col_offsets_[c];
        std::vector<unsigned int> next(col_offsets_.begin(), col_offsets_.end() - 1), rows(nz);
        std::vector<T> vals(nz);
        for (size_t e = 0; e < nz; ++e)
        {
            const unsigned int d = next[tc_[e]]++;
            rows[d] = row_indices_[e]; vals[d] = data_[e];
        }
        row_indices_.swap(rows); data_.swap(vals);
        tc_.clear(); tc_.shrink_to_fit();
        loading_ = false;
    }

private:
    unsigned int height_, width_;
    bool loading_;
    std::vector<unsigned int> col_offsets_, row_indices_;
    std::vector<T> data_;
    std::vector<unsigned int> tc_;      // column of each triplet while loading
};
