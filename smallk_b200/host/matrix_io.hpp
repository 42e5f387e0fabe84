// smallk_b200 host — the file formats either side of the NMF path:
//   * delimited text, row-major rows of "%.{p}e" values (common/include/delimited_file.hpp:50-195)
//   * MatrixMarket coordinate files, real/integer/pattern x general/symmetric/skew-symmetric, expanded
//     and compressed to CSC by a stable counting sort on the column, rows kept in file order,
//     duplicates kept (common/include/sparse_matrix_io.hpp:118-260, sparse_matrix_impl.hpp:184-258)
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace smallk_io {

// Column-major buffer (ld = height) from a delimited text file; rows of the file are matrix rows.
// The reference's reader (delimited_file.hpp:79-135, delimited_file.cpp:36-103) defines the format by what it does, and
// files in the wild depend on it, so its rules are kept one by one (tests/test_file_formats.py holds both readers to the
// same buffers on well-formed and malformed input):
//   * blank lines and lines starting with '#' or '%' are skipped at the top of the file only;
//   * width = 1 + number of delimiters in the first data line; every later line is a row, blank or not;
//   * a last line without a terminating newline is not a row (and if it is the first data line the buffer stays zero);
//   * a row is read with formatted stream extraction, value then one non-blank separator character, `width` times:
//     short rows leave zeros, text that is not a number reads as 0 and ends the row, and so do "nan" / "inf".
inline bool LoadDelimitedFile(std::vector<double>& buffer, unsigned int& height, unsigned int& width,
                              const std::string& filename, const char delim = ',')
{
    std::ifstream in(filename);
    if (!in) return false;
    // every line with a flag: was it terminated by a newline (the stream had not hit end-of-file when it was read)?
    std::vector<std::pair<std::string, bool>> lines;
    {
        std::string text;
        while (std::getline(in, text)) lines.emplace_back(text, !in.eof());
    }
    size_t first = 0;
    while (first < lines.size() && (lines[first].first.empty() || lines[first].first[0] == '#' || lines[first].first[0] == '%')) ++first;
    if (first == lines.size()) return false;
    width = 1;
    for (const char ch : lines[first].first) if (ch == delim) ++width;
    height = 1;
    for (size_t i = first + 1; i < lines.size(); ++i) if (lines[i].second) ++height;
    buffer.resize(static_cast<size_t>(height) * width);
    unsigned int r = 0;
    for (size_t i = first; i < lines.size() && lines[i].second; ++i, ++r)
    {
        std::istringstream row(lines[i].first);
        char separator;
        for (unsigned int c = 0; c != width; ++c)
        {
            row >> buffer[static_cast<size_t>(c) * height + r];
            row >> separator;
        }
    }
    return true;
}

inline bool WriteDelimitedFile(const double* buffer, unsigned int ldim, unsigned int height, unsigned int width,
                               const std::string& filename, unsigned int precision, const char delim = ',')
{
    std::ofstream out(filename);
    if (!out) return false;
    out << std::scientific;
    out.precision(precision);
    for (unsigned int r = 0; r != height; ++r)
    {
        for (unsigned int c = 0; c != width - 1; ++c) out << buffer[static_cast<size_t>(c) * ldim + r] << delim;
        out << buffer[static_cast<size_t>(width - 1) * ldim + r] << std::endl;
    }
    return true;
}

struct CscMatrix
{
    unsigned int height = 0, width = 0;
    std::vector<unsigned int> col_offsets, row_indices;
    std::vector<double> data;
    unsigned int nnz() const { return static_cast<unsigned int>(data.size()); }
};

inline bool IsMatrixMarketFile(const std::string& filename)
{
    return filename.size() >= 4 && filename.compare(filename.size() - 4, 4, ".mtx") == 0;
}

// Sparse (coordinate) MatrixMarket -> CSC. Behaviour on malformed files follows the reference's reader
// (sparse_matrix_io.hpp:117-260 over SparseMatrix::Load / Compress, sparse_matrix_impl.hpp:158-258), which the tests hold it to:
//   * every line after the size line counts as an entry line, blank ones included (they carry no entry), and the count
//     must equal the declared number of entries, otherwise the file is rejected (returns false);
//   * an index of 0 is rejected; an index beyond the declared shape throws std::runtime_error("SparseMatrix::Load: row | col
//     index out of bounds"), a file without entries throws std::runtime_error("SparseMatrix::Compress: matrix has no data");
//   * entries are read with formatted stream extraction (a missing value reads as 0).
inline bool LoadMatrixMarketFile(const std::string& filename, CscMatrix& A)
{
    std::ifstream in(filename);
    if (!in) return false;
    std::string line;
    if (!std::getline(in, line)) return false;
    std::istringstream hdr(line);
    std::string banner, object, format, field, symmetry;
    hdr >> banner >> object >> format >> field >> symmetry;
    for (auto* s : {&object, &format, &field, &symmetry}) for (auto& ch : *s) ch = static_cast<char>(std::tolower(ch));
    if (banner != "%%MatrixMarket" || object != "matrix" || format != "coordinate") return false;
    const bool pattern = (field == "pattern");
    if (!pattern && field != "real" && field != "integer") return false;
    const bool symmetric = (symmetry == "symmetric"), skew = (symmetry == "skew-symmetric");
    if (!symmetric && !skew && symmetry != "general") return false;
    while (std::getline(in, line)) if (!line.empty() && line[0] != '%') break;
    unsigned long long h = 0, w = 0, nz = 0;
    { std::istringstream s(line); s >> h >> w >> nz; }
    if (h == 0 || w == 0) return false;
    std::vector<unsigned int> tr, tc;
    std::vector<double> tv;
    tr.reserve((symmetric || skew) ? 2 * nz : nz); tc.reserve(tr.capacity()); tv.reserve(tr.capacity());
    unsigned long long line_count = 0;
    double v = 0.0;
    while (std::getline(in, line))
    {
        ++line_count;
        if (line.empty()) continue;
        std::istringstream entry(line);
        unsigned int r = 0, c = 0;
        entry >> r; entry >> c;
        if (pattern) v = 1.0; else entry >> v;
        if (r == 0 || c == 0) return false;
        if (c > w) throw std::runtime_error("SparseMatrix::Load: col index out of bounds");
        if (r > h) throw std::runtime_error("SparseMatrix::Load: row index out of bounds");
        tr.push_back(r - 1); tc.push_back(c - 1); tv.push_back(v);
        if ((symmetric || skew) && r != c)
        {
            if (r > w) throw std::runtime_error("SparseMatrix::Load: col index out of bounds");
            if (c > h) throw std::runtime_error("SparseMatrix::Load: row index out of bounds");
            tr.push_back(c - 1); tc.push_back(r - 1); tv.push_back(skew ? -v : v);
        }
    }
    if (tv.empty()) throw std::runtime_error("SparseMatrix::Compress: matrix has no data");
    if (line_count != nz) return false;
    A.height = static_cast<unsigned int>(h); A.width = static_cast<unsigned int>(w);
    const size_t n = tv.size();
    A.col_offsets.assign(static_cast<size_t>(w) + 1, 0u);
    for (size_t e = 0; e < n; ++e) A.col_offsets[tc[e] + 1]++;
    for (size_t c = 0; c < w; ++c) A.col_offsets[c + 1] += A.col_offsets[c];
    std::vector<unsigned int> next(A.col_offsets.begin(), A.col_offsets.end() - 1);
    A.row_indices.resize(n); A.data.resize(n);
    for (size_t e = 0; e < n; ++e)          // stable: file order kept inside a column
    {
        const unsigned int d = next[tc[e]]++;
        A.row_indices[d] = tr[e];
        A.data[d] = tv[e];
    }
    return true;
}

} // namespace smallk_io
