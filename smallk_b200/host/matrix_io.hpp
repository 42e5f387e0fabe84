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
#include <string>
#include <vector>

namespace smallk_io {

// Column-major buffer (ld = height) from a delimited text file; rows of the file are matrix rows.
inline bool LoadDelimitedFile(std::vector<double>& buffer, unsigned int& height, unsigned int& width,
                              const std::string& filename, const char delim = ',')
{
    std::ifstream in(filename);
    if (!in) return false;
    std::vector<std::vector<double>> rows;
    std::string line;
    while (std::getline(in, line))
    {
        if (line.empty() || line == "\r") continue;
        std::vector<double> vals;
        const char* p = line.c_str();
        while (*p)
        {
            char* end = nullptr;
            double v = std::strtod(p, &end);
            if (end == p) break;
            vals.push_back(v);
            p = end;
            while (*p == delim || *p == ' ' || *p == '\t' || *p == '\r') ++p;
        }
        if (!vals.empty()) rows.push_back(std::move(vals));
    }
    if (rows.empty()) return false;
    height = static_cast<unsigned int>(rows.size());
    width = static_cast<unsigned int>(rows[0].size());
    for (const auto& r : rows) if (r.size() != width) return false;
    buffer.assign(static_cast<size_t>(height) * width, 0.0);
    for (unsigned int r = 0; r < height; ++r)
        for (unsigned int c = 0; c < width; ++c) buffer[static_cast<size_t>(c) * height + r] = rows[r][c];
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

// Sparse (coordinate) MatrixMarket -> CSC.
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
    for (unsigned long long e = 0; e < nz; ++e)
    {
        if (!std::getline(in, line)) return false;
        const char* p = line.c_str();
        char* end = nullptr;
        unsigned long r = std::strtoul(p, &end, 10); p = end;
        unsigned long c = std::strtoul(p, &end, 10); p = end;
        double v = pattern ? 1.0 : std::strtod(p, &end);
        if (r == 0 || c == 0 || r > h || c > w) return false;
        tr.push_back(static_cast<unsigned int>(r - 1)); tc.push_back(static_cast<unsigned int>(c - 1)); tv.push_back(v);
        if ((symmetric || skew) && r != c)
        {
            tr.push_back(static_cast<unsigned int>(c - 1)); tc.push_back(static_cast<unsigned int>(r - 1));
            tv.push_back(skew ? -v : v);
        }
    }
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
