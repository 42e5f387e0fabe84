// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// El.hpp: a header-only stand-in for the ~27 Elemental 0.85 symbols that the
// reference's hot path touches, so the reference's own
// sources under /root/reference/{common,nmf,flatclust,hierclust,smallk} can be
// compiled UNMODIFIED into oracle/_ref/ by oracle/Makefile. Elemental itself
// cannot be built in this image (hard dependency on <mpi.h> and Fortran).
//
// What is restated from Elemental (sequential El::Matrix<double> only):
//   * Matrix<T> storage/view/resize semantics   modules/libelemental/src/core/Matrix.cpp:1033-1046
//   * Cholesky(UPPER) right-looking unblocked   modules/libelemental/src/lapack_like/factor/Cholesky/UVar3.hpp:17-44
//     (Elemental switches to a blocked variant above n=128; the shim keeps the
//      unblocked recurrence at every n — identical arithmetic for k <= 128.)
//   * cholesky::SolveAfter = two Trsm            .../factor/Cholesky/SolveAfter.hpp:17-42
//   * FrobeniusNorm scaled-square accumulation  .../props/Norm/Frobenius.hpp:16-29
//   * Transpose / Axpy / Scale / Zeros / Copy / DiagonalScale plain loops
// BLAS level-2/3 (dgemm, dgemv, dnrm2) go to the OpenBLAS that ships inside the
// Python venv (scipy_* symbol prefix) when SHIM_USE_OPENBLAS is defined, else
// to the plain loops below. Which one was linked is printed by ref_capi.
#pragma once

#include <cmath>
#include <cstring>
#include <vector>
#include <memory>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <functional>
#include <complex>

#ifdef SHIM_USE_OPENBLAS
extern "C" {
void scipy_dgemm_(const char*, const char*, const int*, const int*, const int*,
                  const double*, const double*, const int*, const double*, const int*,
                  const double*, double*, const int*);
void scipy_dgemv_(const char*, const int*, const int*, const double*, const double*,
                  const int*, const double*, const int*, const double*, double*, const int*);
double scipy_dnrm2_(const int*, const double*, const int*);
void scipy_openblas_set_num_threads(int);
}
#endif

namespace El {

typedef int Int;

enum UpperOrLower { LOWER, UPPER };
enum Orientation { NORMAL, TRANSPOSE, ADJOINT };
enum LeftOrRight { LEFT, RIGHT };
enum UnitOrNonUnit { NON_UNIT, UNIT };
enum NormType { ONE_NORM, INFINITY_NORM, ENTRYWISE_ONE_NORM, MAX_NORM, NUCLEAR_NORM, FROBENIUS_NORM, TWO_NORM };

class NonHPSDMatrixException : public std::runtime_error
{
public:
    NonHPSDMatrixException(const char* msg = "Matrix was not HPSD") : std::runtime_error(msg) {}
};

inline void LogicError(const std::string& s) { throw std::logic_error(s); }

namespace shim_state { inline bool& initialized() { static bool b = false; return b; } }
inline void Initialize(int&, char**&) { shim_state::initialized() = true; }
inline bool Initialized() { return shim_state::initialized(); }
inline void Finalize() { shim_state::initialized() = false; }

template <typename T>
class Matrix
{
public:
    Matrix() : h_(0), w_(0), ld_(1), data_(nullptr), view_(false), locked_(false) {}
    Matrix(Int h, Int w) : h_(h), w_(w), ld_(std::max(h, 1)), view_(false), locked_(false)
    {
        mem_.assign(static_cast<size_t>(ld_) * w_, T(0));
        data_ = mem_.data();
    }
    Matrix(Int h, Int w, T* buf, Int ld) : h_(h), w_(w), ld_(ld), data_(buf), view_(true), locked_(false) {}
    Matrix(Int h, Int w, const T* buf, Int ld)
        : h_(h), w_(w), ld_(ld), data_(const_cast<T*>(buf)), view_(true), locked_(true) {}
    Matrix(const Matrix<T>& o) : h_(0), w_(0), ld_(1), data_(nullptr), view_(false), locked_(false) { *this = o; }

    const Matrix<T>& operator=(const Matrix<T>& o)
    {
        if (this == &o) return *this;
        if (view_)
        {
            if (h_ != o.h_ || w_ != o.w_) LogicError("shim: cannot assign to a view of different size");
        }
        else
            Resize(o.h_, o.w_);
        for (Int j = 0; j < w_; ++j)
            std::memcpy(data_ + static_cast<size_t>(j) * ld_, o.data_ + static_cast<size_t>(j) * o.ld_, sizeof(T) * h_);
        return *this;
    }

    Int Height() const { return h_; }
    Int Width() const { return w_; }
    Int LDim() const { return ld_; }
    T* Buffer() { return data_; }
    const T* LockedBuffer() const { return data_; }
    T* Buffer(Int i, Int j) { return data_ + i + static_cast<size_t>(j) * ld_; }
    const T* LockedBuffer(Int i, Int j) const { return data_ + i + static_cast<size_t>(j) * ld_; }
    bool Locked() const { return locked_; }
    bool Viewing() const { return view_; }
    T Get(Int i, Int j) const { return data_[i + static_cast<size_t>(j) * ld_]; }
    void Set(Int i, Int j, T a) { data_[i + static_cast<size_t>(j) * ld_] = a; }
    void Update(Int i, Int j, T a) { data_[i + static_cast<size_t>(j) * ld_] += a; }

    void Attach(Int h, Int w, T* buf, Int ld)
    {
        mem_.clear(); h_ = h; w_ = w; ld_ = ld; data_ = buf; view_ = true; locked_ = false;
    }
    void LockedAttach(Int h, Int w, const T* buf, Int ld)
    {
        mem_.clear(); h_ = h; w_ = w; ld_ = ld; data_ = const_cast<T*>(buf); view_ = true; locked_ = true;
    }
    void Empty()
    {
        mem_.clear(); mem_.shrink_to_fit(); h_ = 0; w_ = 0; ld_ = 1; data_ = nullptr; view_ = false; locked_ = false;
    }
    // Matrix.cpp:1033-1046 — shrink in place when possible, else re-allocate
    // (contents are not preserved across a re-allocation).
    void Resize(Int h, Int w)
    {
        if (view_)
        {
            if (h > h_ || w > w_) LogicError("shim: cannot grow a view");
            h_ = h; w_ = w; return;
        }
        bool reallocate = h > ld_ || w > w_ || data_ == nullptr;
        h_ = h; w_ = w;
        if (reallocate)
        {
            ld_ = std::max(h, 1);
            size_t need = static_cast<size_t>(ld_) * w;
            if (mem_.size() < need) mem_.assign(need, T(0));
            data_ = mem_.data();
        }
    }

private:
    Int h_, w_, ld_;
    T* data_;
    std::vector<T> mem_;
    bool view_, locked_;
};

template <typename T>
inline void View(Matrix<T>& A, Matrix<T>& B, Int i, Int j, Int h, Int w)
{ A.Attach(h, w, B.Buffer(i, j), B.LDim()); }

template <typename T>
inline void LockedView(Matrix<T>& A, const Matrix<T>& B, Int i, Int j, Int h, Int w)
{ A.LockedAttach(h, w, B.LockedBuffer(i, j), B.LDim()); }

template <typename T>
inline void Zeros(Matrix<T>& A, Int h, Int w)
{
    A.Resize(h, w);
    for (Int j = 0; j < w; ++j)
        std::memset(A.Buffer(0, j), 0, sizeof(T) * h);
}

template <typename T>
inline void Copy(const Matrix<T>& X, Matrix<T>& Y) { Y = X; }

template <typename T>
inline void Axpy(T alpha, const Matrix<T>& X, Matrix<T>& Y)
{
    // Elemental's Axpy also accepts a row vector added to a column vector;
    // smallk only uses conforming shapes.
    if (X.Height() != Y.Height() || X.Width() != Y.Width()) LogicError("shim Axpy: nonconformal");
    const Int h = X.Height(), w = X.Width();
    for (Int j = 0; j < w; ++j)
    {
        const T* x = X.LockedBuffer(0, j);
        T* y = Y.Buffer(0, j);
        for (Int i = 0; i < h; ++i) y[i] += alpha * x[i];
    }
}

template <typename T>
inline void Scale(T alpha, Matrix<T>& X)
{
    const Int h = X.Height(), w = X.Width();
    for (Int j = 0; j < w; ++j)
    {
        T* x = X.Buffer(0, j);
        for (Int i = 0; i < h; ++i) x[i] *= alpha;
    }
}

template <typename T>
inline void Transpose(const Matrix<T>& A, Matrix<T>& B)
{
    const Int m = A.Height(), n = A.Width();
    B.Resize(n, m);
    for (Int j = 0; j < n; ++j)
        for (Int i = 0; i < m; ++i)
            B.Set(j, i, A.Get(i, j));
}

template <typename T>
inline T Nrm2(const Matrix<T>& x)
{
    if (x.Height() != 1 && x.Width() != 1) LogicError("shim Nrm2: expected vector");
    const int n = (x.Width() == 1 ? x.Height() : x.Width());
    const int inc = (x.Width() == 1 ? 1 : x.LDim());
#ifdef SHIM_USE_OPENBLAS
    return scipy_dnrm2_(&n, x.LockedBuffer(), &inc);
#else
    // reference-BLAS dnrm2 scaled accumulation
    T scale = 0, ssq = 1;
    const T* p = x.LockedBuffer();
    for (int i = 0; i < n; ++i)
    {
        T v = p[static_cast<size_t>(i) * inc];
        if (v != T(0))
        {
            T a = std::abs(v);
            if (scale < a) { ssq = T(1) + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    }
    return scale * std::sqrt(ssq);
#endif
}

// modules/libelemental/include/El/core/... UpdateScaledSquare
template <typename T>
inline void UpdateScaledSquare(T alpha, T& scale, T& scaledSquare)
{
    T alphaAbs = std::abs(alpha);
    if (alphaAbs != 0)
    {
        if (alphaAbs <= scale)
        {
            const T relMag = alphaAbs / scale;
            scaledSquare += relMag * relMag;
        }
        else
        {
            const T relMag = scale / alphaAbs;
            scaledSquare = scaledSquare * relMag * relMag + 1;
            scale = alphaAbs;
        }
    }
}

template <typename T>
inline T FrobeniusNorm(const Matrix<T>& A)
{
    T scale = 0, scaledSquare = 1;
    for (Int j = 0; j < A.Width(); ++j)
        for (Int i = 0; i < A.Height(); ++i)
            UpdateScaledSquare(A.Get(i, j), scale, scaledSquare);
    return scale * std::sqrt(scaledSquare);
}

template <typename T>
inline T Norm(const Matrix<T>& A, NormType type = FROBENIUS_NORM)
{
    switch (type)
    {
    case FROBENIUS_NORM: return FrobeniusNorm(A);
    case MAX_NORM:
    {
        T m = 0;
        for (Int j = 0; j < A.Width(); ++j)
            for (Int i = 0; i < A.Height(); ++i) m = std::max(m, std::abs(A.Get(i, j)));
        return m;
    }
    case ONE_NORM:
    {
        T m = 0;
        for (Int j = 0; j < A.Width(); ++j)
        {
            T s = 0;
            for (Int i = 0; i < A.Height(); ++i) s += std::abs(A.Get(i, j));
            m = std::max(m, s);
        }
        return m;
    }
    case INFINITY_NORM:
    {
        T m = 0;
        for (Int i = 0; i < A.Height(); ++i)
        {
            T s = 0;
            for (Int j = 0; j < A.Width(); ++j) s += std::abs(A.Get(i, j));
            m = std::max(m, s);
        }
        return m;
    }
    default: LogicError("shim Norm: unsupported norm type");
    }
    return T(0);
}

template <typename T>
inline void Gemm(Orientation oA, Orientation oB, T alpha, const Matrix<T>& A, const Matrix<T>& B, T beta, Matrix<T>& C)
{
    const int m = C.Height(), n = C.Width();
    const int kk = (oA == NORMAL ? A.Width() : A.Height());
    {
        const int am = (oA == NORMAL ? A.Height() : A.Width());
        const int bk = (oB == NORMAL ? B.Height() : B.Width());
        const int bn = (oB == NORMAL ? B.Width() : B.Height());
        if (am != m || bn != n || bk != kk) LogicError("shim Gemm: nonconformal");
    }
#ifdef SHIM_USE_OPENBLAS
    const char ta = (oA == NORMAL ? 'N' : 'T'), tb = (oB == NORMAL ? 'N' : 'T');
    const int lda = A.LDim(), ldb = B.LDim(), ldc = C.LDim();
    if (m > 0 && n > 0)
        scipy_dgemm_(&ta, &tb, &m, &n, &kk, &alpha, A.LockedBuffer(), &lda, B.LockedBuffer(), &ldb, &beta, C.Buffer(), &ldc);
#else
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i)
        {
            T s = 0;
            for (int p = 0; p < kk; ++p)
            {
                T a = (oA == NORMAL ? A.Get(i, p) : A.Get(p, i));
                T b = (oB == NORMAL ? B.Get(p, j) : B.Get(j, p));
                s += a * b;
            }
            C.Set(i, j, (beta == T(0) ? T(0) : beta * C.Get(i, j)) + alpha * s);
        }
#endif
}

template <typename T>
inline void Gemv(Orientation o, T alpha, const Matrix<T>& A, const Matrix<T>& x, T beta, Matrix<T>& y)
{
    // Elemental treats x and y as vectors regardless of row/column storage.
    const int m = A.Height(), n = A.Width();
    const int incx = (x.Width() == 1 ? 1 : x.LDim());
    const int incy = (y.Width() == 1 ? 1 : y.LDim());
    const int xlen = (x.Width() == 1 ? x.Height() : x.Width());
    const int ylen = (y.Width() == 1 ? y.Height() : y.Width());
    if (o == NORMAL) { if (xlen != n || ylen != m) LogicError("shim Gemv: nonconformal"); }
    else             { if (xlen != m || ylen != n) LogicError("shim Gemv: nonconformal"); }
#ifdef SHIM_USE_OPENBLAS
    const char t = (o == NORMAL ? 'N' : 'T');
    const int lda = A.LDim();
    if (m > 0 && n > 0)
        scipy_dgemv_(&t, &m, &n, &alpha, A.LockedBuffer(), &lda, x.LockedBuffer(), &incx, &beta, y.Buffer(), &incy);
#else
    const T* xp = x.LockedBuffer();
    T* yp = y.Buffer();
    for (int i = 0; i < ylen; ++i)
    {
        T s = 0;
        for (int p = 0; p < xlen; ++p)
            s += (o == NORMAL ? A.Get(i, p) : A.Get(p, i)) * xp[static_cast<size_t>(p) * incx];
        T& yy = yp[static_cast<size_t>(i) * incy];
        yy = (beta == T(0) ? T(0) : beta * yy) + alpha * s;
    }
#endif
}

template <typename T>
inline void DiagonalScale(LeftOrRight side, Orientation, const Matrix<T>& d, Matrix<T>& X)
{
    const Int m = X.Height(), n = X.Width();
    if (side == LEFT)
    {
        for (Int j = 0; j < n; ++j)
            for (Int i = 0; i < m; ++i) X.Set(i, j, X.Get(i, j) * d.Get(i, 0));
    }
    else
    {
        for (Int j = 0; j < n; ++j)
        {
            const T s = d.Get(j, 0);
            for (Int i = 0; i < m; ++i) X.Set(i, j, X.Get(i, j) * s);
        }
    }
}

// Plain-loop triangular solves, one right-hand side at a time.
template <typename T>
inline void Trsm(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag,
                 T alpha, const Matrix<T>& A, Matrix<T>& B)
{
    const Int n = A.Height();
    const bool unit = (diag == UNIT);
    if (alpha != T(1)) Scale(alpha, B);
    if (side == LEFT)
    {
        if (B.Height() != n) LogicError("shim Trsm: nonconformal");
        for (Int c = 0; c < B.Width(); ++c)
        {
            T* b = B.Buffer(0, c);
            const bool upper_eff = (uplo == UPPER) == (o == NORMAL);   // effective op(A) is upper?
            if (!upper_eff)
            {
                // forward substitution with op(A) lower
                for (Int i = 0; i < n; ++i)
                {
                    T s = b[i];
                    for (Int p = 0; p < i; ++p)
                        s -= (o == NORMAL ? A.Get(i, p) : A.Get(p, i)) * b[p];
                    b[i] = unit ? s : s / A.Get(i, i);
                }
            }
            else
            {
                for (Int i = n - 1; i >= 0; --i)
                {
                    T s = b[i];
                    for (Int p = i + 1; p < n; ++p)
                        s -= (o == NORMAL ? A.Get(i, p) : A.Get(p, i)) * b[p];
                    b[i] = unit ? s : s / A.Get(i, i);
                }
            }
        }
    }
    else
    {
        // X op(A) = B  <=>  op(A)^T X^T = B^T ; solve row by row
        if (B.Width() != n) LogicError("shim Trsm: nonconformal");
        for (Int r = 0; r < B.Height(); ++r)
        {
            const bool upper_eff = (uplo == UPPER) == (o == NORMAL);
            // op(A)^T is lower when op(A) is upper
            if (upper_eff)
            {
                for (Int i = 0; i < n; ++i)
                {
                    T s = B.Get(r, i);
                    for (Int p = 0; p < i; ++p)
                        s -= (o == NORMAL ? A.Get(p, i) : A.Get(i, p)) * B.Get(r, p);
                    B.Set(r, i, unit ? s : s / A.Get(i, i));
                }
            }
            else
            {
                for (Int i = n - 1; i >= 0; --i)
                {
                    T s = B.Get(r, i);
                    for (Int p = i + 1; p < n; ++p)
                        s -= (o == NORMAL ? A.Get(p, i) : A.Get(i, p)) * B.Get(r, p);
                    B.Set(r, i, unit ? s : s / A.Get(i, i));
                }
            }
        }
    }
}

namespace cholesky {

// factor/Cholesky/UVar3.hpp:17-44 (right-looking, upper, unblocked)
template <typename T>
inline void UVar3Unb(Matrix<T>& A)
{
    const Int n = A.Height();
    const Int lda = A.LDim();
    T* a = A.Buffer();
    for (Int j = 0; j < n; ++j)
    {
        T alpha = a[j + static_cast<size_t>(j) * lda];
        if (alpha <= T(0)) throw NonHPSDMatrixException("A was not numerically HPD");
        alpha = std::sqrt(alpha);
        a[j + static_cast<size_t>(j) * lda] = alpha;
        for (Int k = j + 1; k < n; ++k) a[j + static_cast<size_t>(k) * lda] /= alpha;
        for (Int k = j + 1; k < n; ++k)
            for (Int i = j + 1; i <= k; ++i)
                a[i + static_cast<size_t>(k) * lda] -= a[j + static_cast<size_t>(i) * lda] * a[j + static_cast<size_t>(k) * lda];
    }
}

template <typename T>
inline void LVar3Unb(Matrix<T>& A)
{
    const Int n = A.Height();
    for (Int j = 0; j < n; ++j)
    {
        T alpha = A.Get(j, j);
        if (alpha <= T(0)) throw NonHPSDMatrixException("A was not numerically HPD");
        alpha = std::sqrt(alpha);
        A.Set(j, j, alpha);
        for (Int i = j + 1; i < n; ++i) A.Set(i, j, A.Get(i, j) / alpha);
        for (Int k = j + 1; k < n; ++k)
            for (Int i = k; i < n; ++i)
                A.Set(i, k, A.Get(i, k) - A.Get(i, j) * A.Get(k, j));
    }
}

template <typename T>
inline void SolveAfter(UpperOrLower uplo, Orientation, const Matrix<T>& A, Matrix<T>& B)
{
    if (uplo == LOWER)
    {
        Trsm(LEFT, LOWER, NORMAL, NON_UNIT, T(1), A, B);
        Trsm(LEFT, LOWER, ADJOINT, NON_UNIT, T(1), A, B);
    }
    else
    {
        Trsm(LEFT, UPPER, ADJOINT, NON_UNIT, T(1), A, B);
        Trsm(LEFT, UPPER, NORMAL, NON_UNIT, T(1), A, B);
    }
}

} // namespace cholesky

template <typename T>
inline void Cholesky(UpperOrLower uplo, Matrix<T>& A)
{
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (uplo == UPPER) cholesky::UVar3Unb(A); else cholesky::LVar3Unb(A);
}

// src/lapack_like/solve/HPDSolve.cpp:14-21
template <typename T>
inline void HPDSolve(UpperOrLower uplo, Orientation o, Matrix<T>& A, Matrix<T>& B)
{
    Cholesky(uplo, A);
    cholesky::SolveAfter(uplo, o, A, B);
}

// ---- wrappers that exist in dense_matrix_ops.hpp but are never reached by the
// ---- solvers (SURVEY.md §2.3 last row): present so the headers compile.
namespace lu {
template <typename T, typename P>
inline void SolveAfter(Orientation, const Matrix<T>&, const Matrix<P>&, Matrix<T>&)
{ LogicError("shim: lu::SolveAfter not provided (off the hot path)"); }
}
template <typename T, typename P>
inline void LU(Matrix<T>&, Matrix<P>&) { LogicError("shim: LU not provided (off the hot path)"); }
template <typename T>
inline void GaussianElimination(Matrix<T>&, Matrix<T>&) { LogicError("shim: GaussianElimination not provided"); }
template <typename T>
inline void Pseudoinverse(Matrix<T>&) { LogicError("shim: Pseudoinverse not provided"); }
template <typename T>
inline void LeastSquares(Orientation, Matrix<T>&, const Matrix<T>&, Matrix<T>&) { LogicError("shim: LeastSquares not provided"); }
template <typename T>
inline void MakeUniform(Matrix<T>&) { LogicError("shim: MakeUniform not provided"); }
template <typename T>
inline T Dot(const Matrix<T>& x, const Matrix<T>& y)
{
    T s = 0;
    const Int n = (x.Width() == 1 ? x.Height() : x.Width());
    for (Int i = 0; i < n; ++i)
        s += (x.Width() == 1 ? x.Get(i, 0) : x.Get(0, i)) * (y.Width() == 1 ? y.Get(i, 0) : y.Get(0, i));
    return s;
}
template <typename T>
inline void Print(const Matrix<T>&, const std::string& = "") {}

} // namespace El
