// TEST INFRASTRUCTURE ONLY - a minimal stand-in for the parts of Eigen 3 that the reference's B-spline, Givens / implicit-QR SVD and
// fixed-corotated model headers use, so that those headers compile FROM WHERE THEY LIE under /root/reference into
// oracle/_ref/libziran_ref.so (Eigen itself is not in this image; see oracle/Makefile).
// Fixed-size, column-major, eager evaluation (every operator returns a plain Matrix; col / row / block are lvalue views).
// Arithmetic conventions that matter for parity are Eigen's: products sum over k = 0, 1, 2 in order; the 3x3 determinant is the
// cofactor expansion along the first row.  SelfAdjointEigenSolver is a cyclic Jacobi iteration (NOT Eigen's tridiagonal QL): the
// PSD projection built from it (EigenDecomposition.h:126-135) is unique mathematically, so results agree to rounding.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STRONG_INLINE inline
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3

namespace Eigen {

enum { Dynamic = -1 };
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
typedef std::ptrdiff_t Index;
template <class T>
using aligned_allocator = std::allocator<T>;

template <class T, int R, int C, int O = 0, int MR = R, int MC = C>
class Matrix;
template <class X, int BR, int BC>
class Block;
template <class T, int N, int M = N>
class DiagonalMatrix;
template <class M>
class Map; // declared only
template <class T, int O = 0, class I = int>
class SparseMatrix; // declared only
template <class T, class I = int>
class Triplet; // declared only
template <class V>
class VectorBlock; // declared only
template <class M, int Q = 0>
class JacobiSVD; // declared only

template <class T, int R, int C>
struct FixedArray;
template <class D>
struct traits;
template <class T, int R, int C, int O, int MR, int MC>
struct traits<Matrix<T, R, C, O, MR, MC>> {
    typedef T Scalar;
    enum { Rows = R, Cols = C };
};
template <class X, int BR, int BC>
struct traits<Block<X, BR, BC>> {
    typedef typename traits<X>::Scalar Scalar;
    enum { Rows = BR, Cols = BC };
};

template <class Derived>
class MatrixBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    typedef std::ptrdiff_t Index;
    enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols,
        SizeAtCompileTime = (traits<Derived>::Rows == Dynamic || traits<Derived>::Cols == Dynamic) ? Dynamic : traits<Derived>::Rows * traits<Derived>::Cols };
    typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;

    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    static constexpr int rows() { return RowsAtCompileTime; }
    static constexpr int cols() { return ColsAtCompileTime; }
    static constexpr int size() { return RowsAtCompileTime * ColsAtCompileTime; }

    Scalar& operator()(int i, int j) { return derived().coeffRef(i, j); }
    const Scalar& operator()(int i, int j) const { return derived().coeff(i, j); }
    Scalar& operator()(int i) { return derived().coeffRef(i % RowsAtCompileTime, i / RowsAtCompileTime); }
    const Scalar& operator()(int i) const { return derived().coeff(i % RowsAtCompileTime, i / RowsAtCompileTime); }
    Scalar& operator[](int i) { return (*this)(i); }
    const Scalar& operator[](int i) const { return (*this)(i); }
    Scalar& x() { return (*this)(0); }
    Scalar& y() { return (*this)(1); }
    Scalar& z() { return (*this)(2); }
    const Scalar& x() const { return (*this)(0); }
    const Scalar& y() const { return (*this)(1); }
    const Scalar& z() const { return (*this)(2); }

    PlainObject eval() const
    {
        PlainObject r;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) r(i, j) = (*this)(i, j);
        return r;
    }
    Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> transpose() const
    {
        Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> r;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) r(j, i) = (*this)(i, j);
        return r;
    }
    // lvalue views (the reference writes through const references by const_cast, ImplicitQRSVD.h:63,108: so do these)
    Block<Derived, RowsAtCompileTime, 1> col(int j) const { return Block<Derived, RowsAtCompileTime, 1>(const_cast<Derived&>(derived()), 0, j); }
    Block<Derived, 1, ColsAtCompileTime> row(int i) const { return Block<Derived, 1, ColsAtCompileTime>(const_cast<Derived&>(derived()), i, 0); }
    template <int BR, int BC>
    Block<Derived, BR, BC> block(int i, int j) const { return Block<Derived, BR, BC>(const_cast<Derived&>(derived()), i, j); }

    // (same-type assignment through the base must copy coefficients too: the implicit copy assignment of this empty base would not)
    MatrixBase& operator=(const MatrixBase& o)
    {
        if (this != &o) {
            const PlainObject t = o.eval();
            for (int j = 0; j < cols(); ++j)
                for (int i = 0; i < rows(); ++i) (*this)(i, j) = t(i, j);
        }
        return *this;
    }
    MatrixBase() = default;
    MatrixBase(const MatrixBase&) = default;
    template <class O>
    Derived& operator=(const MatrixBase<O>& o)
    {
        static_assert((int)O::RowsAtCompileTime == (int)RowsAtCompileTime && (int)O::ColsAtCompileTime == (int)ColsAtCompileTime, "size mismatch");
        const typename MatrixBase<O>::PlainObject t = o.eval(); // aliasing-safe
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) = t(i, j);
        return derived();
    }
    Derived& noalias() { return derived(); }
    template <class O>
    Derived& operator+=(const MatrixBase<O>& o)
    {
        const typename MatrixBase<O>::PlainObject t = o.eval();
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) += t(i, j);
        return derived();
    }
    template <class O>
    Derived& operator-=(const MatrixBase<O>& o)
    {
        const typename MatrixBase<O>::PlainObject t = o.eval();
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) -= t(i, j);
        return derived();
    }
    Derived& operator*=(const Scalar& s)
    {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) *= s;
        return derived();
    }
    Derived& operator/=(const Scalar& s)
    {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) /= s;
        return derived();
    }
    template <class O>
    void swap(const MatrixBase<O>& o_)
    {
        MatrixBase<O>& o = const_cast<MatrixBase<O>&>(o_);
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) std::swap((*this)(i, j), o(i, j));
    }
    Derived& setZero()
    {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) = Scalar(0);
        return derived();
    }
    Derived& setConstant(const Scalar& v)
    {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) = v;
        return derived();
    }
    Derived& setIdentity()
    {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) (*this)(i, j) = i == j ? Scalar(1) : Scalar(0);
        return derived();
    }
    Scalar squaredNorm() const
    {
        Scalar s = 0;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s += (*this)(i, j) * (*this)(i, j);
        return s;
    }
    Scalar norm() const { using std::sqrt; return sqrt(squaredNorm()); }
    Scalar sum() const
    {
        Scalar s = 0;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s += (*this)(i, j);
        return s;
    }
    Scalar prod() const
    {
        Scalar s = 1;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s *= (*this)(i, j);
        return s;
    }
    Scalar trace() const
    {
        Scalar s = 0;
        for (int i = 0; i < rows(); ++i) s += (*this)(i, i);
        return s;
    }
    template <class O>
    Scalar dot(const MatrixBase<O>& o) const
    {
        Scalar s = 0;
        for (int i = 0; i < size(); ++i) s += (*this)(i) * o(i);
        return s;
    }
    template <class O>
    PlainObject cwiseProduct(const MatrixBase<O>& o) const
    {
        PlainObject r;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) r(i, j) = (*this)(i, j) * o(i, j);
        return r;
    }
    Scalar determinant() const
    {
        static_assert((int)RowsAtCompileTime == (int)ColsAtCompileTime && RowsAtCompileTime <= 3, "determinant: up to 3x3");
        const MatrixBase& m = *this;
        if (RowsAtCompileTime == 1) return m(0, 0);
        if (RowsAtCompileTime == 2) return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
        return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0))
            + m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
    }
    DiagonalMatrix<Scalar, RowsAtCompileTime> asDiagonal() const { return DiagonalMatrix<Scalar, RowsAtCompileTime>(eval()); }
    Matrix<Scalar, RowsAtCompileTime, 1> diagonal() const
    {
        Matrix<Scalar, RowsAtCompileTime, 1> d;
        for (int i = 0; i < rows(); ++i) d(i) = (*this)(i, i);
        return d;
    }
    // 3 x 3 inverse the way Eigen computes it (Eigen/src/LU/InverseImpl.h, compute_inverse_size3_helper): cofactors, determinant from
    // the first column, one multiplication by 1 / det per entry
    PlainObject inverse() const
    {
        static_assert((RowsAtCompileTime == 3 && ColsAtCompileTime == 3) || (RowsAtCompileTime == 2 && ColsAtCompileTime == 2), "mini_eigen: inverse() is implemented for 2 x 2 and 3 x 3 only");
        const MatrixBase& m = *this;
        if (RowsAtCompileTime == 2) { // Eigen/src/LU/InverseImpl.h, compute_inverse<.., 2>: adjugate times 1 / det
            const Scalar invdet2 = Scalar(1) / (m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1));
            PlainObject r2;
            r2(0, 0) = m(1, 1) * invdet2; r2(1, 0) = -m(1, 0) * invdet2; r2(0, 1) = -m(0, 1) * invdet2; r2(1, 1) = m(0, 0) * invdet2;
            return r2;
        }
        auto cof = [&m](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
        };
        const Scalar c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const Scalar det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
        const Scalar invdet = Scalar(1) / det;
        PlainObject r;
        r(0, 0) = c00 * invdet; r(0, 1) = c10 * invdet; r(0, 2) = c20 * invdet;
        r(1, 0) = cof(0, 1) * invdet; r(1, 1) = cof(1, 1) * invdet; r(1, 2) = cof(2, 1) * invdet;
        r(2, 0) = cof(0, 2) * invdet; r(2, 1) = cof(1, 2) * invdet; r(2, 2) = cof(2, 2) * invdet;
        return r;
    }
    template <class U>
    Matrix<U, RowsAtCompileTime, ColsAtCompileTime> cast() const
    {
        Matrix<U, RowsAtCompileTime, ColsAtCompileTime> r;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) r(i, j) = (U)(*this)(i, j);
        return r;
    }
    static PlainObject Zero() { PlainObject r; r.setZero(); return r; }
    static PlainObject Zero(int, int) { return Zero(); }
    template <class O> void resizeLike(const O&) {}
    void resize(int, int) {}   // fixed size: Eigen accepts a resize to the same dimensions (BinaryIO.h:139)
    PlainObject normalized() const { PlainObject r = eval(); const Scalar n = norm(); if (n > Scalar(0)) { for (int k = 0; k < size(); ++k) r(k) = r(k) / n; } return r; }
    void normalize() { derived() = normalized(); }
    template <class O>
    PlainObject cross(const MatrixBase<O>& o) const
    {
        static_assert(RowsAtCompileTime * ColsAtCompileTime == 3, "cross: 3-vectors");
        PlainObject r;
        const MatrixBase& a = *this;
        r(0) = a(1) * o(2) - a(2) * o(1); r(1) = a(2) * o(0) - a(0) * o(2); r(2) = a(0) * o(1) - a(1) * o(0);
        return r;
    }
    template <int BR, int BC> Block<Derived, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    void transposeInPlace() { const PlainObject t = eval(); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) (*this)(i, j) = t(j, i); }
    template <class O>
    bool operator!=(const MatrixBase<O>& o) const { for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) if ((*this)(i, j) != o(i, j)) return true; return false; }
    template <int N> Block<Derived, N, 1> head() const { return Block<Derived, N, 1>(const_cast<Derived&>(derived()), 0, 0); }
    FixedArray<Scalar, RowsAtCompileTime, ColsAtCompileTime> array() const;
    template <int BR, int BC> Block<Derived, BR, BC> topRightCorner() const { return block<BR, BC>(0, ColsAtCompileTime - BC); }
    Scalar maxCoeff() const { Scalar m = (*this)(0); for (int k = 1; k < size(); ++k) if ((*this)(k) > m) m = (*this)(k); return m; }
    Scalar minCoeff() const { Scalar m = (*this)(0); for (int k = 1; k < size(); ++k) if ((*this)(k) < m) m = (*this)(k); return m; }
    static PlainObject Identity() { PlainObject r; r.setIdentity(); return r; }
    static PlainObject Constant(const Scalar& v) { PlainObject r; r.setConstant(v); return r; }
    static PlainObject Ones() { return Constant(Scalar(1)); }
    static PlainObject Unit(int k) { PlainObject r; r.setZero(); r(k) = Scalar(1); return r; }

    // m << a, b, c;  (row-major fill like Eigen's CommaInitializer)
    struct Comma {
        MatrixBase& m;
        int k;
        Comma& operator,(const Scalar& v)
        {
            m(k / ColsAtCompileTime, k % ColsAtCompileTime) = v;
            ++k;
            return *this;
        }
    };
    Comma operator<<(const Scalar& v)
    {
        (*this)(0, 0) = v;
        return Comma{*this, 1};
    }
};

template <class T, int R, int C, int O, int MR, int MC>
class Matrix : public MatrixBase<Matrix<T, R, C, O, MR, MC>> {
    // Eigen aligns fixed-size objects whose byte size is a multiple of 16 (of 32 in an AVX build: EIGEN_MAX_STATIC_ALIGN_BYTES); GridState's "size is a power of
    // two" static assertions (Lib/MPM/MpmGrid.h:39-42) rely on it for the 2D records
    static constexpr std::size_t bytes_ = sizeof(T) * (R > 0 ? R : 1) * (C > 0 ? C : 1);
    alignas(bytes_ % 32 == 0 ? 32 : bytes_ % 16 == 0 ? 16 : alignof(T)) T m_[(R > 0 ? R : 1) * (C > 0 ? C : 1)];

public:
    typedef MatrixBase<Matrix> Base;
    typedef T Scalar;
    Matrix() {}
    template <class Od>
    Matrix(const MatrixBase<Od>& o) { Base::template operator=<Od>(o); }
    Matrix(const Matrix& o) : Base() { for (int k = 0; k < (R > 0 ? R : 1) * (C > 0 ? C : 1); ++k) m_[k] = o.m_[k]; }
    template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 2, int>::type = 0>
    Matrix(const T& a, const T& b) { m_[0] = a; m_[1] = b; }
    template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 3, int>::type = 0>
    Matrix(const T& a, const T& b, const T& c) { m_[0] = a; m_[1] = b; m_[2] = c; }
    template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 4, int>::type = 0>
    Matrix(const T& a, const T& b, const T& c, const T& d) { m_[0] = a; m_[1] = b; m_[2] = c; m_[3] = d; }
    Matrix& operator=(const Matrix& o) { for (int k = 0; k < (R > 0 ? R : 1) * (C > 0 ? C : 1); ++k) m_[k] = o.m_[k]; return *this; }
    template <class Od>
    Matrix& operator=(const MatrixBase<Od>& o) { return Base::operator=(o); }
    T& coeffRef(int i, int j) { return m_[i + j * R]; }
    const T& coeff(int i, int j) const { return m_[i + j * R]; }
    T* data() { return m_; }
    const T* data() const { return m_; }
};

// coefficient-wise view of a fixed-size matrix: what the reference's geometry code uses of Eigen's Array (comparisons + any / all, min / max, abs,
// maxCoeff, difference); converts back to a Matrix
template <class T, int R, int C>
struct FixedArray {
    Matrix<T, R, C> m;
    struct Mask {
        bool b[(R > 0 ? R : 1) * (C > 0 ? C : 1)];
        bool any() const { for (int k = 0; k < R * C; ++k) if (b[k]) return true; return false; }
        bool all() const { for (int k = 0; k < R * C; ++k) if (!b[k]) return false; return true; }
    };
    Mask operator<(const FixedArray& o) const { Mask r; for (int k = 0; k < R * C; ++k) r.b[k] = m(k) < o.m(k); return r; }
    Mask operator>(const FixedArray& o) const { Mask r; for (int k = 0; k < R * C; ++k) r.b[k] = m(k) > o.m(k); return r; }
    Mask operator<=(const FixedArray& o) const { Mask r; for (int k = 0; k < R * C; ++k) r.b[k] = m(k) <= o.m(k); return r; }
    Mask operator>=(const FixedArray& o) const { Mask r; for (int k = 0; k < R * C; ++k) r.b[k] = m(k) >= o.m(k); return r; }
    FixedArray min(const FixedArray& o) const { FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = o.m(k) < m(k) ? o.m(k) : m(k); return r; }
    FixedArray max(const FixedArray& o) const { FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = m(k) < o.m(k) ? o.m(k) : m(k); return r; }
    FixedArray abs() const { using std::abs; FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = abs(m(k)); return r; }
    FixedArray operator-(const FixedArray& o) const { FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = m(k) - o.m(k); return r; }
    FixedArray operator+(const FixedArray& o) const { FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = m(k) + o.m(k); return r; }
    FixedArray operator*(const FixedArray& o) const { FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = m(k) * o.m(k); return r; }
    FixedArray operator/(const FixedArray& o) const { FixedArray r; for (int k = 0; k < R * C; ++k) r.m(k) = m(k) / o.m(k); return r; }
    T maxCoeff() const { return m.maxCoeff(); }
    T minCoeff() const { return m.minCoeff(); }
    T sum() const { return m.sum(); }
    operator Matrix<T, R, C>() const { return m; }
    Matrix<T, R, C> matrix() const { return m; }
};
template <class Derived>
FixedArray<typename MatrixBase<Derived>::Scalar, MatrixBase<Derived>::RowsAtCompileTime, MatrixBase<Derived>::ColsAtCompileTime> MatrixBase<Derived>::array() const
{
    FixedArray<Scalar, RowsAtCompileTime, ColsAtCompileTime> a;
    a.m = eval();
    return a;
}

template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
    X& x_;
    int i0_, j0_;

public:
    typedef MatrixBase<Block> Base;
    typedef typename traits<X>::Scalar Scalar;
    Block(X& x, int i0, int j0) : x_(x), i0_(i0), j0_(j0) {}
    Block(const Block&) = default;
    Scalar& coeffRef(int i, int j) { return x_.coeffRef(i0_ + i, j0_ + j); }
    const Scalar& coeff(int i, int j) const { return const_cast<const X&>(x_).coeff(i0_ + i, j0_ + j); }
    Block& operator=(const Block& o) { Base::operator=(static_cast<const Base&>(o)); return *this; }
    template <class Od>
    Block& operator=(const MatrixBase<Od>& o) { return Base::operator=(o); }
};

template <class T, int N, int M>
class DiagonalMatrix {
public:
    Matrix<T, N, 1> d;
    DiagonalMatrix() {}
    template <class Od>
    explicit DiagonalMatrix(const MatrixBase<Od>& v) : d(v) {}
    DiagonalMatrix inverse() const
    {
        DiagonalMatrix r;
        for (int i = 0; i < N; ++i) r.d(i) = T(1) / d(i);
        return r;
    }
    operator Matrix<T, N, N>() const
    {
        Matrix<T, N, N> r;
        r.setZero();
        for (int i = 0; i < N; ++i) r(i, i) = d(i);
        return r;
    }
};

// ---- operators: all eager --------------------------------------------------------------------------------------------------
#define MINI_EIGEN_CWISE(OP)                                                                                                     \
    template <class A, class B>                                                                                                  \
    typename MatrixBase<A>::PlainObject operator OP(const MatrixBase<A>& a, const MatrixBase<B>& b)                               \
    {                                                                                                                            \
        static_assert((int)A::RowsAtCompileTime == (int)B::RowsAtCompileTime && (int)A::ColsAtCompileTime == (int)B::ColsAtCompileTime, "size"); \
        typename MatrixBase<A>::PlainObject r;                                                                                   \
        for (int j = 0; j < a.cols(); ++j)                                                                                       \
            for (int i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) OP b(i, j);                                                     \
        return r;                                                                                                                \
    }
MINI_EIGEN_CWISE(+)
MINI_EIGEN_CWISE(-)
#undef MINI_EIGEN_CWISE
template <class A>
typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A>& a)
{
    typename MatrixBase<A>::PlainObject r;
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) r(i, j) = -a(i, j);
    return r;
}
template <class A, class B>
Matrix<typename MatrixBase<A>::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    static_assert((int)A::ColsAtCompileTime == (int)B::RowsAtCompileTime, "product size");
    Matrix<typename MatrixBase<A>::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> r;
    for (int j = 0; j < b.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) {
            typename MatrixBase<A>::Scalar s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.cols(); ++k) s += a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
template <class A, class S, typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, const S& s)
{
    typename MatrixBase<A>::PlainObject r;
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) * (typename MatrixBase<A>::Scalar)s;
    return r;
}
template <class A, class S, typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>
typename MatrixBase<A>::PlainObject operator*(const S& s, const MatrixBase<A>& a)
{
    typename MatrixBase<A>::PlainObject r;
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) r(i, j) = (typename MatrixBase<A>::Scalar)s * a(i, j);
    return r;
}
template <class A, class S, typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, const S& s)
{
    typename MatrixBase<A>::PlainObject r;
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) / (typename MatrixBase<A>::Scalar)s;
    return r;
}
template <class A, class T, int N>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, const DiagonalMatrix<T, N>& d)
{
    typename MatrixBase<A>::PlainObject r;
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) * d.d(j);
    return r;
}
template <class A, class T, int N>
typename MatrixBase<A>::PlainObject operator*(const DiagonalMatrix<T, N>& d, const MatrixBase<A>& a)
{
    typename MatrixBase<A>::PlainObject r;
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) r(i, j) = d.d(i) * a(i, j);
    return r;
}
template <class A>
std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a)
{
    for (int i = 0; i < a.rows(); ++i) {
        for (int j = 0; j < a.cols(); ++j) os << a(i, j) << (j + 1 < a.cols() ? " " : "");
        os << "\n";
    }
    return os;
}

// symmetric eigen-decomposition by cyclic Jacobi; eigenvalues ascending like Eigen's
template <class M>
class SelfAdjointEigenSolver {
    typedef typename M::Scalar T;
    enum { N = M::RowsAtCompileTime };
    Matrix<T, N, 1> w_;
    M v_;

public:
    SelfAdjointEigenSolver() {}
    explicit SelfAdjointEigenSolver(const M& a) { compute(a); }
    SelfAdjointEigenSolver& compute(const M& a_in)
    {
        M a = a_in;
        v_.setIdentity();
        for (int sweep = 0; sweep < 64; ++sweep) {
            T off = 0;
            for (int p = 0; p < N; ++p)
                for (int q = p + 1; q < N; ++q) off += a(p, q) * a(p, q);
            if (off == T(0)) break;
            for (int p = 0; p < N; ++p)
                for (int q = p + 1; q < N; ++q) {
                    if (a(p, q) == T(0)) continue;
                    const T theta = (a(q, q) - a(p, p)) / (2 * a(p, q));
                    const T t = (theta >= 0 ? T(1) : T(-1)) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                    const T c = 1 / std::sqrt(t * t + 1), s = t * c;
                    for (int k = 0; k < N; ++k) { // A <- A G
                        const T akp = a(k, p), akq = a(k, q);
                        a(k, p) = c * akp - s * akq;
                        a(k, q) = s * akp + c * akq;
                    }
                    for (int k = 0; k < N; ++k) { // A <- G^T A
                        const T apk = a(p, k), aqk = a(q, k);
                        a(p, k) = c * apk - s * aqk;
                        a(q, k) = s * apk + c * aqk;
                    }
                    for (int k = 0; k < N; ++k) {
                        const T vkp = v_(k, p), vkq = v_(k, q);
                        v_(k, p) = c * vkp - s * vkq;
                        v_(k, q) = s * vkp + c * vkq;
                    }
                }
        }
        for (int i = 0; i < N; ++i) w_(i) = a(i, i);
        for (int i = 0; i < N; ++i) // ascending
            for (int j = i + 1; j < N; ++j)
                if (w_(j) < w_(i)) {
                    std::swap(w_(i), w_(j));
                    v_.col(i).swap(v_.col(j));
                }
        return *this;
    }
    const Matrix<T, N, 1>& eigenvalues() const { return w_; }
    const M& eigenvectors() const { return v_; }
};

typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 3, 1> Vector3f;

} // namespace Eigen
#include "mini_eigen_dyn.h"
#include "mini_eigen_geom.h"
