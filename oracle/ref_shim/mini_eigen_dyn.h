// TEST INFRASTRUCTURE ONLY - dynamic-size part of the Eigen stand-in (see mini_eigen.h): exactly what the reference's multigrid code
// (Projects/multigrid/{SquareMatrix.h,MPMMultigridMatrix.h,MultigridPreconditioner.h}) uses of
//   Eigen::Matrix<T, R, Dynamic>  ("TVStack": R x n, column-major, one column per grid node)  and  Eigen::Matrix<T, Dynamic, 1>  ("Vec"),
// all eager, plus inert definitions of the sparse-solver classes those headers hold as members but the pinned paths never call.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <vector>

namespace Eigen {

// a column of a dynamic-column matrix: a fixed-size R x 1 lvalue, so every fixed-size operator of mini_eigen.h applies to it
template <class T, int R>
class DynColRef;
template <class T, int R>
struct traits<DynColRef<T, R>> {
    typedef T Scalar;
    enum { Rows = R, Cols = 1 };
};
template <class T, int R>
class DynColRef : public MatrixBase<DynColRef<T, R>> {
    T* p_;

public:
    typedef MatrixBase<DynColRef> Base;
    typedef T Scalar;
    explicit DynColRef(T* p) : p_(p) {}
    DynColRef(const DynColRef&) = default;
    T& coeffRef(int i, int) { return p_[i]; }
    const T& coeff(int i, int) const { return p_[i]; }
    DynColRef& operator=(const DynColRef& o)
    {
        T t[R];
        for (int i = 0; i < R; ++i) t[i] = o.p_[i];
        for (int i = 0; i < R; ++i) p_[i] = t[i];
        return *this;
    }
    template <class Od>
    DynColRef& operator=(const MatrixBase<Od>& o) { return Base::operator=(o); }
};

struct DynSum { // result of .array() chains that end in .sum()
    double s;
    double sum() const { return s; }
    DynSum array() const { return *this; }
};
template <class T>
struct DynArray { // .array() of a dynamic matrix / vector: coefficient-wise product, .sum(), .unaryExpr is not needed
    std::vector<T> v;
    T sum() const { T s = 0; for (const T& x : v) s += x; return s; }
    DynArray abs() const { DynArray r; r.v.resize(v.size()); for (size_t i = 0; i < v.size(); ++i) r.v[i] = std::abs(v[i]); return r; }
    T maxCoeff() const { T m = v.empty() ? T(0) : v[0]; for (const T& x : v) if (x > m) m = x; return m; }
};
template <class T>
DynArray<T> operator*(const DynArray<T>& a, const DynArray<T>& b)
{
    DynArray<T> r; r.v.resize(a.v.size());
    for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = a.v[i] * b.v[i];
    return r;
}

// Vec::segment(start, n), its transpose, (fixed column) * (segment transposed) and (fixed row) * middleCols  (ImplicitSolver.h:138,268-271)
template <class T>
struct DynSegment {
    const T* p; int n;
    struct Transposed { const T* p; int n; };
    Transposed transpose() const { return Transposed{p, n}; }
};
template <class T>
struct DynRow { // (1 x n) result of rowvector * middleCols
    std::vector<T> v;
    T dot(const DynSegment<T>& s) const { T r = 0; for (int j = 0; j < s.n; ++j) r += v[j] * s.p[j]; return r; }
};
template <class T, int R, int O, int MR, int MC>
class Matrix<T, R, Dynamic, O, MR, MC> {
    std::vector<T> m_;
    int c_ = 0;

public:
    typedef T Scalar;
    typedef std::ptrdiff_t Index;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = Dynamic, SizeAtCompileTime = Dynamic };
    Matrix() {}
    Matrix(int r, int c) { resize(r, c); }
    void resize(int, int c) { c_ = c; m_.resize((size_t)R * c); }
    template <class M> void resizeLike(const M& o) { resize(R, (int)o.cols()); }
    int rows() const { return R; }
    int cols() const { return c_; }
    long size() const { return (long)m_.size(); }
    T* data() { return m_.data(); }
    const T* data() const { return m_.data(); }
    Matrix& setZero() { for (T& x : m_) x = T(0); return *this; }
    Matrix& setZero(int r, int c) { resize(r, c); return setZero(); }
    DynColRef<T, R> col(int j) const { return DynColRef<T, R>(const_cast<T*>(m_.data()) + (size_t)R * j); }
    T& operator()(int i, int j) { return m_[(size_t)R * j + i]; }
    const T& operator()(int i, int j) const { return m_[(size_t)R * j + i]; }
    Matrix& operator+=(const Matrix& o) { for (size_t k = 0; k < m_.size(); ++k) m_[k] += o.m_[k]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (size_t k = 0; k < m_.size(); ++k) m_[k] -= o.m_[k]; return *this; }
    Matrix& operator*=(const T& s) { for (T& x : m_) x *= s; return *this; }
    Matrix& operator/=(const T& s) { for (T& x : m_) x /= s; return *this; }
    DynArray<T> array() const { return DynArray<T>{m_}; }
    T squaredNorm() const { T s = 0; for (const T& x : m_) s += x * x; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix& noalias() { return *this; }
    void swap(Matrix& o) { m_.swap(o.m_); std::swap(c_, o.c_); }
    Matrix& operator=(const DynArray<T>& a) { m_ = a.v; return *this; }          // x = x.array().abs()
    Matrix& setRandom() { for (T& x : m_) x = T(2) * T(rand()) / T(RAND_MAX) - T(1); return *this; }
    // r.middleCols(start, n).colwise().squaredNorm().array().sum()  (MultigridPreconditioner.h:130-141)
    struct Middle {
        const Matrix* m; int start, n;
        struct Colwise {
            const Middle* b;
            DynSum squaredNorm() const
            {
                double s = 0;
                for (int j = b->start; j < b->start + b->n; ++j) {
                    double cs = 0; // (one column at a time, like colwise())
                    for (int i = 0; i < R; ++i) cs += (double)((*b->m)(i, j) * (*b->m)(i, j));
                    s += cs;
                }
                return DynSum{s};
            }
        };
        Colwise colwise() const { return Colwise{this}; }
        // residual.middleCols(start, n) = dtg * mass.segment(start, n).transpose()  (ImplicitSolver.h:138): column j = dtg * mass(start + j)
        template <int O2, int MR2, int MC2>
        friend DynRow<T> operator*(const Matrix<T, 1, R, O2, MR2, MC2>& g, const Middle& b) // gravity.transpose() * dv.middleCols(..)
        {
            DynRow<T> r;
            r.v.resize(b.n);
            for (int j = 0; j < b.n; ++j) {
                T s = 0;
                for (int i = 0; i < R; ++i) s += g(0, i) * (*b.m)(i, b.start + j);
                r.v[j] = s;
            }
            return r;
        }
        // ((a.middleCols(..).transpose() * b.middleCols(..)).diagonal().array()).sum()  (ImplicitSolver.h:228): sum over the columns of a_j . b_j
        struct MiddleT {
            const Middle* a;
            struct Prod {
                const Middle *a, *b;
                const Prod& diagonal() const { return *this; }
                const Prod& array() const { return *this; }
                T sum() const
                {
                    T s = 0;
                    for (int j = 0; j < a->n; ++j) {
                        T d = 0;
                        for (int i = 0; i < R; ++i) d += (*a->m)(i, a->start + j) * (*b->m)(i, b->start + j);
                        s += d;
                    }
                    return s;
                }
            };
            Prod operator*(const Middle& b) const { return Prod{a, &b}; }
        };
        MiddleT transpose() const { return MiddleT{this}; }
        template <class OP>
        const Middle& operator=(const OP& op) const
        {
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < R; ++i) const_cast<Matrix&>(*m)(i, start + j) = op.v[i] * op.p[j];
            return *this;
        }
    };
    Middle middleCols(int start, int n) const { return Middle{this, start, n}; }
};
template <class T, int R>
struct DynOuter { T v[R]; const T* p; int n; };
template <class T, int R, int O, int MR, int MC>
DynOuter<T, R> operator*(const Matrix<T, R, 1, O, MR, MC>& v, const typename DynSegment<T>::Transposed& s)
{
    DynOuter<T, R> r;
    for (int i = 0; i < R; ++i) r.v[i] = v(i);
    r.p = s.p; r.n = s.n;
    return r;
}
template <class T, int R, int O, int MR, int MC, class S, typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>
Matrix<T, R, Dynamic, O, MR, MC> operator/(const Matrix<T, R, Dynamic, O, MR, MC>& a, const S& s) { Matrix<T, R, Dynamic, O, MR, MC> r = a; r /= (T)s; return r; }
template <class T, int R, int O, int MR, int MC, class S, typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>
Matrix<T, R, Dynamic, O, MR, MC> operator*(const Matrix<T, R, Dynamic, O, MR, MC>& a, const S& s) { Matrix<T, R, Dynamic, O, MR, MC> r = a; r *= (T)s; return r; }
template <class T, int R, int O, int MR, int MC, class S, typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>
Matrix<T, R, Dynamic, O, MR, MC> operator*(const S& s, const Matrix<T, R, Dynamic, O, MR, MC>& a) { return a * s; }
template <class T, int R, int O, int MR, int MC>
Matrix<T, R, Dynamic, O, MR, MC> operator+(const Matrix<T, R, Dynamic, O, MR, MC>& a, const Matrix<T, R, Dynamic, O, MR, MC>& b) { Matrix<T, R, Dynamic, O, MR, MC> r = a; r += b; return r; }
template <class T, int R, int O, int MR, int MC>
Matrix<T, R, Dynamic, O, MR, MC> operator-(const Matrix<T, R, Dynamic, O, MR, MC>& a, const Matrix<T, R, Dynamic, O, MR, MC>& b) { Matrix<T, R, Dynamic, O, MR, MC> r = a; r -= b; return r; }

template <class T, int O, int MR, int MC>
class Matrix<T, Dynamic, 1, O, MR, MC> {
    std::vector<T> m_;

public:
    typedef T Scalar;
    typedef std::ptrdiff_t Index;
    enum { RowsAtCompileTime = Dynamic, ColsAtCompileTime = 1, SizeAtCompileTime = Dynamic };
    Matrix() {}
    explicit Matrix(int n) { resize(n); }
    void resize(int n) { m_.resize((size_t)n); }
    void resize(int n, int) { m_.resize((size_t)n); }
    template <class M> void resizeLike(const M& o) { resize((int)o.size()); }
    int rows() const { return (int)m_.size(); }
    int cols() const { return 1; }
    long size() const { return (long)m_.size(); }
    T* data() { return m_.data(); }
    const T* data() const { return m_.data(); }
    Matrix& setZero() { for (T& x : m_) x = T(0); return *this; }
    Matrix& setZero(int n) { resize(n); return setZero(); }
    Matrix& setOnes() { for (T& x : m_) x = T(1); return *this; }
    T& operator()(int i) { return m_[i]; }
    const T& operator()(int i) const { return m_[i]; }
    T& operator[](int i) { return m_[i]; }
    const T& operator[](int i) const { return m_[i]; }
    Matrix& operator+=(const Matrix& o) { for (size_t k = 0; k < m_.size(); ++k) m_[k] += o.m_[k]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (size_t k = 0; k < m_.size(); ++k) m_[k] -= o.m_[k]; return *this; }
    Matrix& operator*=(const T& s) { for (T& x : m_) x *= s; return *this; }
    Matrix& operator/=(const T& s) { for (T& x : m_) x /= s; return *this; }
    DynArray<T> array() const { return DynArray<T>{m_}; }
    T squaredNorm() const { T s = 0; for (const T& x : m_) s += x * x; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    T dot(const Matrix& o) const { T s = 0; for (size_t k = 0; k < m_.size(); ++k) s += m_[k] * o.m_[k]; return s; }
    DynSegment<T> segment(int start, int n) const { return DynSegment<T>{m_.data() + start, n}; }
};

// Map of a dynamic vector (EIGEN_EXT::vec() of DenseExt.h views a TVStack as one long vector; used by direct-solver paths only)
template <class T, int O, int MR, int MC>
class Map<Matrix<T, Dynamic, 1, O, MR, MC>> {
    T* p_; long n_;
public:
    Map(T* p, long n, long = 1) : p_(p), n_(n) {}
    long size() const { return n_; }
    T& operator()(long i) { return p_[i]; }
    const T& operator()(long i) const { return p_[i]; }
    Map& noalias() { return *this; }
    template <class X> Map& operator=(const X&) { return *this; }
};
template <class T, int O, int MR, int MC>
class Map<const Matrix<T, Dynamic, 1, O, MR, MC>> {
    const T* p_; long n_;
public:
    Map(const T* p, long n, long = 1) : p_(p), n_(n) {}
    long size() const { return n_; }
    const T& operator()(long i) const { return p_[i]; }
};

// ---- sparse matrix / triplet: inert (held as a member by SquareMatrix, filled only by debugging / direct-solver paths that are not pinned)
template <class T, class I>
class Triplet {
public:
    Triplet() {}
    Triplet(int, int, const T&) {}
};
template <class T, int Options, class I>
class SparseMatrix {
public:
    struct InnerIterator {
        InnerIterator(const SparseMatrix&, int) {}
        operator bool() const { return false; }
        InnerIterator& operator++() { return *this; }
        int row() const { return 0; }
        int col() const { return 0; }
        T value() const { return T(0); }
    };
    void resize(int, int) {}
    template <class It> void setFromTriplets(It, It) {}
    int outerSize() const { return 0; }
    int rows() const { return 0; }
    int cols() const { return 0; }
    T coeffRef(int, int) { return T(0); }
    void makeCompressed() {}
    SparseMatrix transpose() const { return *this; }
    template <class V> V operator*(const V& v) const { return v; }
    Matrix<T, Dynamic, 1> diagonal() const { return Matrix<T, Dynamic, 1>(); }
};

// ---- sparse solvers the reference's SquareMatrix holds as members / names in code paths that are not pinned: inert
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
template <class T, int UpLo = 1, class Ordering = int>
class IncompleteCholesky {
public:
    template <class M> void compute(const M&) {}
    template <class M> IncompleteCholesky& analyzePattern(const M&) { return *this; }
    template <class M> IncompleteCholesky& factorize(const M&) { return *this; }
    template <class V> V solve(const V& b) const { return b; }
    ComputationInfo info() const { return Success; }
    void setInitialShift(double) {}
};
template <class M, int UpLo = 1, class Ordering = int>
class SimplicialLDLT {
public:
    SimplicialLDLT() {}
    template <class X> explicit SimplicialLDLT(const X&) {}
    template <class X> void compute(const X&) {}
    template <class V> V solve(const V& b) const { return b; }
    ComputationInfo info() const { return Success; }
};
template <class M, int UpLo = 1, class Ordering = int>
class SimplicialLLT : public SimplicialLDLT<M, UpLo, Ordering> {
public:
    SimplicialLLT() {}
    template <class X> explicit SimplicialLLT(const X&) {}
};
template <class M, class Solver = int, bool B = false>
class ArpackGeneralizedSelfAdjointEigenSolver {
public:
    template <class... A> ArpackGeneralizedSelfAdjointEigenSolver(const A&...) {}
    template <class... A> ArpackGeneralizedSelfAdjointEigenSolver& compute(const A&...) { return *this; }
    Matrix<double, Dynamic, 1> eigenvalues() const { return Matrix<double, Dynamic, 1>(1); }
    ComputationInfo info() const { return Success; }
};

} // namespace Eigen
