// TEST INFRASTRUCTURE ONLY - the parts of Eigen/Geometry and of Eigen's scalar-traits machinery that the reference's collision-object code names
// (Lib/Ziran/Math/Geometry/{Rotation.h, CollisionObject.*, AnalyticLevelSet.*}, Lib/Ziran/Math/Nonlinear/AutoDiff.h): Quaternion (w, x, y, z) with
// Eigen's formulas (Eigen/src/Geometry/Quaternion.h: toRotationMatrix, normalized, FromTwoVectors for non-antiparallel vectors), Rotation2D,
// NumTraits / IOFormat / internal::cast as declarations.  Third-party arithmetic restated, like the rest of this stand-in.
#pragma once
namespace Eigen {
template <class T>
struct NumTraits {
    typedef T Real;
    typedef T NonInteger;
    typedef T Nested;
    enum { IsComplex = 0, IsInteger = 0, IsSigned = 1, RequireInitialization = 0, ReadCost = 1, AddCost = 1, MulCost = 1 };
};
struct IOFormat {
    IOFormat(int = 0, int = 0, const char* = "", const char* = "", const char* = "", const char* = "") {}
};
namespace internal {
template <class A, class B>
inline B cast(const A& a) { return static_cast<B>(a); }
} // namespace internal

template <class T>
class Quaternion {
    T w_, x_, y_, z_;

public:
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(const T& w, const T& x, const T& y, const T& z) : w_(w), x_(x), y_(y), z_(z) {}
    explicit Quaternion(const Matrix<T, 4, 1>& v) : w_(v(3)), x_(v(0)), y_(v(1)), z_(v(2)) {} // Eigen: coefficients are stored (x, y, z, w)
    T w() const { return w_; }
    T x() const { return x_; }
    T y() const { return y_; }
    T z() const { return z_; }
    T norm() const { return std::sqrt(w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_); }
    Quaternion normalized() const
    {
        const T n = norm();
        return Quaternion(w_ / n, x_ / n, y_ / n, z_ / n);
    }
    void normalize() { *this = normalized(); }
    Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
    Quaternion inverse() const
    {
        const T n2 = w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_;
        return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
    }
    Quaternion operator*(const Quaternion& b) const
    {
        const Quaternion& a = *this;
        return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
            a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
    }
    Matrix<T, 3, 3> toRotationMatrix() const
    {
        Matrix<T, 3, 3> res;
        const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
        const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
        const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
        const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
        res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
        return res;
    }
    Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const { return toRotationMatrix() * v; }
    Matrix<T, 3, 1> _transformVector(const Matrix<T, 3, 1>& v) const { return toRotationMatrix() * v; }
    operator Matrix<T, 3, 3>() const { return toRotationMatrix(); }
    template <class A, class B>
    Quaternion& setFromTwoVectors(const MatrixBase<A>& a, const MatrixBase<B>& b) { *this = FromTwoVectors(a, b); return *this; }
    template <class A, class B>
    static Quaternion FromTwoVectors(const MatrixBase<A>& a, const MatrixBase<B>& b)
    { // Eigen's setFromTwoVectors away from the antiparallel case: axis = v0 x v1, s = sqrt(2 (1 + v0 . v1)), vec = axis / s, w = s / 2
        Matrix<T, 3, 1> v0 = a.normalized(), v1 = b.normalized();
        const T c = v1.dot(v0);
        if (c < T(-1) + T(1e-12)) {
            // antiparallel: Eigen takes the rotation axis from a JacobiSVD of [v0; v1] (any unit vector orthogonal to v0 is a valid answer, which one it
            // returns is a property of that SVD).  Here: the axis orthogonal to v0 closest to e_z (rotation by pi); results on this branch are not
            // compared with anything (tests/test_collider_ref.py excludes them)
            Matrix<T, 3, 1> ez = Matrix<T, 3, 1>::Zero();
            ez(2) = T(1);
            Matrix<T, 3, 1> ax = ez - v0 * v0.dot(ez);
            if (ax.norm() < T(1e-6)) { ax = Matrix<T, 3, 1>::Zero(); ax(1) = T(1); ax = ax - v0 * v0.dot(ax); }
            ax = ax.normalized();
            return Quaternion(T(0), ax(0), ax(1), ax(2));
        }
        Matrix<T, 3, 1> axis = v0.cross(v1);
        const T s = std::sqrt((T(1) + c) * T(2));
        const T invs = T(1) / s;
        return Quaternion(s * T(0.5), axis(0) * invs, axis(1) * invs, axis(2) * invs);
    }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class T>
class Rotation2D {
    T a_;

public:
    explicit Rotation2D(const T& a = T(0)) : a_(a) {}
    T angle() const { return a_; }
    Matrix<T, 2, 2> toRotationMatrix() const
    {
        const T s = std::sin(a_), c = std::cos(a_);
        Matrix<T, 2, 2> r;
        r(0, 0) = c; r(0, 1) = -s; r(1, 0) = s; r(1, 1) = c;
        return r;
    }
};
typedef Matrix<int, 4, 1> Vector4i;
} // namespace Eigen
