// ORACLE / TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the parts of Eigen 3 that the reference sources compiled
// into oracle/_ref/ (see oracle/ref/Makefile) name: fixed and dynamic dense matrices of scalars with element access,
// the comma initialiser, + - * /, norm / normalize, 3x3 inverse, cast, and a quaternion.  Every operation is the plain
// scalar expression in the order Eigen evaluates it for these sizes (element-wise; products as left-to-right dot
// products), so double arithmetic rounds like the reference binary's.  Not a product file; never shipped.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <memory>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

constexpr int Dynamic = -1;

template <class T>
using aligned_allocator = std::allocator<T>;

template <class T, int R, int C>
class Matrix;

template <class M>
class CommaInit {
public:
    CommaInit(M& m, typename M::Scalar v) : m_(m), i_(0) { put(v); }
    CommaInit& operator,(typename M::Scalar v) { put(v); return *this; }
    template <class T2, int R2, int C2>
    CommaInit& operator,(const Matrix<T2, R2, C2>& blk) {      // row-vector blocks are not used by the reference
        for (int r = 0; r < blk.rows(); r++)
            for (int c = 0; c < blk.cols(); c++) put((typename M::Scalar)blk(r, c));
        return *this;
    }
private:
    void put(typename M::Scalar v) {
        const int cols = m_.cols();
        m_(i_ / cols, i_ % cols) = v;
        i_++;
    }
    M& m_;
    int i_;
};

template <class T, int R, int C>
class Matrix {
public:
    typedef T Scalar;
    static constexpr bool kDyn = (R == Dynamic || C == Dynamic);

    Matrix() : rows_(R == Dynamic ? 0 : R), cols_(C == Dynamic ? 0 : C) { alloc(); }
    // two scalars: a 2-vector's coefficients, or the size of a dynamic matrix
    template <class A, class B>
    Matrix(const A& a, const B& b) : rows_(R == Dynamic ? (int)a : R), cols_(C == Dynamic ? (int)b : C) {
        alloc();
        if (!kDyn) { d_[0] = (T)a; d_[1] = (T)b; }
    }
    // same coefficients, other static shape (a dynamic block assigned to a fixed vector and the like)
    template <int R2, int C2, typename std::enable_if<R2 != R || C2 != C, int>::type = 0>
    Matrix(const Matrix<T, R2, C2>& o) : rows_(o.rows()), cols_(o.cols()) {
        alloc();
        for (int i = 0; i < rows_; i++) for (int j = 0; j < cols_; j++) (*this)(i, j) = o(i, j);
    }
    Matrix(const T& a, const T& b, const T& c) : rows_(R), cols_(C) { alloc(); d_[0] = a; d_[1] = b; d_[2] = c; }
    Matrix(const T& a, const T& b, const T& c, const T& d) : rows_(R), cols_(C) { alloc(); d_[0] = a; d_[1] = b; d_[2] = c; d_[3] = d; }

    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int size() const { return rows_ * cols_; }

    T& operator()(int i) { return d_[i]; }
    const T& operator()(int i) const { return d_[i]; }
    T& operator[](int i) { return d_[i]; }
    const T& operator[](int i) const { return d_[i]; }
    // column-major like Eigen's default (only matters for data())
    T& operator()(int r, int c) { return d_[(size_t)c * rows_ + r]; }
    const T& operator()(int r, int c) const { return d_[(size_t)c * rows_ + r]; }
    T& x() { return d_[0]; }
    T& y() { return d_[1]; }
    T& z() { return d_[2]; }
    const T& x() const { return d_[0]; }
    const T& y() const { return d_[1]; }
    const T& z() const { return d_[2]; }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }

    CommaInit<Matrix> operator<<(const T& v) { return CommaInit<Matrix>(*this, v); }

    static Matrix Zero() { Matrix m; for (auto& v : m.d_) v = T(0); return m; }
    static Matrix Identity() {
        Matrix m = Zero();
        for (int i = 0; i < m.rows_ && i < m.cols_; i++) m(i, i) = T(1);
        return m;
    }
    void setZero() { for (auto& v : d_) v = T(0); }

    Matrix operator+(const Matrix& o) const { Matrix r = *this; for (int i = 0; i < size(); i++) r.d_[i] = d_[i] + o.d_[i]; return r; }
    Matrix operator-(const Matrix& o) const { Matrix r = *this; for (int i = 0; i < size(); i++) r.d_[i] = d_[i] - o.d_[i]; return r; }
    Matrix operator-() const { Matrix r = *this; for (int i = 0; i < size(); i++) r.d_[i] = -d_[i]; return r; }
    Matrix operator*(const T& s) const { Matrix r = *this; for (int i = 0; i < size(); i++) r.d_[i] = d_[i] * s; return r; }
    Matrix operator/(const T& s) const { Matrix r = *this; for (int i = 0; i < size(); i++) r.d_[i] = d_[i] / s; return r; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < size(); i++) d_[i] += o.d_[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < size(); i++) d_[i] -= o.d_[i]; return *this; }
    Matrix& operator*=(const T& s) { for (int i = 0; i < size(); i++) d_[i] *= s; return *this; }
    Matrix& operator/=(const T& s) { for (int i = 0; i < size(); i++) d_[i] /= s; return *this; }

    template <int C2>
    Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& o) const {
        Matrix<T, R, C2> r = make<R, C2>(rows_, o.cols());
        for (int i = 0; i < rows_; i++)
            for (int j = 0; j < o.cols(); j++) {
                T acc = (*this)(i, 0) * o(0, j);
                for (int k = 1; k < cols_; k++) acc += (*this)(i, k) * o(k, j);
                r(i, j) = acc;
            }
        return r;
    }
    // dynamic * fixed vector (Camera::projectPoints: MatrixXd(3,3) * Vector3d)
    template <int R2, int C2, int RR = R, typename std::enable_if<RR == Dynamic && R2 != Dynamic, int>::type = 0>
    Matrix<T, R2, C2> operator*(const Matrix<T, R2, C2>& o) const {
        Matrix<T, R2, C2> r;
        for (int i = 0; i < rows_; i++)
            for (int j = 0; j < o.cols(); j++) {
                T acc = (*this)(i, 0) * o(0, j);
                for (int k = 1; k < cols_; k++) acc += (*this)(i, k) * o(k, j);
                r(i, j) = acc;
            }
        return r;
    }

    template <int N>
    Matrix<T, N, 1> head() const { Matrix<T, N, 1> r; for (int i = 0; i < N; i++) r(i) = d_[i]; return r; }
    template <int N>
    Matrix<T, N, 1> tail() const { Matrix<T, N, 1> r; for (int i = 0; i < N; i++) r(i) = d_[size() - N + i]; return r; }

    Matrix<T, Dynamic, C> topRows(int n) const {
        Matrix<T, Dynamic, C> r = Matrix<T, Dynamic, C>::sized(n, cols_);
        for (int i = 0; i < n; i++) for (int j = 0; j < cols_; j++) r(i, j) = (*this)(i, j);
        return r;
    }
    Matrix<T, Dynamic, C> bottomRows(int n) const {
        Matrix<T, Dynamic, C> r = Matrix<T, Dynamic, C>::sized(n, cols_);
        for (int i = 0; i < n; i++) for (int j = 0; j < cols_; j++) r(i, j) = (*this)(rows_ - n + i, j);
        return r;
    }
    template <int N>
    Matrix<T, N, C> topRows() const { Matrix<T, N, C> r; for (int i = 0; i < N; i++) for (int j = 0; j < cols_; j++) r(i, j) = (*this)(i, j); return r; }

    T squaredNorm() const { T s = d_[0] * d_[0]; for (int i = 1; i < size(); i++) s += d_[i] * d_[i]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const T n = norm(); if (n > T(0)) for (auto& v : d_) v /= n; }
    Matrix normalized() const { Matrix r = *this; r.normalize(); return r; }
    T dot(const Matrix& o) const { T s = d_[0] * o.d_[0]; for (int i = 1; i < size(); i++) s += d_[i] * o.d_[i]; return s; }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> r = make<C, R>(cols_, rows_);
        for (int i = 0; i < rows_; i++) for (int j = 0; j < cols_; j++) r(j, i) = (*this)(i, j);
        return r;
    }
    Matrix inverse() const {       // 2x2 / 3x3 by cofactors (only used off the parity path: initUndistortRectifyMap)
        Matrix r = *this;
        const Matrix& m = *this;
        if (rows_ == 2) {
            const T det = m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
            r(0, 0) = m(1, 1) / det; r(0, 1) = -m(0, 1) / det; r(1, 0) = -m(1, 0) / det; r(1, 1) = m(0, 0) / det;
        } else {
            const T c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1), c01 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2),
                    c02 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
            const T det = m(0, 0) * c00 + m(0, 1) * c01 + m(0, 2) * c02;
            r(0, 0) = c00 / det; r(1, 0) = c01 / det; r(2, 0) = c02 / det;
            r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) / det;
            r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) / det;
            r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) / det;
            r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) / det;
            r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) / det;
            r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) / det;
        }
        return r;
    }
    template <class U>
    Matrix<U, R, C> cast() const {
        Matrix<U, R, C> r = Matrix<U, R, C>::sized(rows_, cols_);
        for (int i = 0; i < rows_; i++) for (int j = 0; j < cols_; j++) r(i, j) = (U)(*this)(i, j);
        return r;
    }
    static Matrix sized(int r, int c) { Matrix m; m.rows_ = r; m.cols_ = c; m.alloc(); return m; }

private:
    template <int R2, int C2>
    static Matrix<T, R2, C2> make(int r, int c) { return Matrix<T, R2, C2>::sized(r, c); }
    void alloc() { d_.assign((size_t)rows_ * cols_, T(0)); }
    int rows_, cols_;
    std::vector<T> d_;
};

template <class T, int R, int C>
Matrix<T, R, C> operator*(const T& s, const Matrix<T, R, C>& m) { return m * s; }

template <class T>
class Quaternion {
public:
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
    T& w() { return w_; } T& x() { return x_; } T& y() { return y_; } T& z() { return z_; }
    const T& w() const { return w_; } const T& x() const { return x_; } const T& y() const { return y_; } const T& z() const { return z_; }
    Matrix<T, 3, 3> toRotationMatrix() const {
        Matrix<T, 3, 3> m;
        const T tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
        const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_,
                tyz = tz * y_, tzz = tz * z_;
        m(0, 0) = 1 - (tyy + tzz); m(0, 1) = txy - twz; m(0, 2) = txz + twy;
        m(1, 0) = txy + twz; m(1, 1) = 1 - (txx + tzz); m(1, 2) = tyz - twx;
        m(2, 0) = txz - twy; m(2, 1) = tyz + twx; m(2, 2) = 1 - (txx + tyy);
        return m;
    }
private:
    T w_, x_, y_, z_;
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

}  // namespace Eigen
