// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY, never linked into the product.
//
// The reference's src/wass_stereo/PovMesh.{h,cpp} and src/wass_lib/triangulate.hpp include <opencv2/opencv.hpp> and
// <boost/filesystem.hpp>; neither library's C++ headers exist in this image.  This header provides, in namespace cv,
// exactly the small-matrix vocabulary those two files use -- Vec, Matx, Point3_, a dense Mat / Mat_<T>, solve, SVD,
// norm, normalize, determinant, and no-op imwrite / resize -- so that the UNMODIFIED reference sources compile where
// they lie (oracle/build_ref.sh) into oracle/_ref/povmesh_ref, the checker that pins oracle/pipeline.py.
//
// What is OpenCV's own arithmetic and what is not:
//   * Vec / Matx operators, cross, dot, norm, normalize: the textbook left-to-right formulas OpenCV's templates expand to.
//   * cv::solve(3x3, DECOMP_LU): OpenCV does not run LU for n <= 3 but Cramer's rule on the cofactors with one
//     reciprocal of the determinant; restated here in that operation order (the repo's own restatement of the same
//     formula, oracle/pipeline.py::solve3_lu, is pinned to cv2.solve at rtol 1e-11 in tests/test_oracle_pipeline.py).
//   * cv::SVD of the symmetric 3x3 scatter matrix in PovMesh::refine_plane: a cyclic Jacobi eigen-solver; the reference
//     only takes the singular vector of the smallest singular value, normalises it and fixes its sign, so any
//     converged solver gives the same plane to rounding (compared at 1e-9).
//   * DECOMP_SVD least squares (only in triangulation variants the reference compiles out, ITERATIVE_TRIANGULATION OFF,
//     wass_stereo.cpp:100): normal equations, enough to compile and run.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_16SC1 3
#define CV_64FC1 6
#define CV_64F 6

namespace cv {

enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_CHOLESKY = 3 };

// ---------------------------------------------------------------------------------------------- Matx / Vec
template <typename T, int M, int N> struct Matx {
    T val[M * N];
    Matx() { for (int i = 0; i < M * N; ++i) val[i] = T(0); }
    template <typename... A, typename = typename std::enable_if<sizeof...(A) == M * N && (M * N > 1)>::type>
    Matx(A... a) { T tmp[] = {T(a)...}; for (int i = 0; i < M * N; ++i) val[i] = tmp[i]; }
    static Matx zeros() { return Matx(); }
    static Matx eye() { Matx m; for (int i = 0; i < (M < N ? M : N); ++i) m(i, i) = T(1); return m; }
    T& operator()(int i, int j) { return val[i * N + j]; }
    const T& operator()(int i, int j) const { return val[i * N + j]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    Matx<T, N, M> t() const { Matx<T, N, M> r; for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) r(j, i) = (*this)(i, j); return r; }
};

template <typename T, int N> struct Vec : Matx<T, N, 1> {
    Vec() {}
    Vec(const Matx<T, N, 1>& m) { for (int i = 0; i < N; ++i) this->val[i] = m.val[i]; }
    template <typename... A, typename = typename std::enable_if<sizeof...(A) == N && (N > 1)>::type>
    Vec(A... a) { T tmp[] = {T(a)...}; for (int i = 0; i < N; ++i) this->val[i] = tmp[i]; }
    T& operator[](int i) { return this->val[i]; }
    const T& operator[](int i) const { return this->val[i]; }
    template <typename U, typename = typename std::enable_if<!std::is_same<U, T>::value>::type>
    operator Vec<U, N>() const { Vec<U, N> r; for (int i = 0; i < N; ++i) r[i] = (U)this->val[i]; return r; }
    Vec cross(const Vec& o) const
    {
        static_assert(N == 3, "cross");
        return Vec(this->val[1] * o.val[2] - this->val[2] * o.val[1], this->val[2] * o.val[0] - this->val[0] * o.val[2],
                   this->val[0] * o.val[1] - this->val[1] * o.val[0]);
    }
    double ddot(const Vec& o) const { double s = 0; for (int i = 0; i < N; ++i) s += (double)this->val[i] * o.val[i]; return s; }
    T dot(const Vec& o) const { T s = 0; for (int i = 0; i < N; ++i) s += this->val[i] * o.val[i]; return s; }
};

template <typename T, int N> Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }
template <typename T, int N> Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }
template <typename T, int N> Vec<T, N> operator-(const Vec<T, N>& a) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }
template <typename T, int N> Vec<T, N> operator*(const Vec<T, N>& a, double s) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = T(a[i] * s); return r; }
template <typename T, int N> Vec<T, N> operator*(double s, const Vec<T, N>& a) { return a * s; }
template <typename T, int N> Vec<T, N> operator/(const Vec<T, N>& a, double s) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = T(a[i] / s); return r; }
template <typename T, int N> Vec<T, N>& operator+=(Vec<T, N>& a, const Vec<T, N>& b) { for (int i = 0; i < N; ++i) a[i] += b[i]; return a; }
template <typename T, int N> Vec<T, N>& operator-=(Vec<T, N>& a, const Vec<T, N>& b) { for (int i = 0; i < N; ++i) a[i] -= b[i]; return a; }
template <typename T, int N> Vec<T, N>& operator*=(Vec<T, N>& a, double s) { for (int i = 0; i < N; ++i) a[i] = T(a[i] * s); return a; }
template <typename T, int N> Vec<T, N>& operator/=(Vec<T, N>& a, double s) { for (int i = 0; i < N; ++i) a[i] = T(a[i] / s); return a; }

template <typename T, int M, int N> Matx<T, M, N> operator+(const Matx<T, M, N>& a, const Matx<T, M, N>& b) { Matx<T, M, N> r; for (int i = 0; i < M * N; ++i) r.val[i] = a.val[i] + b.val[i]; return r; }
template <typename T, int M, int N> Matx<T, M, N> operator-(const Matx<T, M, N>& a, const Matx<T, M, N>& b) { Matx<T, M, N> r; for (int i = 0; i < M * N; ++i) r.val[i] = a.val[i] - b.val[i]; return r; }
template <typename T, int M, int N> Matx<T, M, N> operator-(const Matx<T, M, N>& a) { Matx<T, M, N> r; for (int i = 0; i < M * N; ++i) r.val[i] = -a.val[i]; return r; }
template <typename T, int M, int N> Matx<T, M, N> operator*(const Matx<T, M, N>& a, double s) { Matx<T, M, N> r; for (int i = 0; i < M * N; ++i) r.val[i] = T(a.val[i] * s); return r; }
template <typename T, int M, int N> Matx<T, M, N> operator*(double s, const Matx<T, M, N>& a) { return a * s; }

template <typename T, int M, int N> Vec<T, M> operator*(const Matx<T, M, N>& A, const Vec<T, N>& x)
{
    Vec<T, M> r;
    for (int i = 0; i < M; ++i) { T s = 0; for (int j = 0; j < N; ++j) s += A(i, j) * x[j]; r[i] = s; }
    return r;
}
template <typename T, int M, int N, int L> Matx<T, M, L> operator*(const Matx<T, M, N>& A, const Matx<T, N, L>& B)
{
    Matx<T, M, L> r;
    for (int i = 0; i < M; ++i) for (int j = 0; j < L; ++j) { T s = 0; for (int k = 0; k < N; ++k) s += A(i, k) * B(k, j); r(i, j) = s; }
    return r;
}

template <typename T, int N> double norm(const Vec<T, N>& v) { double s = 0; for (int i = 0; i < N; ++i) s += (double)v[i] * (double)v[i]; return std::sqrt(s); }
template <typename T, int N> Vec<T, N> normalize(const Vec<T, N>& v) { const double n = norm(v); return v * (n == 0 ? 0.0 : 1.0 / n); }
template <typename T> double determinant(const Matx<T, 3, 3>& m)
{
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}

typedef Vec<double, 2> Vec2d;
typedef Vec<float, 2> Vec2f;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;
typedef Matx<double, 4, 1> Matx41d;
typedef Vec<int, 2> Vec2i;
typedef Vec<unsigned char, 3> Vec3b;
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<double, 3, 4> Matx34d;
typedef Matx<double, 4, 3> Matx43d;

template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<double> Point3d;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

template <typename T, int N> std::ostream& operator<<(std::ostream& o, const Vec<T, N>& v)
{
    o << "[";
    for (int i = 0; i < N; ++i) o << (i ? ", " : "") << +v[i];
    return o << "]";
}

// ---------------------------------------------------------------------------------------------- Mat
inline int shim_elem_size(int type) { return type == CV_64FC1 ? 8 : type == CV_32FC1 ? 4 : type == CV_8UC3 ? 3 : type == CV_16SC1 ? 2 : 1; }

class Mat {
public:
    int rows = 0, cols = 0, mtype = CV_64FC1;
    unsigned char* data = nullptr;       // view into `own` or into user memory (the (rows, cols, type, ptr) constructor)
    std::vector<unsigned char> own;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* user) : rows(r), cols(c), mtype(type), data((unsigned char*)user) {}
    Mat(const Mat& o) { *this = o; }
    template <typename T, int M, int N> Mat(const Matx<T, M, N>& m)
    {
        create(M, N, CV_64FC1);
        for (int i = 0; i < M * N; ++i) ((double*)data)[i] = (double)m.val[i];
    }
    Mat& operator=(const Mat& o)
    {
        if (this == &o) return *this;
        rows = o.rows; cols = o.cols; mtype = o.mtype;
        if (!o.own.empty() || o.data == nullptr) { own = o.own; data = own.empty() ? nullptr : own.data(); }
        else { own.clear(); data = o.data; }                     // header over user memory: shallow, like OpenCV
        return *this;
    }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; mtype = type;
        own.assign((size_t)r * c * shim_elem_size(type), 0);
        data = own.data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    Mat clone() const { Mat o(rows, cols, mtype); if (data) memcpy(o.data, data, o.own.size()); return o; }
    Mat t() const { Mat o(cols, rows, CV_64FC1); for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) o.at<double>(j, i) = at<double>(i, j); return o; }
    Mat mul(const Mat& b) const      // element-wise product of 8-bit images, saturated like cv::Mat::mul
    {
        Mat o(rows, cols, mtype);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) { const int v = (int)data[i] * (int)b.data[i]; o.data[i] = (unsigned char)(v > 255 ? 255 : v); }
        return o;
    }
    template <typename T, int M, int N, typename = typename std::enable_if<(N > 1)>::type> operator Matx<T, M, N>() const
    {
        Matx<T, M, N> r;
        for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) r(i, j) = (T)at<double>(i, j);
        return r;
    }
    template <typename T, int N> operator Vec<T, N>() const { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)((const double*)data)[i]; return r; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    int type() const { return mtype; }
    template <typename T> T& at(int i, int j) { return *(T*)(data + ((size_t)i * cols + j) * sizeof(T)); }
    template <typename T> const T& at(int i, int j) const { return *(const T*)(data + ((size_t)i * cols + j) * sizeof(T)); }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * cols * shim_elem_size(mtype); }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * cols * shim_elem_size(mtype); }
};

inline std::ostream& operator<<(std::ostream& o, const Mat& m)
{
    o << "[";
    for (int i = 0; i < m.rows; ++i) {
        for (int j = 0; j < m.cols; ++j) o << (j ? ", " : "") << (m.mtype == CV_64FC1 ? m.at<double>(i, j) : 0.0);
        o << (i + 1 < m.rows ? ";\n " : "");
    }
    return o << "]";
}

template <typename T> class Mat_;
template <typename T> struct MatCommaInitializer_ {
    Mat_<T>* m; int idx;
    template <typename U> MatCommaInitializer_& operator,(U v);
    operator Mat_<T>() const;
    operator Mat() const;
};

template <typename T> class Mat_ : public Mat {
public:
    Mat_() { mtype = CV_64FC1; }
    Mat_(int r, int c) : Mat(r, c, CV_64FC1) { static_assert(std::is_same<T, double>::value, "shim: Mat_<double> only"); }
    Mat_(const Mat& m) : Mat(m) {}
    template <int M, int N> Mat_(const Matx<T, M, N>& m) : Mat(m) {}
    T& operator()(int i, int j) { return this->template at<T>(i, j); }
    const T& operator()(int i, int j) const { return this->template at<T>(i, j); }
    T& operator()(int i) { return ((T*)data)[i]; }
    const T& operator()(int i) const { return ((const T*)data)[i]; }
    Mat_ row(int r) const { Mat_ o(1, cols); for (int j = 0; j < cols; ++j) o(0, j) = (*this)(r, j); return o; }
    template <typename U> MatCommaInitializer_<T> operator<<(U v) { (*this)(0) = (T)v; return MatCommaInitializer_<T>{this, 1}; }
};
template <typename T> template <typename U> MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(U v) { (*m)(idx++) = (T)v; return *this; }
template <typename T> MatCommaInitializer_<T>::operator Mat_<T>() const { return *m; }
template <typename T> MatCommaInitializer_<T>::operator Mat() const { return *m; }

template <typename T> Mat_<T> operator*(const Mat_<T>& A, const Mat_<T>& B)
{
    Mat_<T> r(A.rows, B.cols);
    for (int i = 0; i < A.rows; ++i) for (int j = 0; j < B.cols; ++j) { T s = 0; for (int k = 0; k < A.cols; ++k) s += A(i, k) * B(k, j); r(i, j) = s; }
    return r;
}

// ---------------------------------------------------------------------------------------------- solve / SVD
namespace shim {
struct In {      // dense row-major double copy of any of the argument types the reference passes
    int rows = 0, cols = 0; std::vector<double> v;
    In(const Mat& m) : rows(m.rows), cols(m.cols), v((size_t)m.rows * m.cols) { for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) v[(size_t)i * cols + j] = m.at<double>(i, j); }
    template <typename T, int M, int N> In(const Matx<T, M, N>& m) : rows(M), cols(N), v(m.val, m.val + M * N) {}
    double operator()(int i, int j) const { return v[(size_t)i * cols + j]; }
};
inline void store(Mat& dst, const std::vector<double>& x, int n)
{
    if (dst.data == nullptr || dst.rows * dst.cols != n) dst.create(n, 1, CV_64FC1);
    for (int i = 0; i < n; ++i) ((double*)dst.data)[i] = x[i];
}
template <typename T, int N> void store(Vec<T, N>& dst, const std::vector<double>& x, int n) { for (int i = 0; i < n && i < N; ++i) dst[i] = (T)x[i]; }

// OpenCV's closed form for 3x3 systems (cv::solve, n <= 3: no pivoting, one reciprocal of the determinant)
inline bool solve3(const In& A, const In& b, std::vector<double>& x)
{
    const double a00 = A(0, 0), a01 = A(0, 1), a02 = A(0, 2), a10 = A(1, 0), a11 = A(1, 1), a12 = A(1, 2), a20 = A(2, 0),
                 a21 = A(2, 1), a22 = A(2, 2);
    double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
    x.assign(3, 0.0);
    if (d == 0) return false;
    const double b0 = b.v[0], b1 = b.v[1], b2 = b.v[2];
    d = 1. / d;
    x[0] = d * (b0 * (a11 * a22 - a12 * a21) - a01 * (b1 * a22 - a12 * b2) + a02 * (b1 * a21 - a11 * b2));
    x[1] = d * (a00 * (b1 * a22 - a12 * b2) - b0 * (a10 * a22 - a12 * a20) + a02 * (a10 * b2 - b1 * a20));
    x[2] = d * (a00 * (a11 * b2 - b1 * a21) - a01 * (a10 * b2 - b1 * a20) + b0 * (a10 * a21 - a11 * a20));
    return true;
}
// Gaussian elimination with partial pivoting (normal equations of the least-squares variants; never on the tested path)
inline bool gauss(std::vector<double> A, std::vector<double> b, int n, std::vector<double>& x)
{
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
        if (A[p * n + c] == 0) return false;
        for (int j = 0; j < n; ++j) std::swap(A[c * n + j], A[p * n + j]);
        std::swap(b[c], b[p]);
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r * n + c] / A[c * n + c];
            for (int j = c; j < n; ++j) A[r * n + j] -= f * A[c * n + j];
            b[r] -= f * b[c];
        }
    }
    x.assign(n, 0.0);
    for (int r = n - 1; r >= 0; --r) { double s = b[r]; for (int j = r + 1; j < n; ++j) s -= A[r * n + j] * x[j]; x[r] = s / A[r * n + r]; }
    return true;
}
}  // namespace shim

template <typename TA, typename TB, typename TX> bool solve(const TA& A_, const TB& b_, TX& x_, int flags = DECOMP_LU)
{
    const shim::In A(A_), b(b_);
    std::vector<double> x;
    bool ok;
    if (A.rows == 3 && A.cols == 3 && flags != DECOMP_SVD) {
        ok = shim::solve3(A, b, x);
    } else {
        const int n = A.cols;
        std::vector<double> N((size_t)n * n, 0.0), r(n, 0.0);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) for (int k = 0; k < A.rows; ++k) N[i * n + j] += A(k, i) * A(k, j);
            for (int k = 0; k < A.rows; ++k) r[i] += A(k, i) * b.v[k];
        }
        ok = shim::gauss(N, r, n, x);
    }
    shim::store(x_, x, A.cols);
    return ok;
}

// Singular value decomposition of a SYMMETRIC positive semi-definite matrix (all the reference asks of cv::SVD,
// PovMesh.cpp:641-644): cyclic Jacobi rotations, singular values sorted in descending order, vt rows = vectors.
class SVD {
public:
    Mat u, w, vt;
    SVD() {}
    SVD& operator()(const Mat& src, int = 0)
    {
        const int n = src.rows;
        std::vector<double> A((size_t)n * n), V((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[i * n + j] = 0.5 * (src.at<double>(i, j) + src.at<double>(j, i));
        for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0;
            for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
            if (off < 1e-300) break;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) {
                    if (A[p * n + q] == 0) continue;
                    const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * A[p * n + q]);
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                    const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                    for (int k = 0; k < n; ++k) {
                        const double akp = A[k * n + p], akq = A[k * n + q];
                        A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
                    }
                    for (int k = 0; k < n; ++k) {
                        const double apk = A[p * n + k], aqk = A[q * n + k];
                        A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
                    }
                    for (int k = 0; k < n; ++k) {
                        const double vkp = V[k * n + p], vkq = V[k * n + q];
                        V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
                    }
                }
        }
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(A[a * n + a]) > std::fabs(A[b * n + b]); });
        w.create(n, 1, CV_64FC1); vt.create(n, n, CV_64FC1); u.create(n, n, CV_64FC1);
        for (int i = 0; i < n; ++i) {
            w.at<double>(i, 0) = std::fabs(A[order[i] * n + order[i]]);
            for (int k = 0; k < n; ++k) { vt.at<double>(i, k) = V[k * n + order[i]]; u.at<double>(k, i) = V[k * n + order[i]]; }
        }
        return *this;
    }
};

// ---------------------------------------------------------------------------------------------- Mat arithmetic the
// triangulation stage of wass_stereo.cpp uses: 3x3 / 3x4 double products, and 8-bit mask images built as
// `img.clone()*0 + 1`, `mask.mul(1 - aux)`, cv::threshold(..., THRESH_BINARY)
inline unsigned char shim_sat8(double v) { v = std::nearbyint(v); return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
inline Mat operator*(const Mat& A, const Mat& B)
{
    Mat r(A.rows, B.cols, CV_64FC1);
    for (int i = 0; i < A.rows; ++i) for (int j = 0; j < B.cols; ++j) { double s = 0; for (int k = 0; k < A.cols; ++k) s += A.at<double>(i, k) * B.at<double>(k, j); r.at<double>(i, j) = s; }
    return r;
}
inline Mat operator-(const Mat& A) { Mat r = A.clone(); for (size_t i = 0; i < (size_t)A.rows * A.cols; ++i) ((double*)r.data)[i] = -((const double*)A.data)[i]; return r; }
inline Mat operator*(const Mat& A, double s)
{
    Mat r = A.clone();
    if (A.mtype == CV_64FC1) for (size_t i = 0; i < (size_t)A.rows * A.cols; ++i) ((double*)r.data)[i] *= s;
    else for (size_t i = 0; i < r.own.size(); ++i) r.data[i] = shim_sat8(A.data[i] * s);
    return r;
}
inline Mat operator+(const Mat& A, double s) { Mat r = A.clone(); for (size_t i = 0; i < r.own.size(); ++i) r.data[i] = shim_sat8(A.data[i] + s); return r; }
inline Mat operator-(double s, const Mat& A) { Mat r = A.clone(); for (size_t i = 0; i < r.own.size(); ++i) r.data[i] = shim_sat8(s - A.data[i]); return r; }
enum { THRESH_BINARY = 0, IMREAD_GRAYSCALE = 0 };
inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int)
{
    Mat r(src.rows, src.cols, CV_8UC1);
    for (size_t i = 0; i < (size_t)src.rows * src.cols; ++i) r.data[i] = src.data[i] > thresh ? shim_sat8(maxval) : 0;
    dst = r;
    return thresh;
}
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
// cv::imread(..., IMREAD_GRAYSCALE) for binary PGM (P5, maxval 255) only: all the checker's mask images need
inline Mat imread(const std::string& fn, int)
{
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) return Mat();
    int w = 0, h = 0, mx = 0;
    Mat r;
    if (fscanf(f, "P5 %d %d %d", &w, &h, &mx) == 3 && mx == 255 && fgetc(f) != EOF) {
        r.create(h, w, CV_8UC1);
        if (fread(r.data, 1, (size_t)w * h, f) != (size_t)w * h) r = Mat();
    }
    fclose(f);
    return r;
}

// ---------------------------------------------------------------------------------------------- image IO: not needed by the checker
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void resize(const Mat& src, Mat& dst, Size, double = 0, double = 0, int = 1) { dst = src; }

}  // namespace cv
