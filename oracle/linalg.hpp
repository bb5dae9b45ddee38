// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under oracle/ is linked into or called from the product library.
//
// Minimal dense linear algebra used by the restatement (the reference uses
// Eigen, which is not available in this image).  Row-major, double only.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#include <array>
#include <algorithm>

namespace orc {

typedef std::array<double, 3> Vec3;
struct Mat3 {
    double m[9];  // row-major
    double& operator()(int r, int c) { return m[3 * r + c]; }
    double operator()(int r, int c) const { return m[3 * r + c]; }
    static Mat3 Zero() { Mat3 a; for (double& v : a.m) v = 0; return a; }
    static Mat3 Identity() { Mat3 a = Zero(); a.m[0] = a.m[4] = a.m[8] = 1; return a; }
};

inline Mat3 mul(const Mat3& a, const Mat3& b) {
    Mat3 c = Mat3::Zero();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
            c(i, j) = s;
        }
    return c;
}
inline Mat3 transpose(const Mat3& a) {
    Mat3 c;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) c(i, j) = a(j, i);
    return c;
}
inline Vec3 mul(const Mat3& a, const Vec3& v) {
    Vec3 r;
    for (int i = 0; i < 3; i++) r[i] = a(i, 0) * v[0] + a(i, 1) * v[1] + a(i, 2) * v[2];
    return r;
}
inline Mat3 scale(const Mat3& a, double s) { Mat3 c; for (int i = 0; i < 9; i++) c.m[i] = a.m[i] * s; return c; }
inline Mat3 add(const Mat3& a, const Mat3& b) { Mat3 c; for (int i = 0; i < 9; i++) c.m[i] = a.m[i] + b.m[i]; return c; }
inline Mat3 sub(const Mat3& a, const Mat3& b) { Mat3 c; for (int i = 0; i < 9; i++) c.m[i] = a.m[i] - b.m[i]; return c; }
inline double trace(const Mat3& a) { return a.m[0] + a.m[4] + a.m[8]; }
inline double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline Vec3 add(const Vec3& a, const Vec3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline Vec3 sub(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline Vec3 scale(const Vec3& a, double s) { return {a[0] * s, a[1] * s, a[2] * s}; }

// Dynamic dense matrix, row-major.
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    void setZero(int r_, int c_) { r = r_; c = c_; a.assign((size_t)r_ * c_, 0.0); }
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
    double* row(int i) { return &a[(size_t)i * c]; }
    const double* row(int i) const { return &a[(size_t)i * c]; }
};
typedef std::vector<double> Vec;

// In-place lower Cholesky of the leading n x n block of a (row-major, ld).
// Returns false if a non-positive pivot is met (Eigen LLT: NumericalIssue).
inline bool cholesky_inplace(double* a, int n, int ld) {
    for (int j = 0; j < n; j++) {
        double* aj = a + (size_t)j * ld;
        double d = aj[j];
        for (int k = 0; k < j; k++) d -= aj[k] * aj[k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        aj[j] = d;
        double inv = 1.0 / d;
        for (int i = j + 1; i < n; i++) {
            double* ai = a + (size_t)i * ld;
            double s = ai[j];
            for (int k = 0; k < j; k++) s -= ai[k] * aj[k];
            ai[j] = s * inv;
        }
    }
    return true;
}
// Eigen::LLT's verdict (info() != NumericalIssue): false at the first pivot x <= 0; NaN pivots are not flagged.
inline bool llt_no_numerical_issue(double* a, int n, int ld) {
    for (int j = 0; j < n; j++) {
        double* aj = a + (size_t)j * ld;
        double d = aj[j];
        for (int k = 0; k < j; k++) d -= aj[k] * aj[k];
        if (d <= 0.0) return false;
        d = std::sqrt(d);
        aj[j] = d;
        double inv = 1.0 / d;
        for (int i = j + 1; i < n; i++) {
            double* ai = a + (size_t)i * ld;
            double s = ai[j];
            for (int k = 0; k < j; k++) s -= ai[k] * aj[k];
            ai[j] = s * inv;
        }
    }
    return true;
}
// Solve L L^T x = b in place (L lower, row-major).
inline void cholesky_solve(const double* L, int n, int ld, double* x) {
    for (int i = 0; i < n; i++) {
        const double* Li = L + (size_t)i * ld;
        double s = x[i];
        for (int k = 0; k < i; k++) s -= Li[k] * x[k];
        x[i] = s / Li[i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[(size_t)k * ld + i] * x[k];
        x[i] = s / L[(size_t)i * ld + i];
    }
}
// Forward substitution only: solve L y = b in place.
inline void cholesky_forward(const double* L, int n, int ld, double* x) {
    for (int i = 0; i < n; i++) {
        const double* Li = L + (size_t)i * ld;
        double s = x[i];
        for (int k = 0; k < i; k++) s -= Li[k] * x[k];
        x[i] = s / Li[i];
    }
}

// Determinant of an n x n matrix (n <= 8) by LU with partial pivoting
// (what Eigen's MatrixXd::determinant() does for dynamic sizes > 4).
inline double det_lu(const double* in, int n) {
    double a[64];
    for (int i = 0; i < n * n; i++) a[i] = in[i];
    double det = 1.0;
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = std::fabs(a[k * n + k]);
        for (int i = k + 1; i < n; i++) {
            double v = std::fabs(a[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        if (best == 0.0) return 0.0;
        if (p != k) {
            for (int j = 0; j < n; j++) std::swap(a[k * n + j], a[p * n + j]);
            det = -det;
        }
        double piv = a[k * n + k];
        det *= piv;
        for (int i = k + 1; i < n; i++) {
            double f = a[i * n + k] / piv;
            for (int j = k + 1; j < n; j++) a[i * n + j] -= f * a[k * n + j];
        }
    }
    return det;
}

}  // namespace orc
