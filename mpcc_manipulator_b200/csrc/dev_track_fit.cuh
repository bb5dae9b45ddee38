// Track ingestion on the device: ArcLengthSpline::gen6DSpline / fitSpline for a whole batch of heterogeneous tracks
// (reference cpp/src/Spline/arc_length_spline.cpp:213-265; cubic_spline.cpp:65-124; cubic_spline_rot.cpp:139-238).
//
// One THREAD per track.  The fit is a chain of short sequential recurrences (three natural-spline tridiagonal sweeps per
// coordinate, two resampling rounds, SO(3) log / exp per resampled knot); the parallel axis that matters for configuration C4
// is the 65 536 tracks, not the 100 knots of one.  Every per-track array lives in a global scratch block laid out
// [element][track] (track fastest), so the threads of a warp -- which all walk the same element index -- read and write
// consecutive addresses.  The same code compiled for the host (stride 1) is what tests/emul checks against the host fit.
//
// Operation order follows csrc/host/track_fit.cpp line by line (which follows the reference), so both produce the same
// table up to FMA contraction.
#pragma once
#include "mpcc_types.h"
#include "dev_track.cuh"

namespace mpcc {

// strided view of one per-track array
struct TArr {
    double* p; size_t stride;
    MPCC_HD double& operator[](int i) const { return p[(size_t)i * stride]; }
    MPCC_HD TArr at(int off) const { return TArr{p + (size_t)off * stride, stride}; }
};

// scratch doubles per track for waypoint counts up to n (n >= 3)
MPCC_HD size_t track_fit_scratch_doubles(int n) {
    const size_t m = (size_t)((n > N_SPLINE) ? n : N_SPLINE);
    return 15 * m /* s | b,c,d x 3 | h, al, l, mu, z */ + 2 * (size_t)(12 * N_SPLINE + N_SPLINE) /* two resampled rounds: X,Y,Z,R9,s */;
}

// natural cubic spline through (x_i, a_i), i < n  (cubic_spline.cpp:65-124): b, c, d out; h, al, l, mu, z scratch
MPCC_HD void tf_cubic_fit(int n, const TArr& x, const TArr& a, const TArr& b, const TArr& c, const TArr& d, const TArr& h, const TArr& al, const TArr& l,
                          const TArr& mu, const TArr& z) {
    for (int i = 0; i < n - 1; i++) h[i] = x[i + 1] - x[i];
    for (int i = 1; i < n - 1; i++) al[i] = 3.0 / h[i] * (a[i + 1] - a[i]) - 3.0 / h[i - 1] * (a[i] - a[i - 1]);
    l[0] = 1.0; mu[0] = 0.0; z[0] = 0.0;
    for (int i = 1; i < n - 1; i++) {
        l[i] = 2.0 * (x[i + 1] - x[i - 1]) - h[i - 1] * mu[i - 1];
        mu[i] = h[i] / l[i];
        z[i] = (al[i] - h[i - 1] * z[i - 1]) / l[i];
    }
    c[n - 1] = 0.0;
    for (int i = n - 2; i >= 0; i--) {
        c[i] = z[i] - mu[i] * c[i + 1];
        b[i] = (a[i + 1] - a[i]) / h[i] - (h[i] * (c[i + 1] + 2.0 * c[i])) / 3.0;
        d[i] = (c[i + 1] - c[i]) / (3.0 * h[i]);
    }
}
// last knot <= v on strictly increasing knots (std::map::upper_bound - 1, cubic_spline.cpp:144-152); exact end -> n - 1
MPCC_HD int tf_segment(int n, const TArr& x, double v) {
    if (v == x[n - 1]) return n - 1;
    int lo = 0, hi = n;  // first index with x[i] > v
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] > v) hi = mid; else lo = mid + 1; }
    return lo - 1;
}
MPCC_HD double tf_cubic_eval(int n, const TArr& x, const TArr& a, const TArr& b, const TArr& c, const TArr& d, double v) {
    v = fmax(0.0, fmin(v, x[n - 1]));
    const int i = tf_segment(n, x, v);
    if (i == n - 1) return a[n - 1];
    const double dx = v - x[i];
    return a[i] + b[i] * dx + c[i] * (dx * dx) + d[i] * (dx * (dx * dx));
}
// SO(3) cubic spline through rotations on irregular knots (cubic_spline_rot.cpp:139-238); R: n x 9 (element-major view)
MPCC_HD void tf_rot_eval(int n, const TArr& x, const TArr& R, double v, double* out) {
    v = fmax(0.0, fmin(v, x[n - 1]));
    const int i = tf_segment(n, x, v);
    double Ri[9];
    if (i == n - 1) { for (int k = 0; k < 9; k++) out[k] = R[9 * (n - 1) + k]; return; }
    double Rn[9];
    for (int k = 0; k < 9; k++) { Ri[k] = R[9 * i + k]; Rn[k] = R[9 * (i + 1) + k]; }
    const double h = x[i + 1] - x[i];
    const double c = 3.0 / (h * h), d = -2.0 / (h * h * h);
    const double dx = v - x[i], dx2 = dx * dx, dx3 = dx * dx2;
    double RR[9], w[3], E[9];
    mat3_tmul(Ri, Rn, RR);
    so3_log(RR, w);
    const double f = c * dx2 + d * dx3;
    const double wv[3] = {w[0] * f, w[1] * f, w[2] * f};
    so3_exp(wv, E);
    mat3_mul(Ri, E, out);
}
MPCC_HD void tf_arc_length(int n, const TArr& X, const TArr& Y, const TArr& Z, const TArr& s) {
    s[0] = 0.0;
    for (int i = 0; i + 1 < n; i++) {
        const double dx = X[i + 1] - X[i], dy = Y[i + 1] - Y[i], dz = Z[i + 1] - Z[i];
        s[i + 1] = s[i] + sqrt(dx * dx + dy * dy + dz * dz);
    }
}

// one "fit on s, resample at N_SPLINE equidistant arc lengths" round (arc_length_spline.cpp:89-118)
// in: n knots (s, X, Y, Z, R); out: N_SPLINE knots (so, Xo, Yo, Zo, Ro); w: 14 m doubles of scratch (b,c,d x 3 | h,al,l,mu,z)
MPCC_HD void tf_fit_resample(int n, int m, const TArr& s, const TArr& X, const TArr& Y, const TArr& Z, const TArr& R, const TArr& w, const TArr& so,
                             const TArr& Xo, const TArr& Yo, const TArr& Zo, const TArr& Ro) {
    const TArr in[3] = {X, Y, Z};
    const TArr out[3] = {Xo, Yo, Zo};
    const TArr h = w.at(9 * m), al = w.at(10 * m), l = w.at(11 * m), mu = w.at(12 * m), z = w.at(13 * m);
    for (int a = 0; a < 3; a++) {
        for (int i = 0; i < n; i++) al[i] = 0.0;
        tf_cubic_fit(n, s, in[a], w.at((3 * a) * m), w.at((3 * a + 1) * m), w.at((3 * a + 2) * m), h, al, l, mu, z);
    }
    const double total = s[n - 1];
    const double step = (total - 0.0) / (N_SPLINE - 1);  // Eigen setLinSpaced(size, 0, total)
    for (int i = 0; i < N_SPLINE; i++) so[i] = (i == N_SPLINE - 1) ? total : 0.0 + i * step;
    for (int i = 0; i < N_SPLINE; i++) {
        const double v = so[i];
        for (int a = 0; a < 3; a++) out[a][i] = tf_cubic_eval(n, s, in[a], w.at((3 * a) * m), w.at((3 * a + 1) * m), w.at((3 * a + 2) * m), v);
        double Rv[9];
        tf_rot_eval(n, s, R, v, Rv);
        for (int k = 0; k < 9; k++) Ro[9 * i + k] = Rv[k];
    }
}

// the final regular spline on 100 knots (setRegularData + genSpline(..., true), arc_length_spline.cpp:244-252) -> table
MPCC_HD void tf_table_from_knots(const TArr& s2, const TArr& X, const TArr& Y, const TArr& Z, const TArr& R, const TArr& w, int m, TrackTable& t) {
    const TArr in[3] = {X, Y, Z};
    const TArr h = w.at(9 * m), al = w.at(10 * m), l = w.at(11 * m), mu = w.at(12 * m), z = w.at(13 * m);
    for (int a = 0; a < 3; a++) {
        const TArr b = w.at((3 * a) * m), c = w.at((3 * a + 1) * m), d = w.at((3 * a + 2) * m);
        for (int i = 0; i < N_SPLINE; i++) al[i] = 0.0;
        tf_cubic_fit(N_SPLINE, s2, in[a], b, c, d, h, al, l, mu, z);
        for (int i = 0; i < N_SPLINE; i++) {
            t.a[a][i] = in[a][i];
            t.c[a][i] = c[i];
            t.b[a][i] = (i < N_SPLINE - 1) ? b[i] : 0.0;
            t.d[a][i] = (i < N_SPLINE - 1) ? d[i] : 0.0;
        }
    }
    for (int i = 0; i < N_SPLINE; i++) {
        t.s[i] = s2[i];
        for (int k = 0; k < 9; k++) t.R[i][k] = R[9 * i + k];
    }
    for (int i = 0; i < N_SPLINE; i++) {
        if (i < N_SPLINE - 1) {
            const double hh = s2[i + 1] - s2[i];
            t.rc[i] = 3.0 / (hh * hh);
            t.rd[i] = -2.0 / (hh * hh * hh);
            double RR[9], wv[3];
            mat3_tmul(t.R[i], t.R[i + 1], RR);
            so3_log(RR, wv);
            t.w[i][0] = wv[0]; t.w[i][1] = wv[1]; t.w[i][2] = wv[2];
        } else {
            t.rc[i] = t.rd[i] = 0.0;
            t.w[i][0] = t.w[i][1] = t.w[i][2] = 0.0;
        }
    }
    t.delta = s2[1] - s2[0];
    t.length = s2[N_SPLINE - 1];
    t.pad[0] = t.pad[1] = 0.0;
}

// fitSpline: chord-length fit -> resample 100 -> re-measure -> refit -> resample 100 -> regular spline.
// X, Y, Z [n], R [n][9]: this track's waypoints (views); ws: track_fit_scratch_doubles(n) doubles (view)
MPCC_HD void tf_fit_track(int n, const TArr& X, const TArr& Y, const TArr& Z, const TArr& R, const TArr& ws, TrackTable& t) {
    const int m = (n > N_SPLINE) ? n : N_SPLINE;
    const TArr s = ws, w = ws.at(m);
    const TArr r1 = ws.at(15 * m), r2 = ws.at(15 * m + 13 * N_SPLINE);
    const TArr X1 = r1, Y1 = r1.at(N_SPLINE), Z1 = r1.at(2 * N_SPLINE), R1 = r1.at(3 * N_SPLINE), s1 = r1.at(12 * N_SPLINE);
    const TArr X2 = r2, Y2 = r2.at(N_SPLINE), Z2 = r2.at(2 * N_SPLINE), R2 = r2.at(3 * N_SPLINE), s2 = r2.at(12 * N_SPLINE);
    tf_arc_length(n, X, Y, Z, s);
    tf_fit_resample(n, m, s, X, Y, Z, R, w, s1, X1, Y1, Z1, R1);
    tf_arc_length(N_SPLINE, X1, Y1, Z1, s);
    tf_fit_resample(N_SPLINE, m, s, X1, Y1, Z1, R1, w, s2, X2, Y2, Z2, R2);
    tf_table_from_knots(s2, X2, Y2, Z2, R2, w, m, t);
}

}  // namespace mpcc
