#include "track_fit.h"
#include "json_min.h"
#include "../dev_track.cuh"
#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace mpcc {

Waypoints load_track_json(const std::string& file) {
    json::Value j = json::parse_file(file);
    Waypoints w;
    w.X = j.numbers("X");
    w.Y = j.numbers("Y");
    w.Z = j.numbers("Z");
    std::vector<double> qx = j.numbers("quat_X"), qy = j.numbers("quat_Y"), qz = j.numbers("quat_Z"), qw = j.numbers("quat_W");
    size_t n = w.X.size();
    if (w.Y.size() != n || w.Z.size() != n || qx.size() != n || qy.size() != n || qz.size() != n || qw.size() != n)
        throw std::runtime_error("track: arrays of different length in '" + file + "'");
    w.R.resize(9 * n);
    for (size_t i = 0; i < n; i++) {
        // Eigen::Quaterniond(x,y,z,w).normalized().toRotationMatrix() (track.cpp:44-52)
        double nn = std::sqrt(qx[i] * qx[i] + qy[i] * qy[i] + qz[i] * qz[i] + qw[i] * qw[i]);
        double x = qx[i] / nn, y = qy[i] / nn, z = qz[i] / nn, s = qw[i] / nn;
        double tx = 2 * x, ty = 2 * y, tz = 2 * z;
        double twx = tx * s, twy = ty * s, twz = tz * s, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
        double* R = &w.R[9 * i];
        R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
        R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
        R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
    }
    return w;
}

void shift_track(Waypoints& w, const double p[3]) {
    double x0 = w.X[0], y0 = w.Y[0], z0 = w.Z[0];
    for (size_t i = 0; i < w.size(); i++) {
        w.X[i] = w.X[i] - x0 + p[0];
        w.Y[i] = w.Y[i] - y0 + p[1];
        w.Z[i] = w.Z[i] - z0 + p[2];
    }
}

namespace {

// natural cubic spline through (x_i, y_i) (cubic_spline.cpp:65-124): tridiagonal sweep
struct Cubic {
    std::vector<double> x, a, b, c, d;
    void fit(const std::vector<double>& xs, const std::vector<double>& ys) {
        const int n = (int)xs.size();
        x = xs; a = ys;
        b.assign(n - 1, 0.0); c.assign(n, 0.0); d.assign(n - 1, 0.0);
        std::vector<double> h(n - 1), al(n - 1, 0.0), l(n), mu(n - 1), z(n);
        for (int i = 0; i < n - 1; i++) h[i] = x[i + 1] - x[i];
        for (int i = 1; i < n - 1; i++) al[i] = 3.0 / h[i] * (a[i + 1] - a[i]) - 3.0 / h[i - 1] * (a[i] - a[i - 1]);
        l[0] = 1.0; mu[0] = 0.0; z[0] = 0.0;
        for (int i = 1; i < n - 1; i++) {
            l[i] = 2.0 * (x[i + 1] - x[i - 1]) - h[i - 1] * mu[i - 1];
            mu[i] = h[i] / l[i];
            z[i] = (al[i] - h[i - 1] * z[i - 1]) / l[i];
        }
        l[n - 1] = 1.0; z[n - 1] = 0.0; c[n - 1] = 0.0;
        for (int i = n - 2; i >= 0; i--) {
            c[i] = z[i] - mu[i] * c[i + 1];
            b[i] = (a[i + 1] - a[i]) / h[i] - (h[i] * (c[i + 1] + 2.0 * c[i])) / 3.0;
            d[i] = (c[i + 1] - c[i]) / (3.0 * h[i]);
        }
    }
    // segment of an irregular spline: the reference keeps a std::map from knot value to index and
    // takes upper_bound(x) - 1 (cubic_spline.cpp:144-152); on strictly increasing knots that is the
    // last knot <= x.  Exact end point -> last knot (:133-136).
    int segment(double v) const {
        const int n = (int)x.size();
        if (v == x[n - 1]) return n - 1;
        return (int)(std::upper_bound(x.begin(), x.end(), v) - x.begin()) - 1;
    }
    double eval(double v) const {
        const int n = (int)x.size();
        v = std::max(0.0, std::min(v, x[n - 1]));
        int i = segment(v);
        if (i == n - 1) return a[n - 1];
        double dx = v - x[i];
        return a[i] + b[i] * dx + c[i] * (dx * dx) + d[i] * (dx * (dx * dx));
    }
};

// SO(3) cubic spline through rotations (cubic_spline_rot.cpp:139-238), irregular knots
struct RotSpline {
    std::vector<double> x, R;
    void eval(double v, double* out) const {
        const int n = (int)x.size();
        v = std::max(0.0, std::min(v, x[n - 1]));
        int i = (v == x[n - 1]) ? n - 1 : (int)(std::upper_bound(x.begin(), x.end(), v) - x.begin()) - 1;
        if (i == n - 1) { for (int k = 0; k < 9; k++) out[k] = R[9 * (n - 1) + k]; return; }
        double h = x[i + 1] - x[i];
        double c = 3.0 / std::pow(h, 2), d = -2.0 / std::pow(h, 3);
        double dx = v - x[i], dx2 = dx * dx, dx3 = dx * dx2;
        double RR[9], w[3], E[9];
        mat3_tmul(&R[9 * i], &R[9 * (i + 1)], RR);
        so3_log(RR, w);
        double f = c * dx2 + d * dx3;
        double wv[3] = {w[0] * f, w[1] * f, w[2] * f};
        so3_exp(wv, E);
        mat3_mul(&R[9 * i], E, out);
    }
};

std::vector<double> arc_length(const std::vector<double>& X, const std::vector<double>& Y, const std::vector<double>& Z) {
    std::vector<double> s(X.size(), 0.0);
    for (size_t i = 0; i + 1 < X.size(); i++) {
        double dx = X[i + 1] - X[i], dy = Y[i + 1] - Y[i], dz = Z[i + 1] - Z[i];
        s[i + 1] = s[i] + std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    return s;
}

// one "fit on s, resample at N_SPLINE equidistant arc lengths" round (arc_length_spline.cpp:89-118)
void fit_resample(const std::vector<double>& s, const Waypoints& in, double total, Waypoints& out, std::vector<double>& s_out) {
    Cubic cx, cy, cz;
    RotSpline cr;
    cx.fit(s, in.X); cy.fit(s, in.Y); cz.fit(s, in.Z);
    cr.x = s; cr.R = in.R;
    s_out.resize(N_SPLINE);
    double step = (total - 0.0) / (N_SPLINE - 1);  // Eigen setLinSpaced(size, 0, total)
    for (int i = 0; i < N_SPLINE; i++) s_out[i] = (i == N_SPLINE - 1) ? total : 0.0 + i * step;
    out.X.resize(N_SPLINE); out.Y.resize(N_SPLINE); out.Z.resize(N_SPLINE); out.R.resize(9 * N_SPLINE);
    for (int i = 0; i < N_SPLINE; i++) {
        out.X[i] = cx.eval(s_out[i]); out.Y[i] = cy.eval(s_out[i]); out.Z[i] = cz.eval(s_out[i]);
        cr.eval(s_out[i], &out.R[9 * i]);
    }
}

}  // namespace

void fit_track(const Waypoints& w, TrackTable& t) {
    if (w.size() < 3) throw std::runtime_error("track: need at least 3 waypoints");
    std::vector<double> s = arc_length(w.X, w.Y, w.Z), s1, s2;
    Waypoints first, second;
    fit_resample(s, w, s.back(), first, s1);
    s = arc_length(first.X, first.Y, first.Z);
    fit_resample(s, first, s.back(), second, s2);
    // final regular spline on (s2, second): setRegularData + genSpline(...,true) (:244-252)
    Cubic c[3];
    c[0].fit(s2, second.X); c[1].fit(s2, second.Y); c[2].fit(s2, second.Z);
    for (int i = 0; i < N_SPLINE; i++) {
        t.s[i] = s2[i];
        for (int a = 0; a < 3; a++) {
            t.a[a][i] = c[a].a[i];
            t.c[a][i] = c[a].c[i];
            t.b[a][i] = (i < N_SPLINE - 1) ? c[a].b[i] : 0.0;
            t.d[a][i] = (i < N_SPLINE - 1) ? c[a].d[i] : 0.0;
        }
        for (int k = 0; k < 9; k++) t.R[i][k] = second.R[9 * i + k];
    }
    for (int i = 0; i < N_SPLINE; i++) {
        if (i < N_SPLINE - 1) {
            double h = s2[i + 1] - s2[i];
            t.rc[i] = 3.0 / std::pow(h, 2);
            t.rd[i] = -2.0 / std::pow(h, 3);
            double RR[9];
            mat3_tmul(t.R[i], t.R[i + 1], RR);
            so3_log(RR, t.w[i]);
        } else {
            t.rc[i] = t.rd[i] = 0.0;
            t.w[i][0] = t.w[i][1] = t.w[i][2] = 0.0;
        }
    }
    t.delta = s2[1] - s2[0];
    t.length = s2[N_SPLINE - 1];
    t.pad[0] = t.pad[1] = 0.0;
}

}  // namespace mpcc
