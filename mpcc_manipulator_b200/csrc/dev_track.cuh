// Track (arc-length spline) evaluation on the fitted table, SO(3) log/exp, projection.
// Follows cpp/src/Spline/{cubic_spline,cubic_spline_rot,arc_length_spline}.cpp of the reference.
#pragma once
#include "mpcc_types.h"
#include <math.h>

namespace mpcc {

MPCC_HD void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A^T B
MPCC_HD void mat3_tmul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}

// vee(Log(R)) of a rotation matrix (cubic_spline_rot.cpp:44-79).
// Branches: |tr+1|<1e-6 -> rotation by pi about the eigenvector of eigenvalue 1 (returned as
// -v*pi like the reference; the eigenvector's sign is implementation-defined in Eigen, here the
// largest component is made positive); |tr-3|<1e-6 -> 0; else theta/(2 sin theta) (R - R^T)^vee.
MPCC_HD void so3_log(const double* R, double* w) {
    double tr = R[0] + R[4] + R[8];
    if (fabs(tr + 1.0) < 1e-6) {
        double d0 = (R[0] + 1) * 0.5, d1 = (R[4] + 1) * 0.5, d2 = (R[8] + 1) * 0.5;
        int c = 0;
        if (d1 > d0) c = 1;
        if (d2 > ((c == 0) ? d0 : d1)) c = 2;
        double v[3];
#pragma unroll
        for (int a = 0; a < 3; a++) v[a] = (R[3 * a + c] + (a == c ? 1.0 : 0.0)) * 0.5;
        double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        int big = 0;
        if (fabs(v[1]) > fabs(v[big])) big = 1;
        if (fabs(v[2]) > fabs(v[big])) big = 2;
        double sg = (v[big] < 0 ? -1.0 : 1.0) / n;
#pragma unroll
        for (int a = 0; a < 3; a++) w[a] = -(v[a] * sg) * M_PI;
        return;
    }
    if (fabs(tr - 3.0) < 1e-6) { w[0] = w[1] = w[2] = 0.0; return; }
    double th = acos((tr - 1.0) / 2.0);
    double f = 1.0 / 2.0 * th / sin(th);
    w[0] = f * (R[7] - R[5]);
    w[1] = f * (R[2] - R[6]);
    w[2] = f * (R[3] - R[1]);
}

// Exp of the skew matrix of v (cubic_spline_rot.cpp:81-95), incl. the small-angle branch
// I + cos(|v|) [v]x  (the reference's "1/2" is an integer division, so no quadratic term).
MPCC_HD void so3_exp(const double* v, double* E) {
    double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double a, b;
    if (vn <= 1e-8) { a = cos(vn); b = 0.0; }
    else { a = sin(vn) / vn; b = (1 - cos(vn)) / (vn * vn); }
    double x = v[0], y = v[1], z = v[2];
    // [v]x^2 = v v^T - |v|^2 I
    double n2 = x * x + y * y + z * z;
    E[0] = 1.0 + b * (x * x - n2); E[1] = -a * z + b * (x * y);   E[2] = a * y + b * (x * z);
    E[3] = a * z + b * (x * y);    E[4] = 1.0 + b * (y * y - n2); E[5] = -a * x + b * (y * z);
    E[6] = -a * y + b * (x * z);   E[7] = a * x + b * (y * z);    E[8] = 1.0 + b * (z * z - n2);
}

struct TrackPoint {
    double pos[3], dpos[3], ddpos[3];
};

// getIndex for the regular final spline (cubic_spline.cpp:126-153) after unwrapInput (:155-160)
MPCC_HD int track_index(const TrackTable& t, double& s) {
    s = fmax(0.0, fmin(s, t.s[N_SPLINE - 1]));
    if (s == t.s[N_SPLINE - 1]) return N_SPLINE - 1;
    return (int)floor(s / t.delta);
}

// position, first and second derivative of the three cubic splines (cubic_spline.cpp:185-246)
MPCC_HD void track_eval_pos(const TrackTable& t, double s, TrackPoint& o) {
    int idx = track_index(t, s);
    double dx = s - t.s[idx];
    bool last = (idx == N_SPLINE - 1);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (last) {
            o.pos[a] = t.a[a][N_SPLINE - 1];
            o.dpos[a] = 0.0;
            o.ddpos[a] = 2.0 * t.c[a][idx];
        } else {
            double A = t.a[a][idx], B = t.b[a][idx], C = t.c[a][idx], D = t.d[a][idx];
            o.pos[a] = A + B * dx + C * (dx * dx) + D * (dx * (dx * dx));
            o.dpos[a] = B + 2.0 * C * dx + 3.0 * D * (dx * dx);
            o.ddpos[a] = 2.0 * C + 6.0 * D * dx;
        }
    }
}

// reference orientation and its derivative vector (cubic_spline_rot.cpp:216-259)
MPCC_HD void track_eval_rot(const TrackTable& t, double s, double* R, double* dR) {
    int idx = track_index(t, s);
    if (idx == N_SPLINE - 1) {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = t.R[N_SPLINE - 1][i];
        dR[0] = dR[1] = dR[2] = 0.0;
        return;
    }
    double dx = s - t.s[idx], dx2 = dx * dx, dx3 = dx * dx2;
    double f = t.rc[idx] * dx2 + t.rd[idx] * dx3;
    double v[3] = {t.w[idx][0] * f, t.w[idx][1] * f, t.w[idx][2] * f};
    double E[9];
    so3_exp(v, E);
    mat3_mul(t.R[idx], E, R);
    double g = 2.0 * t.rc[idx] * dx + 3.0 * t.rd[idx] * dx2;
#pragma unroll
    for (int a = 0; a < 3; a++) dR[a] = t.w[idx][a] * g;
}

// ArcLengthSpline::projectOnSpline (arc_length_spline.cpp:318-379)
MPCC_HD double track_project(const TrackTable& t, double max_dist_proj, double s, const double* ee) {
    const double s_guess = s;
    TrackPoint tp;
    track_eval_pos(t, s_guess, tp);
    double s_opt = s_guess;
    double ex = ee[0] - tp.pos[0], ey = ee[1] - tp.pos[1], ez = ee[2] - tp.pos[2];
    double dist = sqrt(ex * ex + ey * ey + ez * ez);
    const double L = t.s[N_SPLINE - 1];
    if (dist >= max_dist_proj) {
        int min_all = 0, min_valid = -1;
        double best_all = INFINITY, best_valid = INFINITY;
        for (int i = 0; i < N_SPLINE; i++) {
            // knot positions are the spline's a-coefficients (y_data)
            double dx = t.a[0][i] - ee[0], dy = t.a[1][i] - ee[1], dz = t.a[2][i] - ee[2];
            double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best_all) { best_all = d2; min_all = i; }
            bool valid = fabs(t.s[i] - s_guess) <= max_dist_proj;
            if (valid && d2 < best_valid) { best_valid = d2; min_valid = i; }
        }
        s_opt = (min_valid < 0) ? t.s[min_all] : t.s[min_valid];
    }
    if (s_opt >= L) return L;
    double s_old = s_opt;
    for (int i = 0; i < 20; i++) {
        track_eval_pos(t, s_opt, tp);
        double fx = tp.pos[0] - ee[0], fy = tp.pos[1] - ee[1], fz = tp.pos[2] - ee[2];
        double jac = 2.0 * fx * tp.dpos[0] + 2.0 * fy * tp.dpos[1] + 2.0 * fz * tp.dpos[2];
        double hes = 2.0 * tp.dpos[0] * tp.dpos[0] + 2.0 * fx * tp.ddpos[0] + 2.0 * tp.dpos[1] * tp.dpos[1] + 2.0 * fy * tp.ddpos[1] +
                     2.0 * tp.dpos[2] * tp.dpos[2] + 2.0 * fz * tp.ddpos[2];
        s_opt -= jac / hes;
        s_opt = fmax(0.0, fmin(s_opt, L));
        if (fabs(s_old - s_opt) <= 1e-5) return s_opt;
        s_old = s_opt;
    }
    return s_guess;
}

}  // namespace mpcc
