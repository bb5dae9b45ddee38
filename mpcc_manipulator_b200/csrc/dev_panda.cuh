// Panda forward kinematics, geometric Jacobian and manipulability in registers.
// Replaces the RBDL calls of the reference (cpp/src/Model/robot_model.cpp:366-450);
// chain constants from robot_model.cpp:170-263.
#pragma once
#include "mpcc_types.h"
#include <math.h>

namespace mpcc {

struct PandaKin {
    double p[3];      // EE (panda_hand_tcp) position, base frame
    double R[9];      // EE orientation, row-major
    double Jv[21];    // 3x7 linear Jacobian, row-major
    double Jw[21];    // 3x7 angular Jacobian, row-major
};

// Forward chain.  R_i = R_{i-1} E_i^T Rz(q_i), p_i = p_{i-1} + R_{i-1} r_i  (RBDL: child
// orientation in parent = E^T).  E_i is identity (i=1), Ea (i=2,5) or Eb (i=3,4,6,7).
// Writes the joint origins o[i] and joint axes z[i] (third column of R_i) for the Jacobian.
MPCC_HD void panda_chain(const double* q, double o[7][3], double z[7][3], double* p_ee, double* R_ee) {
    const double rx[7] = {0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088};
    const double ry[7] = {0.0, 0.0, -0.316, 0.0, 0.384, 0.0, 0.0};
    const double rz[7] = {0.333, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int et[7] = {0, 1, 2, 2, 1, 2, 2};  // 0: I, 1: Ea, 2: Eb
    // R stored as three columns
    double c0[3] = {1, 0, 0}, c1[3] = {0, 1, 0}, c2[3] = {0, 0, 1};
    double p[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 7; i++) {
        // p_i = p_{i-1} + R_{i-1} r_i
#pragma unroll
        for (int a = 0; a < 3; a++) p[a] += c0[a] * rx[i] + c1[a] * ry[i] + c2[a] * rz[i];
        // M = R E^T  (column permutation with sign)
        double m1[3], m2[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (et[i] == 0) { m1[a] = c1[a]; m2[a] = c2[a]; }
            else if (et[i] == 1) { m1[a] = -c2[a]; m2[a] = c1[a]; }
            else { m1[a] = c2[a]; m2[a] = -c1[a]; }
        }
        double s, c;
        sincos(q[i], &s, &c);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double n0 = c * c0[a] + s * m1[a];
            double n1 = -s * c0[a] + c * m1[a];
            c0[a] = n0; c1[a] = n1; c2[a] = m2[a];
            o[i][a] = p[a];
            z[i][a] = m2[a];
        }
    }
    // link7 -> hand: E8 = Rz(+45deg) with the literal 0.707107 (robot_model.cpp:238-242), R_ee = R7 E8^T;
    // hand -> hand_tcp: (0,0,0.1034) (robot_model.cpp:182); link7 -> hand: (0,0,0.107)
    const double k = 0.707107;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        R_ee[3 * a + 0] = k * c0[a] - k * c1[a];
        R_ee[3 * a + 1] = k * c0[a] + k * c1[a];
        R_ee[3 * a + 2] = c2[a];
        p_ee[a] = (p[a] + c2[a] * 0.107) + c2[a] * 0.1034;
    }
}

MPCC_HD void panda_kinematics(const double* q, PandaKin& k) {
    double o[7][3], z[7][3];
    panda_chain(q, o, z, k.p, k.R);
#pragma unroll
    for (int i = 0; i < 7; i++) {
        double dx = k.p[0] - o[i][0], dy = k.p[1] - o[i][1], dz = k.p[2] - o[i][2];
        k.Jv[0 * 7 + i] = z[i][1] * dz - z[i][2] * dy;
        k.Jv[1 * 7 + i] = z[i][2] * dx - z[i][0] * dz;
        k.Jv[2 * 7 + i] = z[i][0] * dy - z[i][1] * dx;
        k.Jw[0 * 7 + i] = z[i][0];
        k.Jw[1 * 7 + i] = z[i][1];
        k.Jw[2 * 7 + i] = z[i][2];
    }
}

// sqrt(det(J J^T)), J = [Jv; Jw] (robot_model.cpp:431-435).  J J^T is symmetric positive
// semi-definite, so the determinant is taken from an unpivoted LDL^T in registers.
MPCC_HD double panda_manipulability_from(const double* Jv, const double* Jw) {
    double A[6][6];
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int b = 0; b <= a; b++) {
            const double* ra = (a < 3) ? (Jv + 7 * a) : (Jw + 7 * (a - 3));
            const double* rb = (b < 3) ? (Jv + 7 * b) : (Jw + 7 * (b - 3));
            double s = 0;
#pragma unroll
            for (int k = 0; k < 7; k++) s += ra[k] * rb[k];
            A[a][b] = s;
        }
    }
    double det = 1.0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j][j];
        det *= d;
        double inv = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double l = A[i][j] * inv;
#pragma unroll
            for (int k = j + 1; k <= i; k++) A[i][k] -= l * A[k][j];
        }
    }
    return sqrt(det);
}

MPCC_HD double panda_manipulability(const double* q) {
    PandaKin k;
    panda_kinematics(q, k);
    return panda_manipulability_from(k.Jv, k.Jw);
}

// Central differences with delta = 1e-4 per joint, exactly as the reference does
// (robot_model.cpp:437-450): an analytic gradient would differ at ~1e-8 relative.
MPCC_HD void panda_dmanipulability(const double* q, double* d) {
    const double delta = 1e-4;
    double qq[7];
#pragma unroll
    for (int k = 0; k < 7; k++) qq[k] = q[k];
    for (int i = 0; i < 7; i++) {
        qq[i] = q[i] + delta;
        double m1 = panda_manipulability(qq);
        qq[i] = q[i] - delta;
        double m2 = panda_manipulability(qq);
        qq[i] = q[i];
        d[i] = (m1 - m2) / (2 * delta);
    }
}

}  // namespace mpcc
