// Per-stage linearisation of the MPCC problem: cost value / gradient / Gauss-Newton Hessian
// (reference cpp/src/Cost/cost.cpp), polytopic rows (cpp/src/Constraints/constraints.cpp),
// box / rate bounds (cpp/src/Constraints/bounds.cpp, osqp_interface.cpp:254-300) and the
// dynamics defect (osqp_interface.cpp:221-252), produced directly as the blocks of the
// stage-structured QP in the reference's normalised step variables (T_x, T_u scaling).
#pragma once
#include "mpcc_types.h"
#include "dev_track.cuh"

namespace mpcc {

// strided view of one (instance, stage) RobotData record
struct RbView {
    const double* p;
    size_t stride;
    MPCC_HD double operator()(int e) const { return p[(size_t)e * stride]; }
};

// symmetric 9x9 packed lower: index(r,c), r>=c
MPCC_HD int sym9(int r, int c) { return r * (r + 1) / 2 + c; }

struct StageLin {
    double Q[45];          // T_x f_xx T_x  (packed lower)
    double q[NX];          // T_x f_x
    double Rd[NU];         // diagonal of T_u f_uu T_u + joint-acceleration diagonal (osqp_interface.cpp:193-216)
    double r[NU];          // T_u f_u + joint-acceleration gradient (osqp_interface.cpp:176-191)
    double b[NX];          // dynamics defect: xi_{k+1} = A xi_k + B nu_k + b  (k < N)
    double xlo[NX], xhi[NX];   // box on xi_k from the state-bound rows
    double dlo[DOF], dhi[DOF]; // rate rows: nu_k[j] - nu_{k-1}[j] in [dlo, dhi]  (k < N)
    double pg[NPC * DOF];  // gradients grad h_j of the 11 polytopic rows (k < N)
    double pd[NPC];        // RBF'(h_j)
    double prhs[NPC];      // -c_j: row is  sum_m pd*pg*Tx xi[m] - pg*Tu nu[m] <= prhs
    double obj;            // this stage's share of the objective (filter)
    double gap;            // this stage's share of the l1 constraint violation (filter)
};

MPCC_HD double rbf(double delta, double h) {  // constraints.cpp:34-43
    if (h >= delta) return -log(h + 1);
    return -log(delta + 1) - 1 / (delta + 1) * (h - delta) + 1 / (2 * ((delta + 1) * (delta + 1))) * ((h - delta) * (h - delta));
}
MPCC_HD double drbf(double delta, double h) {  // constraints.cpp:52-61
    if (h >= delta) return -1 / (h + 1);
    return -1 / (delta + 1) + 1 / ((delta + 1) * (delta + 1)) * (h - delta);
}
MPCC_HD double smoothstep_weight(double x, double x0, double xf, double y0, double yf) {  // cost.cpp:36-43
    double t = (x - x0) / (xf - x0);
    return y0 + (yf - y0) * (3 * (t * t) - 2 * (t * t * t));
}
MPCC_HD double viol(double c, double l, double u) { return fmax(l - c, 0.0) + fmax(c - u, 0.0); }

// FULL = true : value + gradient + Hessian + constraint rows (setQP with all outputs)
// FULL = false: objective and constraint violation only (filter line search, osqp_interface.cpp:773-775)
//   x, u        this stage's state / input of the current iterate (u ignored at k == N)
//   u_prev      dq of stage k-1, or the currently applied input for k == 0
//   u_next      dq of stage k+1 (read only if k <= N-2)
//   x_next      state of stage k+1 (read only if k < N)
//   rbf_pre     optional: RBF(h_j) [11] | RBF'(h_j) [11] of this stage, computed by an earlier call.  h_j depends on the
//               (frozen) RobotData and the parameters only, so the values are constant over a control cycle.
//   rbf_out     optional: where to store them
//   hess_only   (FULL) fill only the Hessian blocks Q, Rd -- what the positive-definiteness / NaN test of
//               osqp_interface.cpp:454-473 looks at; used when the QP of this linearisation is already known to fail
template <bool FULL>
MPCC_HDN void stage_eval(const Params& P, const TrackTable& T, double Ts, int N, int k, const double* x, const double* u,
                         const double* u_prev, const double* u_next, const double* x_next, const RbView& rb, StageLin& o,
                         const double* rbf_pre = nullptr, double* rbf_out = nullptr, bool hess_only = false) {
    const double s = x[7], vs = x[8];
    const bool term = (k == N);

    // ---- adaptive weights (cost.cpp:293-308) ----
    const double manip = rb(RB_MANIP), sel = rb(RB_SEL);
    double ratio = fmin(sel / (P.tol_selcol * 2.0), manip / (P.tol_sing * 2.0));
    double wc = P.q_c, wl = P.q_l, wo = P.q_ori;
    if (ratio <= 1.0) {
        wc = P.q_c * smoothstep_weight(ratio, 0.5, 1.0, P.q_c_red_ratio, 1.0);
        wl = P.q_l * smoothstep_weight(ratio, 0.5, 1.0, P.q_l_inc_ratio, 1.0);
        wo = P.q_ori * smoothstep_weight(ratio, 0.5, 1.0, P.q_ori_red_ratio, 1.0);
    }
    const double cc0 = term ? P.q_c_N_mult * wc : wc;  // cost.cpp:128

    // ---- reference point (cost.cpp:46-80) ----
    TrackPoint tp;
    track_eval_pos(T, s, tp);
    const double Tg[3] = {tp.dpos[0], tp.dpos[1], tp.dpos[2]};
    const double Nn[3] = {tp.ddpos[0], tp.ddpos[1], tp.ddpos[1]};  // ddz_ref = ddpos_ref(1), cost.cpp:65

    // ---- contouring / lag errors (cost.cpp:82-117) ----
    double e[3], el[3], ec[3];
#pragma unroll
    for (int a = 0; a < 3; a++) e[a] = rb(RB_P + a) - tp.pos[a];
    const double Te = Tg[0] * e[0] + Tg[1] * e[1] + Tg[2] * e[2];
#pragma unroll
    for (int a = 0; a < 3; a++) { el[a] = Te * Tg[a]; ec[a] = e[a] - el[a]; }
    const double ec2 = ec[0] * ec[0] + ec[1] * ec[1] + ec[2] * ec[2];
    const double el2 = el[0] * el[0] + el[1] * el[1] + el[2] * el[2];

    const double L = T.s[N_SPLINE - 1];
    const double v_des = (s < L * P.deacc_ratio) ? P.desired_ee_velocity : -P.desired_ee_velocity / (L * P.deacc_ratio) * (s - L);  // cost.cpp:133-134
    const double dv = vs - v_des;

    // ---- heading error (cost.cpp:164-207) ----
    double Rref[9], dRref[3], Ree[9], Rbar[9], lg[3];
    track_eval_rot(T, s, Rref, dRref);
#pragma unroll
    for (int i = 0; i < 9; i++) Ree[i] = rb(RB_R + i);
    mat3_tmul(Rref, Ree, Rbar);
    so3_log(Rbar, lg);
    const double lg2 = lg[0] * lg[0] + lg[1] * lg[1] + lg[2] * lg[2];

    double obj = cc0 * ec2 + wl * el2 + P.q_vs * (dv * dv) + wo * lg2 - P.q_sing * manip;
    double dq2 = 0;
    if (!term) {
#pragma unroll
        for (int j = 0; j < DOF; j++) dq2 += u[j] * u[j];
        obj += P.r_dq * dq2 + P.r_dVs * (u[7] * u[7]);  // cost.cpp:229-230
        if (k != N - 1) {                              // osqp_interface.cpp:169-175
            double a2 = 0;
#pragma unroll
            for (int j = 0; j < DOF; j++) { double d = u_next[j] - u[j]; a2 += d * d; }
            obj += P.r_ddq_solver * a2;
        }
    }
    o.obj = obj;

    // ---- l1 violation of every row this stage owns (osqp_interface.cpp:824-833) ----
    double gap = 0;
    if (!(FULL && hess_only)) {
        // state box with the s trust region around the iterate itself (bounds.cpp:85-103)
#pragma unroll
        for (int m = 0; m < NX; m++) {
            double lo = P.lx[m], hi = P.ux[m];
            if (m == 7) { lo = fmax(s - P.s_trust_region, 0.0); hi = fmin(s + P.s_trust_region, L); }
            gap += viol(x[m], lo, hi);
            if (FULL) { o.xlo[m] = (lo - x[m]) / P.Tx[m]; o.xhi[m] = (hi - x[m]) / P.Tx[m]; }
        }
    }
    if (!term && !(FULL && hess_only)) {
        // input box rows: value is u itself (osqp_interface.cpp:274-276)
#pragma unroll
        for (int j = 0; j < NU; j++) gap += viol(u[j], P.lu[j], P.uu[j]);
        // joint-acceleration rows (osqp_interface.cpp:279-297)
#pragma unroll
        for (int j = 0; j < DOF; j++) {
            double lo, hi, c;
            if (k == 0) { c = 1. / Ts * u[j]; lo = P.ldd[j] + 1. / Ts * u_prev[j]; hi = P.udd[j] + 1. / Ts * u_prev[j]; }
            else { c = 1. / Ts * (u[j] - u_prev[j]); lo = P.ldd[j]; hi = P.udd[j]; }
            gap += viol(c, lo, hi);
            if (FULL) { o.dlo[j] = (lo - c) * Ts / P.Tu[j]; o.dhi[j] = (hi - c) * Ts / P.Tu[j]; }
        }
        // dynamics defect (osqp_interface.cpp:247), closed-form A_d, B_d (model.cpp:47-91)
#pragma unroll
        for (int m = 0; m < NX; m++) {
            double pred;
            if (m < 7) pred = x[m] + Ts * u[m];
            else if (m == 7) pred = x[7] + Ts * x[8] + 0.5 * Ts * Ts * u[7];
            else pred = x[8] + Ts * u[7];
            double c = (1.0 / P.Tx[m]) * (x_next[m] - pred);
            gap += fabs(c);
            if (FULL) o.b[m] = -c;
        }
        // polytopic rows (constraints.cpp:70-190)
#pragma unroll 1
        for (int j = 0; j < NPC; j++) {
            double h, sc;
            int goff;
            if (j == 0) { h = 0.01 * sel - P.tol_selcol * 0.01; sc = 0.01; goff = RB_DSEL; }
            else if (j == 1) { h = manip - P.tol_sing; sc = 1.0; goff = RB_DMANIP; }
            else {
                h = 0.01 * (rb(RB_ENV + j - 2) - rb(RB_OBSR) * 1.2) - 0.01 * P.tol_envcol;
                sc = 0.01; goff = RB_DENV + (j - 2) * DOF;
            }
            double dotp = 0;
#pragma unroll
            for (int m = 0; m < DOF; m++) {
                double g = sc * rb(goff + m);
                dotp += g * u[m];
                if (FULL) o.pg[j * DOF + m] = g;
            }
            const double rv = rbf_pre ? rbf_pre[j] : rbf(-0.5, h);
            double c = -dotp + rv;
            gap += fmax(c, 0.0);
            if (FULL) { o.pd[j] = rbf_pre ? rbf_pre[NPC + j] : drbf(-0.5, h); o.prhs[j] = -c; }
            if (rbf_out) { rbf_out[j] = rv; rbf_out[NPC + j] = drbf(-0.5, h); }
        }
    }
    o.gap = gap;
    if (!FULL) return;

    // ---- gradient and Gauss-Newton Hessian in x (cost.cpp:119-207) ----
    // d_total = [Jv, -T, 0]; d_lag = T T^T d_total + (T e^T + |e_l| I) [0.. N ..0] (s column); d_cont = d_total - d_lag
    const double eln = sqrt(el2);
    double Ms[3];  // (T e^T + |e_l| I) N
    {
        const double eN = e[0] * Nn[0] + e[1] * Nn[1] + e[2] * Nn[2];
#pragma unroll
        for (int a = 0; a < 3; a++) Ms[a] = Tg[a] * eN + eln * Nn[a];
    }
    // heading Jacobian: d_Log = J_r_inv R_ee^T [Jw, -dR_ref, 0]  (cost.cpp:182-191)
    double Mh[9];
    {
        double Jri[9];
        const double n = sqrt(lg2);
        if (n < 1e-8) { Jri[0] = Jri[4] = Jri[8] = 1.0; Jri[1] = Jri[2] = Jri[3] = Jri[5] = Jri[6] = Jri[7] = 0.0; }
        else {
            // "+" where the closed form has "-" (cost.cpp:188), reproduced
            const double coef = 1. / lg2 + (1. + cos(n)) / (2. * n * sin(n));
            const double xx = lg[0], yy = lg[1], zz = lg[2];
            Jri[0] = 1.0 + coef * (xx * xx - lg2); Jri[1] = -0.5 * zz + coef * (xx * yy);  Jri[2] = 0.5 * yy + coef * (xx * zz);
            Jri[3] = 0.5 * zz + coef * (xx * yy);  Jri[4] = 1.0 + coef * (yy * yy - lg2); Jri[5] = -0.5 * xx + coef * (yy * zz);
            Jri[6] = -0.5 * yy + coef * (xx * zz); Jri[7] = 0.5 * xx + coef * (yy * zz);  Jri[8] = 1.0 + coef * (zz * zz - lg2);
        }
        // Mh = J_r_inv * R_ee^T
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) Mh[3 * a + b] = Jri[3 * a] * Ree[3 * b] + Jri[3 * a + 1] * Ree[3 * b + 1] + Jri[3 * a + 2] * Ree[3 * b + 2];
    }
    // rows of the three 3x9 Jacobians, column by column (columns 0..6: joints, 7: s, 8: vs = 0)
    double dC[3][8], dL[3][8], dH[3][8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double dt[3], dw[3];
        if (c < 7) {
#pragma unroll
            for (int a = 0; a < 3; a++) { dt[a] = rb(RB_JV + 7 * a + c); dw[a] = rb(RB_JW + 7 * a + c); }
        } else {
#pragma unroll
            for (int a = 0; a < 3; a++) { dt[a] = -Tg[a]; dw[a] = -dRref[a]; }
        }
        const double Tdt = Tg[0] * dt[0] + Tg[1] * dt[1] + Tg[2] * dt[2];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double l = Tg[a] * Tdt + (c == 7 ? Ms[a] : 0.0);
            dL[a][c] = l;
            dC[a][c] = dt[a] - l;
            dH[a][c] = Mh[3 * a] * dw[0] + Mh[3 * a + 1] * dw[1] + Mh[3 * a + 2] * dw[2];
        }
    }
    double fx[NX];
    if (!hess_only) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double g = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) g += 2.0 * cc0 * dC[a][c] * ec[a] + 2.0 * wl * dL[a][c] * el[a] + 2.0 * wo * dH[a][c] * lg[a];
        if (c < 7) g += -P.q_sing * rb(RB_DMANIP + c);  // cost.cpp:280-283
        fx[c] = g;
    }
    fx[8] = 2.0 * P.q_vs * dv;
#pragma unroll
    for (int m = 0; m < NX; m++) o.q[m] = P.Tx[m] * fx[m];
    }
#pragma unroll
    for (int r_ = 0; r_ < NX; r_++)
#pragma unroll
        for (int c = 0; c <= r_; c++) {
            double h = 0;
            if (r_ < 8) {
#pragma unroll
                for (int a = 0; a < 3; a++) h += 2.0 * cc0 * dC[a][r_] * dC[a][c] + 2.0 * wl * dL[a][r_] * dL[a][c] + 2.0 * wo * dH[a][r_] * dH[a][c];
            }
            if (r_ == 8 && c == 8) h += 2.0 * P.q_vs;
            if (r_ == c) h += 1e-6;  // cost.cpp:353
            o.Q[sym9(r_, c)] = P.Tx[r_] * h * P.Tx[c];
        }
    // ---- input cost (cost.cpp:209-270) + joint-acceleration coupling (osqp_interface.cpp:167-217) ----
#pragma unroll
    for (int j = 0; j < NU; j++) { o.Rd[j] = 0; o.r[j] = 0; }
    if (!term) {
        const double kap = (k == 0 || k == N - 1) ? 2.0 : 4.0;
#pragma unroll
        for (int j = 0; j < DOF; j++) {
            double fuu = 2.0 * P.r_dq + 1e-6;
            double fu = 2.0 * P.r_dq * u[j];
            double gg;
            if (N == 1) gg = 0.0;  // i == 0 branch would read u[1]; with N == 1 there is no pair
            else if (k == 0) gg = 2. * P.r_ddq_solver * (u[j] - u_next[j]);
            else if (k == N - 1) gg = 2. * P.r_ddq_solver * (u[j] - u_prev[j]);
            else gg = 2. * P.r_ddq_solver * (2. * u[j] - u_next[j] - u_prev[j]);
            o.Rd[j] = P.Tu[j] * (fuu + kap * P.r_ddq_solver) * P.Tu[j];
            o.r[j] = P.Tu[j] * fu + P.Tu[j] * gg;
        }
        o.Rd[7] = P.Tu[7] * (2.0 * P.r_dVs + 1e-6) * P.Tu[7];
        o.r[7] = P.Tu[7] * (2.0 * P.r_dVs * u[7]);
    }
}

}  // namespace mpcc
