// Stage-structured convex QP solve replacing the reference's OSQP call
// (cpp/src/Interfaces/osqp_interface.cpp:592-656) for the QP that setCost / setConstraints
// (osqp_interface.cpp:129-389) assemble; see SURVEY.md Appendix A/B for the flat and stage forms.
//
//   variables   xi_k (9, k = 0..N, xi_0 = 0) and nu_k (8, k = 0..N-1): the reference's normalised steps
//   dynamics    xi_{k+1} = A xi_k + B nu_k + b_k                       (closed-form normalised A_d, B_d)
//   cost        sum 1/2 xi'Q xi + q'xi + 1/2 Rd nu^2 + r'nu + cpl_j nu_k[j] nu_{k+1}[j]
//   box         xlo <= xi_k <= xhi   (state rows intersected with the mis-indexed input rows, quirk 1)
//   rate        dlo <= nu_k[j] - nu_{k-1}[j] <= dhi
//   polytopic   ax' xi_k[0:7] + au' nu_k[0:7] <= prhs   (11 rows, k < N)
//
// Method: Mehrotra predictor-corrector interior point; every Newton system is solved exactly by a
// Riccati recursion on the state augmented with the previous joint-velocity step (16 states, 8
// inputs), one factorisation and two solves per iteration.  This file is the one-thread-per-instance
// formulation (also compiled for the host by the CPU-side unit tests); the warp-cooperative kernel
// in sqp_warp.cuh implements the same iteration.
#pragma once
#include "mpcc_types.h"
#include "dev_stage.cuh"

namespace mpcc {

// element e of a per-instance array lives at p[e * stride] (stride = batch size: coalesced across threads)
struct WsRef {
    double* p;
    size_t stride;
    MPCC_HD double& operator[](int e) const { return p[(size_t)e * stride]; }
    MPCC_HD WsRef off(int e) const { return WsRef{p + (size_t)e * stride, stride}; }
};

// ---- per-stage workspace layout (doubles) ----
constexpr int LIN_Q = 0, LIN_q = 45, LIN_RD = 54, LIN_r = 62, LIN_b = 70, LIN_XLO = 79, LIN_XHI = 88, LIN_DLO = 97, LIN_DHI = 104,
              LIN_PG = 111, LIN_PD = 188, LIN_PRHS = 199, LIN_OBJ = 210, LIN_GAP = 211, LIN_SIZE = 212;
constexpr int NBOX = 18, NRATE = 14, NPOLY = 11, NINEQ = NBOX + NRATE + NPOLY;  // 43 per stage
// constraint index inside a stage: box lower m (0..8), box upper 9+m, rate lower 18+j, rate upper 25+j, poly 32+j
constexpr int V_XI = 0, V_NU = 9, V_Y = 17, V_SIZE = 26;
constexpr int I_T = 0, I_LAM = NINEQ, I_RP = 2 * NINEQ, I_W = 3 * NINEQ, I_V = 4 * NINEQ, I_DT = 5 * NINEQ, I_DLAM = 6 * NINEQ, I_SIZE = 7 * NINEQ;
constexpr int S_DXI = 0, S_DNU = 9, S_GXI = 17, S_GNU = 26, S_KAP = 34, S_PX = 42, S_YN = 51, S_SIZE = 60;
constexpr int F_L = 0, F_LAM = 36, F_PXX = 164, F_PWX = 209, F_SIZE = 272;
constexpr int STAGE_WS = LIN_SIZE + V_SIZE + I_SIZE + S_SIZE + F_SIZE;
constexpr int OFF_LIN = 0, OFF_VAR = LIN_SIZE, OFF_INEQ = OFF_VAR + V_SIZE, OFF_STEP = OFF_INEQ + I_SIZE, OFF_FACT = OFF_STEP + S_SIZE;

MPCC_HD int qp_workspace_doubles(int N) { return (N + 1) * STAGE_WS; }

struct QpStats { int ok; int iters; double res_dual, res_prim, gap; int infeasible = 0; };  // infeasible: stopped by a primal-infeasibility certificate

// normalised dynamics constants
struct DynConst {
    double asv;      // A[s][vs]
    double bq[DOF];  // B[q_j][dq_j]
    double bs, bv;   // B[s][dVs], B[vs][dVs]
    double cpl[DOF]; // Hessian entry between nu_k[j] and nu_{k+1}[j]: -2 r_ddq Tu_j^2
};
MPCC_HD DynConst make_dyn(const Params& P, double Ts) {
    DynConst d;
    d.asv = Ts * P.Tx[8] / P.Tx[7];
#pragma unroll
    for (int j = 0; j < DOF; j++) { d.bq[j] = Ts * P.Tu[j] / P.Tx[j]; d.cpl[j] = -2.0 * P.r_ddq_solver * P.Tu[j] * P.Tu[j]; }
    d.bs = 0.5 * Ts * Ts * P.Tu[7] / P.Tx[7];
    d.bv = Ts * P.Tu[7] / P.Tx[8];
    return d;
}

struct StageQP {
    const Params& P;
    DynConst dyn;
    int N;
    WsRef ws;
    QpOptions opt;

    MPCC_HD WsRef lin(int k) const { return ws.off(k * STAGE_WS + OFF_LIN); }
    MPCC_HD WsRef var(int k) const { return ws.off(k * STAGE_WS + OFF_VAR); }
    MPCC_HD WsRef ineq(int k) const { return ws.off(k * STAGE_WS + OFF_INEQ); }
    MPCC_HD WsRef stp(int k) const { return ws.off(k * STAGE_WS + OFF_STEP); }
    MPCC_HD WsRef fac(int k) const { return ws.off(k * STAGE_WS + OFF_FACT); }

    // polytopic row j of stage k in normalised variables
    MPCC_HD void poly_row(const WsRef& L, int j, double* ax, double* au) const {
        double pd = L[LIN_PD + j];
#pragma unroll
        for (int m = 0; m < DOF; m++) {
            double g = L[LIN_PG + j * DOF + m];
            ax[m] = pd * g * P.Tx[m];
            au[m] = -g * P.Tu[m];
        }
    }
    // is constraint c of stage k present?  box: k >= 1; rate, poly: k < N
    MPCC_HD bool present(int k, int c) const { return (c < NBOX) ? (k >= 1) : (k < N); }

    // G z for constraint c of stage k evaluated on a (xi, nu) pair of vectors stored at offsets ox / on
    // of the VAR or STEP blocks (sign convention: G z <= h)
    template <int OX, int ON, bool STEP>
    MPCC_HD double gz(int k, int c) const {
        WsRef v = STEP ? stp(k) : var(k);
        if (c < 9) return -v[OX + c];
        if (c < 18) return v[OX + c - 9];
        if (c < 32) {
            int j = (c < 25) ? c - 18 : c - 25;
            double d = v[ON + j];
            if (k >= 1) { WsRef vp = STEP ? stp(k - 1) : var(k - 1); d -= vp[ON + j]; }
            return (c < 25) ? -d : d;
        }
        double ax[DOF], au[DOF];
        poly_row(lin(k), c - 32, ax, au);
        double s = 0;
#pragma unroll
        for (int m = 0; m < DOF; m++) s += ax[m] * v[OX + m] + au[m] * v[ON + m];
        return s;
    }
    MPCC_HD double hval(int k, int c) const {
        WsRef L = lin(k);
        if (c < 9) return -L[LIN_XLO + c];
        if (c < 18) return L[LIN_XHI + c - 9];
        if (c < 25) return -L[LIN_DLO + c - 18];
        if (c < 32) return L[LIN_DHI + c - 25];
        return L[LIN_PRHS + c - 32];
    }

    // Apply the mis-indexed input-bound rows (osqp_interface.cpp:273): row (i,kk) bounds
    // T_u[kk] * z[8 i + kk], a state-step entry, by [lu - u_i[kk], uu - u_i[kk]].
    // guess_u(i, kk) returns the current iterate's input.  Intersected into the state box.
    template <class GU>
    MPCC_HD void apply_input_bound_quirk(const GU& guess_u) const {
        for (int c = 0; c < NU * N; c++) {
            int k = c / NX, m = c % NX, i = c / NU, kk = c % NU;
            double uv = guess_u(i, kk);
            double lo = (P.lu[kk] - uv) / P.Tu[kk], hi = (P.uu[kk] - uv) / P.Tu[kk];
            WsRef L = lin(k);
            L[LIN_XLO + m] = fmax(L[LIN_XLO + m], lo);
            L[LIN_XHI + m] = fmin(L[LIN_XHI + m], hi);
        }
    }

    // ---- Riccati factorisation (matrices only) --------------------------------------------------
    // returns false if some M_nunu is not positive definite
    MPCC_HDN bool factor() const {
        double Pxx[81], Pwx[63], Pww[49];  // cost-to-go of stage k+1 (dense, symmetric where applicable)
        // terminal stage: P_N = Q_N + box W
        {
            WsRef L = lin(N), I = ineq(N);
            for (int r = 0; r < 9; r++)
                for (int c = 0; c <= r; c++) { double v = L[LIN_Q + sym9(r, c)]; Pxx[9 * r + c] = v; Pxx[9 * c + r] = v; }
            for (int m = 0; m < 9; m++) Pxx[10 * m] += I[I_W + m] + I[I_W + 9 + m];
            for (int i = 0; i < 63; i++) Pwx[i] = 0;
            for (int i = 0; i < 49; i++) Pww[i] = 0;
        }
        for (int k = N - 1; k >= 0; k--) {
            WsRef L = lin(k), I = ineq(k), F = fac(k);
            // store P_{k+1} blocks needed for the costates of the forward pass
            {
                WsRef Fn = fac(k + 1);
                for (int r = 0; r < 9; r++) for (int c = 0; c <= r; c++) Fn[F_PXX + sym9(r, c)] = Pxx[9 * r + c];
                for (int i = 0; i < 63; i++) Fn[F_PWX + i] = Pwx[i];
            }
            double Mnn[64], Mnx[8 * 16], Mxx[81], Mww[7];
            // FF = B'Pxx + E'Pwx  (8x9)
            double FF[72];
            for (int c = 0; c < 9; c++) {
                for (int j = 0; j < 7; j++) FF[9 * j + c] = dyn.bq[j] * Pxx[9 * j + c] + Pwx[9 * j + c];
                FF[9 * 7 + c] = dyn.bs * Pxx[9 * 7 + c] + dyn.bv * Pxx[9 * 8 + c];
            }
            // Mnn = FF B + (B'Pxw + Pww) E + diag(Rd) + rate W + poly
            for (int i = 0; i < 8; i++) {
                for (int j = 0; j < 7; j++) {
                    double v = dyn.bq[j] * FF[9 * i + j];
                    if (i < 7) v += dyn.bq[i] * Pwx[9 * j + i] + Pww[7 * i + j];
                    else v += dyn.bs * Pwx[9 * j + 7] + dyn.bv * Pwx[9 * j + 8];
                    Mnn[8 * i + j] = v;
                }
                Mnn[8 * i + 7] = dyn.bs * FF[9 * i + 7] + dyn.bv * FF[9 * i + 8];
            }
            for (int j = 0; j < 8; j++) Mnn[9 * j] += L[LIN_RD + j];
            for (int j = 0; j < 7; j++) Mnn[9 * j] += I[I_W + 18 + j] + I[I_W + 25 + j];
            // Mnx = [FF A | Mnw]  (8 x 16)
            for (int i = 0; i < 8; i++) {
                for (int c = 0; c < 9; c++) Mnx[16 * i + c] = FF[9 * i + c];
                Mnx[16 * i + 8] += dyn.asv * FF[9 * i + 7];
                for (int j = 0; j < 7; j++) Mnx[16 * i + 9 + j] = 0;
            }
            for (int j = 0; j < 7; j++) {
                double wr = (k >= 1) ? (I[I_W + 18 + j] + I[I_W + 25 + j]) : 0.0;
                Mww[j] = wr;
                Mnx[16 * j + 9 + j] = (k >= 1) ? (dyn.cpl[j] - wr) : 0.0;
            }
            // Mxx = Q + A'Pxx A + box W + poly
            for (int r = 0; r < 9; r++)
                for (int c = 0; c < 9; c++) {
                    double v = Pxx[9 * r + c];
                    if (c == 8) v += dyn.asv * Pxx[9 * r + 7];
                    Mxx[9 * r + c] = v;
                }
            for (int c = 0; c < 9; c++) Mxx[9 * 8 + c] += dyn.asv * Mxx[9 * 7 + c];
            for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) Mxx[9 * r + c] += L[LIN_Q + ((r >= c) ? sym9(r, c) : sym9(c, r))];
            if (k >= 1) for (int m = 0; m < 9; m++) Mxx[10 * m] += I[I_W + m] + I[I_W + 9 + m];
            // polytopic rank-1 terms
            for (int j = 0; j < NPOLY; j++) {
                double w = I[I_W + 32 + j];
                double ax[DOF], au[DOF];
                poly_row(L, j, ax, au);
                for (int a = 0; a < 7; a++) {
                    double wa = w * au[a], wx = w * ax[a];
                    for (int b2 = 0; b2 < 7; b2++) {
                        Mnn[8 * a + b2] += wa * au[b2];
                        Mnx[16 * a + b2] += wa * ax[b2];
                        Mxx[9 * a + b2] += wx * ax[b2];
                    }
                }
            }
            // Cholesky Mnn = L L'
            double Lm[64];
            for (int j = 0; j < 8; j++) {
                double d = Mnn[9 * j];
                for (int t = 0; t < j; t++) d -= Lm[8 * j + t] * Lm[8 * j + t];
                if (!(d > 0.0)) return false;
                d = sqrt(d);
                Lm[9 * j] = d;
                double inv = 1.0 / d;
                for (int i = j + 1; i < 8; i++) {
                    double s = Mnn[8 * i + j];
                    for (int t = 0; t < j; t++) s -= Lm[8 * i + t] * Lm[8 * j + t];
                    Lm[8 * i + j] = s * inv;
                }
            }
            {
                int q = 0;
                for (int i = 0; i < 8; i++) for (int j = 0; j <= i; j++) F[F_L + q++] = Lm[8 * i + j];
            }
            // Lam = L^-1 Mnx  (8 x 16), in place
            for (int c = 0; c < 16; c++)
                for (int i = 0; i < 8; i++) {
                    double s = Mnx[16 * i + c];
                    for (int t = 0; t < i; t++) s -= Lm[8 * i + t] * Mnx[16 * t + c];
                    Mnx[16 * i + c] = s / Lm[9 * i];
                }
            for (int i = 0; i < 128; i++) F[F_LAM + i] = Mnx[i];
            // P_k = [Mxx 0; 0 Mww] - Lam'Lam
            for (int r = 0; r < 9; r++)
                for (int c = 0; c < 9; c++) {
                    double s = Mxx[9 * r + c];
                    for (int t = 0; t < 8; t++) s -= Mnx[16 * t + r] * Mnx[16 * t + c];
                    Pxx[9 * r + c] = s;
                }
            for (int a = 0; a < 7; a++) {
                for (int c = 0; c < 9; c++) {
                    double s = 0;
                    for (int t = 0; t < 8; t++) s -= Mnx[16 * t + 9 + a] * Mnx[16 * t + c];
                    Pwx[9 * a + c] = s;
                }
                for (int b2 = 0; b2 < 7; b2++) {
                    double s = (a == b2) ? Mww[a] : 0.0;
                    for (int t = 0; t < 8; t++) s -= Mnx[16 * t + 9 + a] * Mnx[16 * t + 9 + b2];
                    Pww[7 * a + b2] = s;
                }
            }
        }
        return true;
    }

    // ---- gradient of the step QP from the per-constraint scalars v (see file header) ------------
    MPCC_HDN void build_gradient() const {
        for (int k = 0; k <= N; k++) {
            WsRef L = lin(k), V = var(k), I = ineq(k), S = stp(k);
            double xi[9], g[9];
            for (int m = 0; m < 9; m++) xi[m] = V[V_XI + m];
            for (int r = 0; r < 9; r++) {
                double s = L[LIN_q + r];
                for (int c = 0; c < 9; c++) s += L[LIN_Q + ((r >= c) ? sym9(r, c) : sym9(c, r))] * xi[c];
                if (k >= 1) s += I[I_V + 9 + r] - I[I_V + r];
                g[r] = s;
            }
            if (k < N) {
                double gn[8];
                for (int j = 0; j < 8; j++) gn[j] = L[LIN_RD + j] * V[V_NU + j] + L[LIN_r + j];
                for (int j = 0; j < 7; j++) {
                    if (k >= 1) gn[j] += dyn.cpl[j] * var(k - 1)[V_NU + j];
                    if (k <= N - 2) gn[j] += dyn.cpl[j] * var(k + 1)[V_NU + j];
                    gn[j] += I[I_V + 25 + j] - I[I_V + 18 + j];
                    if (k + 1 <= N - 1) { WsRef In = ineq(k + 1); gn[j] -= In[I_V + 25 + j] - In[I_V + 18 + j]; }
                }
                for (int j = 0; j < NPOLY; j++) {
                    double v = I[I_V + 32 + j];
                    double ax[DOF], au[DOF];
                    poly_row(L, j, ax, au);
                    for (int m = 0; m < 7; m++) { g[m] += v * ax[m]; gn[m] += v * au[m]; }
                }
                for (int j = 0; j < 8; j++) S[S_GNU + j] = gn[j];
            }
            for (int m = 0; m < 9; m++) S[S_GXI + m] = g[m];
        }
    }

    // ---- Riccati vector passes: solves the Newton system for the current gradient ----------------
    // writes the step into S_DXI / S_DNU and the new costates into S_YN
    MPCC_HDN void solve_step() const {
        double px[9], pw[7];
        {
            WsRef S = stp(N);
            for (int m = 0; m < 9; m++) { px[m] = S[S_GXI + m]; S[S_PX + m] = px[m]; }
            for (int j = 0; j < 7; j++) pw[j] = 0;
        }
        for (int k = N - 1; k >= 0; k--) {
            WsRef S = stp(k), F = fac(k);
            double mn[8], mx[16];
            for (int j = 0; j < 7; j++) mn[j] = S[S_GNU + j] + dyn.bq[j] * px[j] + pw[j];
            mn[7] = S[S_GNU + 7] + dyn.bs * px[7] + dyn.bv * px[8];
            for (int c = 0; c < 9; c++) mx[c] = S[S_GXI + c] + px[c];
            mx[8] += dyn.asv * px[7];
            for (int j = 0; j < 7; j++) mx[9 + j] = 0;
            // kappa = L^-1 mn
            double kap[8];
            {
                int q = 0;
                for (int i = 0; i < 8; i++) {
                    double s = mn[i];
                    for (int t = 0; t < i; t++) s -= F[F_L + q++] * kap[t];
                    kap[i] = s / F[F_L + q++];
                }
            }
            for (int i = 0; i < 8; i++) S[S_KAP + i] = kap[i];
            for (int c = 0; c < 16; c++) {
                double s = mx[c];
                for (int t = 0; t < 8; t++) s -= F[F_LAM + 16 * t + c] * kap[t];
                mx[c] = s;
            }
            for (int c = 0; c < 9; c++) { px[c] = mx[c]; S[S_PX + c] = px[c]; }
            for (int j = 0; j < 7; j++) pw[j] = mx[9 + j];
        }
        // forward
        double dx[9], dw[7];
        for (int m = 0; m < 9; m++) dx[m] = 0;
        for (int j = 0; j < 7; j++) dw[j] = 0;
        for (int m = 0; m < 9; m++) { stp(0)[S_DXI + m] = 0; stp(0)[S_YN + m] = 0; }
        for (int k = 0; k < N; k++) {
            WsRef S = stp(k), F = fac(k);
            double rhs[8], dn[8];
            for (int i = 0; i < 8; i++) {
                double s = S[S_KAP + i];
                for (int c = 0; c < 9; c++) s += F[F_LAM + 16 * i + c] * dx[c];
                for (int j = 0; j < 7; j++) s += F[F_LAM + 16 * i + 9 + j] * dw[j];
                rhs[i] = -s;
            }
            // dn = L^-T rhs
            for (int i = 7; i >= 0; i--) {
                double s = rhs[i];
                for (int t = i + 1; t < 8; t++) s -= F[F_L + t * (t + 1) / 2 + i] * dn[t];
                dn[i] = s / F[F_L + i * (i + 1) / 2 + i];
            }
            for (int j = 0; j < 8; j++) S[S_DNU + j] = dn[j];
            double nx[9];
            for (int j = 0; j < 7; j++) nx[j] = dx[j] + dyn.bq[j] * dn[j];
            nx[7] = dx[7] + dyn.asv * dx[8] + dyn.bs * dn[7];
            nx[8] = dx[8] + dyn.bv * dn[7];
            for (int m = 0; m < 9; m++) dx[m] = nx[m];
            for (int j = 0; j < 7; j++) dw[j] = dn[j];
            WsRef Sn = stp(k + 1), Fn = fac(k + 1);
            for (int m = 0; m < 9; m++) Sn[S_DXI + m] = dx[m];
            // multiplier of the dynamics row k+1: y_{k+1} = -dV_{k+1}/dxi = -(Pxx dxi + Pwx' dw + px)
            for (int r = 0; r < 9; r++) {
                double s = Sn[S_PX + r];
                for (int c = 0; c < 9; c++) s += Fn[F_PXX + ((r >= c) ? sym9(r, c) : sym9(c, r))] * dx[c];
                if (k + 1 < N) for (int j = 0; j < 7; j++) s += Fn[F_PWX + 9 * j + r] * dw[j];
                Sn[S_YN + r] = -s;
            }
        }
    }

    // slack / multiplier steps from the primal step; returns the largest step keeping t, lam > 0
    MPCC_HDN double ineq_steps() const {
        double a = 1.0;
        for (int k = 0; k <= N; k++) {
            WsRef I = ineq(k);
            for (int c = 0; c < NINEQ; c++) {
                if (!present(k, c)) continue;
                double g = gz<S_DXI, S_DNU, true>(k, c);
                double dt = -I[I_RP + c] - g;
                double dl = -I[I_LAM + c] + I[I_V + c] + I[I_W + c] * g;
                I[I_DT + c] = dt;
                I[I_DLAM + c] = dl;
                if (dt < 0) a = fmin(a, -I[I_T + c] / dt);
                if (dl < 0) a = fmin(a, -I[I_LAM + c] / dl);
            }
        }
        return a;
    }

    // ---- the interior-point loop ------------------------------------------------------------------
    MPCC_HDN QpStats solve() const {
        QpStats st;
        st.ok = 0; st.iters = 0; st.res_dual = 0; st.res_prim = 0; st.gap = 0;
        // constant rows of stage 0 (xi_0 = 0): feasible iff 0 is inside the box
        {
            WsRef L = lin(0);
            for (int m = 0; m < 9; m++)
                if (L[LIN_XLO + m] > 1e-9 || L[LIN_XHI + m] < -1e-9) return st;
        }
        for (int k = 1; k <= N; k++) {
            WsRef L = lin(k);
            for (int m = 0; m < 9; m++) if (L[LIN_XLO + m] > L[LIN_XHI + m]) return st;
        }
        // initial point: nu = 0, xi = rollout of the defects, y = 0, t = max(h - Gz, QP_INIT_SLACK), lam = QP_INIT_SLACK / t
        double qn = 0;
        {
            double x[9];
            for (int m = 0; m < 9; m++) x[m] = 0;
            for (int k = 0; k <= N; k++) {
                WsRef V = var(k), L = lin(k);
                for (int m = 0; m < 9; m++) { V[V_XI + m] = x[m]; V[V_Y + m] = 0; qn = fmax(qn, fabs(L[LIN_q + m])); }
                if (k < N) {
                    for (int j = 0; j < 8; j++) { V[V_NU + j] = 0; qn = fmax(qn, fabs(L[LIN_r + j])); }
                    double nx[9];
                    for (int j = 0; j < 7; j++) nx[j] = x[j] + L[LIN_b + j];
                    nx[7] = x[7] + dyn.asv * x[8] + L[LIN_b + 7];
                    nx[8] = x[8] + L[LIN_b + 8];
                    for (int m = 0; m < 9; m++) x[m] = nx[m];
                }
            }
        }
        int m_tot = 0;
        for (int k = 0; k <= N; k++) {
            WsRef I = ineq(k);
            for (int c = 0; c < NINEQ; c++) {
                if (!present(k, c)) { I[I_T + c] = 1; I[I_LAM + c] = 0; I[I_W + c] = 0; I[I_V + c] = 0; I[I_RP + c] = 0; continue; }
                const double t0 = fmax(hval(k, c) - gz<V_XI, V_NU, false>(k, c), QP_INIT_SLACK);
                I[I_T + c] = t0;
                I[I_LAM + c] = QP_INIT_SLACK / t0;
                m_tot++;
            }
        }
        for (int it = 0; it < opt.max_iter; it++) {
            // ---- residuals ----
            double mu = 0, nrp = 0, nrd = 0;
            for (int k = 0; k <= N; k++) {
                WsRef I = ineq(k);
                for (int c = 0; c < NINEQ; c++) {
                    if (!present(k, c)) continue;
                    double rp = gz<V_XI, V_NU, false>(k, c) + I[I_T + c] - hval(k, c);
                    I[I_RP + c] = rp;
                    nrp = fmax(nrp, fabs(rp));
                    mu += I[I_T + c] * I[I_LAM + c];
                    I[I_W + c] = I[I_LAM + c] / I[I_T + c];
                    I[I_V + c] = I[I_LAM + c];  // v = lam -> build_gradient gives the plain Lagrangian gradient
                }
            }
            mu /= m_tot;
            build_gradient();
            for (int k = 0; k <= N; k++) {
                WsRef S = stp(k), V = var(k);
                if (k >= 1) {
                    for (int m = 0; m < 9; m++) {
                        double r = S[S_GXI + m] + V[V_Y + m];
                        if (k < N) { r -= var(k + 1)[V_Y + m]; if (m == 8) r -= dyn.asv * var(k + 1)[V_Y + 7]; }
                        nrd = fmax(nrd, fabs(r));
                    }
                }
                if (k < N) {
                    WsRef Vn = var(k + 1);
                    for (int j = 0; j < 7; j++) nrd = fmax(nrd, fabs(S[S_GNU + j] - dyn.bq[j] * Vn[V_Y + j]));
                    nrd = fmax(nrd, fabs(S[S_GNU + 7] - dyn.bs * Vn[V_Y + 7] - dyn.bv * Vn[V_Y + 8]));
                }
            }
            st.iters = it; st.res_dual = nrd; st.res_prim = nrp; st.gap = mu;
            if (nrd <= opt.eps * (1.0 + qn) && nrp <= opt.eps && mu <= opt.eps) { st.ok = 1; break; }
            if (!(nrd == nrd) || !(mu == mu)) break;
            if (!factor()) break;
            // ---- predictor: v = lam rp / t ----
            for (int k = 0; k <= N; k++) {
                WsRef I = ineq(k);
                for (int c = 0; c < NINEQ; c++) if (present(k, c)) I[I_V + c] = I[I_LAM + c] * I[I_RP + c] / I[I_T + c];
            }
            build_gradient();
            solve_step();
            double a_aff = ineq_steps();
            double mu_aff = 0;
            for (int k = 0; k <= N; k++) {
                WsRef I = ineq(k);
                for (int c = 0; c < NINEQ; c++) if (present(k, c)) mu_aff += (I[I_T + c] + a_aff * I[I_DT + c]) * (I[I_LAM + c] + a_aff * I[I_DLAM + c]);
            }
            mu_aff /= m_tot;
            double sigma = (mu > 0) ? (mu_aff / mu) * (mu_aff / mu) * (mu_aff / mu) : 0.0;
            // ---- corrector: v = (lam rp + sigma mu - dt_a dlam_a) / t ----
            for (int k = 0; k <= N; k++) {
                WsRef I = ineq(k);
                for (int c = 0; c < NINEQ; c++)
                    if (present(k, c)) I[I_V + c] = (I[I_LAM + c] * I[I_RP + c] + sigma * mu - I[I_DT + c] * I[I_DLAM + c]) / I[I_T + c];
            }
            build_gradient();
            solve_step();
            double a = fmin(1.0, qp_step_tau(mu) * ineq_steps());
            for (int k = 0; k <= N; k++) {
                WsRef V = var(k), S = stp(k), I = ineq(k);
                for (int m = 0; m < 9; m++) { V[V_XI + m] += a * S[S_DXI + m]; V[V_Y + m] += a * (S[S_YN + m] - V[V_Y + m]); }
                if (k < N) for (int j = 0; j < 8; j++) V[V_NU + j] += a * S[S_DNU + j];
                for (int c = 0; c < NINEQ; c++) if (present(k, c)) { I[I_T + c] += a * I[I_DT + c]; I[I_LAM + c] += a * I[I_DLAM + c]; }
            }
            st.iters = it + 1;
        }
        if (st.ok) {
            // make the equalities exact: xi = rollout(nu)
            double x[9];
            for (int m = 0; m < 9; m++) x[m] = 0;
            for (int k = 0; k <= N; k++) {
                WsRef V = var(k), L = lin(k);
                for (int m = 0; m < 9; m++) V[V_XI + m] = x[m];
                if (k < N) {
                    double nx[9];
                    for (int j = 0; j < 7; j++) nx[j] = x[j] + dyn.bq[j] * V[V_NU + j] + L[LIN_b + j];
                    nx[7] = x[7] + dyn.asv * x[8] + dyn.bs * V[V_NU + 7] + L[LIN_b + 7];
                    nx[8] = x[8] + dyn.bv * V[V_NU + 7] + L[LIN_b + 8];
                    for (int m = 0; m < 9; m++) x[m] = nx[m];
                }
            }
        }
        return st;
    }
};

}  // namespace mpcc
