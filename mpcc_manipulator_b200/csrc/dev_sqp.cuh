// One control cycle of one MPCC instance: prologue (reference cpp/src/MPC/mpc.cpp:104-124,
// 54-89) and the SQP loop (cpp/src/Interfaces/osqp_interface.cpp:398-590) with the filter
// line search (:759-808).  One-thread-per-instance formulation; also compiled for the host by
// the CPU-side unit tests.  The horizon ("guess") is stored as (N+1) x 17 = [x(9), u(8)] per stage.
#pragma once
#include "mpcc_types.h"
#include "dev_panda.cuh"
#include "dev_track.cuh"
#include "dev_stage.cuh"
#include "dev_qp.cuh"

namespace mpcc {

constexpr int HZ = NX + NU;  // doubles per stage of a horizon

// per-instance persistent state of mpcc::MPC (mpc.h:117-127): warm start + validity + failure counter
struct WarmFlags { int32_t valid; int32_t failed; };

// ---- prologue: projection, vs estimate, warm-start shift / regeneration --------------------------
// x0 (9) is updated in place (s, vs) like the reference (mpc.cpp:108,115).
MPCC_HDN void cycle_prologue(const Params& P, const TrackTable& T, double Ts, int N, double* x0, const double* u0,
                             const WsRef& warm, WarmFlags& fl) {
    PandaKin kin;
    panda_kinematics(x0, kin);
    const double last_s = x0[7];
    x0[7] = track_project(T, P.max_dist_proj, last_s, kin.p);
    TrackPoint tp;
    track_eval_pos(T, x0[7], tp);
    double vs = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double v = 0;
#pragma unroll
        for (int j = 0; j < DOF; j++) v += kin.Jv[7 * a + j] * u0[j];
        vs += v * tp.dpos[a];
    }
    x0[8] = vs;
    if (fabs(last_s - x0[7]) > P.max_dist_proj) { fl.valid = 0; fl.failed++; }
    const double L = T.s[N_SPLINE - 1];
    // unwrapInitialGuess (mpc.cpp:70-77) clamps s of the stages 1..N to the track length: applied as the stages are written
    if (fl.valid) {
        // updateInitialGuess (mpc.cpp:54-68): g[i-1] = g[i] for i = 1..N-1, g[0].xk = x0, g[N-1] = g[N-2], g[N].xk = RK4(g[N-1]), g[N].uk = 0.
        // One thread shifts one instance's warm start through global memory: the loads of stage i + 1 are issued BEFORE the stores of stage i
        // (same array: the compiler must otherwise keep every load behind the previous store -- 340 dependent memory round trips, 0.1 ms).
        double cur[HZ], nxt[HZ];
#pragma unroll
        for (int e = 0; e < HZ; e++) cur[e] = warm[HZ + e];
        for (int i = 1; i < N; i++) {
            if (i + 1 < N) {
#pragma unroll
                for (int e = 0; e < HZ; e++) nxt[e] = warm[(i + 1) * HZ + e];
            }
#pragma unroll
            for (int e = 0; e < HZ; e++) {
                double v = cur[e];
                if (i == 1 && e < NX) v = x0[e];               // g[0].xk = x0 (keeps the shifted uk)
                else if (e == 7 && i > 1) v = fmin(v, L);
                warm[(i - 1) * HZ + e] = v;
            }
            if (i + 1 < N) {
#pragma unroll
                for (int e = 0; e < HZ; e++) cur[e] = nxt[e];
            }
        }
        // cur = the old last stage g[N-1], now also at N-2: g[N-1] = g[N-2] leaves it where it is (only the clamp applies)
        double xs[NX], us[NU];
        for (int e = 0; e < NX; e++) xs[e] = (N == 2 && e < NX) ? x0[e] : cur[e];
        for (int e = 0; e < NU; e++) us[e] = cur[NX + e];
        if (N == 2) { for (int e = 0; e < NX; e++) warm[HZ + e] = x0[e]; }   // N = 2: g[N-2] is g[0], whose xk is x0
        warm[(N - 1) * HZ + 7] = fmin(xs[7], L);
        // RK4 of the linear model is exact (integrator.cpp:29-43 on model.cpp:31-45); it reads the UNclamped g[N-1] (the clamp comes after it in the reference)
        for (int j = 0; j < DOF; j++) warm[N * HZ + j] = xs[j] + Ts * us[j];
        warm[N * HZ + 7] = fmin(xs[7] + Ts * xs[8] + 0.5 * Ts * Ts * us[7], L);
        warm[N * HZ + 8] = xs[8] + Ts * us[7];
        for (int e = 0; e < NU; e++) warm[N * HZ + NX + e] = 0;
    } else {
        // generateNewInitialGuess (mpc.cpp:79-89)
        for (int i = 0; i <= N; i++) {
            for (int e = 0; e < NX; e++) warm[i * HZ + e] = (e == 7 && i > 0) ? fmin(x0[e], L) : x0[e];
            for (int e = 0; e < NU; e++) warm[i * HZ + NX + e] = 0;
        }
        fl.valid = 1;
    }
}

// accept_mask: bit i = the filter accepted the first line-search trial of SQP iteration i (i < 32)
struct SqpResult { int32_t status; int32_t iters; int32_t qp_fail; int32_t qp_iters; uint32_t accept_mask; };

// optional per-iteration record for the parity tests (step of each SQP iteration, alpha, QP flag)
struct SqpLogRef {
    double* steps;   // [max_log][ (N+1)*17 ] normalised step in horizon layout, or nullptr
    double* alphas;
    int32_t* qp_ok;
    int max_log;
    int n;
};

// Evaluate every stage of a horizon.  FULL: fill the QP blocks; else only objective and violation.
template <bool FULL>
MPCC_HDN void eval_horizon(const Params& P, const TrackTable& T, double Ts, int N, const WsRef& g, const double* cur_u,
                           const double* rb, size_t rb_stride, size_t rb_stage, const StageQP* qp, double& obj, double& gap) {
    obj = 0; gap = 0;
    for (int k = 0; k <= N; k++) {
        double x[NX], u[NU], up[DOF], un[DOF], xn[NX];
        for (int e = 0; e < NX; e++) x[e] = g[k * HZ + e];
        for (int e = 0; e < NU; e++) u[e] = g[k * HZ + NX + e];
        for (int j = 0; j < DOF; j++) {
            up[j] = (k == 0) ? cur_u[j] : g[(k - 1) * HZ + NX + j];
            un[j] = (k < N) ? g[(k + 1) * HZ + NX + j] : 0.0;
        }
        for (int e = 0; e < NX; e++) xn[e] = (k < N) ? g[(k + 1) * HZ + e] : 0.0;
        StageLin sl;
        RbView rv{rb + (size_t)k * rb_stage, rb_stride};
        stage_eval<FULL>(P, T, Ts, N, k, x, u, up, un, xn, rv, sl);
        obj += sl.obj;
        gap += sl.gap;
        if (FULL) {
            WsRef L = qp->lin(k);
            const double* src = (const double*)&sl;
            for (int e = 0; e < LIN_SIZE; e++) L[e] = src[e];
        }
    }
}

// The SQP loop.  guess: current horizon (in/out).  step: persistent normalised step (horizon layout).
// trial: scratch horizon.  filt: 2*(max_iter+1) doubles.  Returns status / iteration count.
MPCC_HDN SqpResult sqp_solve(const Params& P, const TrackTable& T, double Ts, int N, const WsRef& guess, const WsRef& step,
                             const WsRef& trial, const WsRef& filt, const double* cur_u, const double* rb, size_t rb_stride,
                             size_t rb_stage, const WsRef& ws, QpOptions opt, SqpLogRef* log) {
    SqpResult res;
    res.status = SOLVED; res.iters = 0; res.qp_fail = 0; res.qp_iters = 0; res.accept_mask = 0;
    StageQP qp{P, make_dyn(P, Ts), N, ws, opt};
    const int max_iter = (int)P.max_iter, ls_max = (int)P.line_search_max_iter;
    for (int e = 0; e < (N + 1) * HZ; e++) step[e] = 0;
    int n_filt = 0;
    int it = 0;
    bool done = false;
    for (it = 0; it < max_iter; it++) {
        double obj, gap;
        eval_horizon<true>(P, T, Ts, N, guess, cur_u, rb, rb_stride, rb_stage, &qp, obj, gap);
        struct GU { const WsRef& g; MPCC_HD double operator()(int i, int kk) const { return g[i * HZ + NX + kk]; } } gu{guess};
        qp.apply_input_bound_quirk(gu);
        // isPosdef / isNan on the block structure of Hess_ (osqp_interface.cpp:454-473)
        {
            bool pd = true, nan = false;
            for (int k = 0; k <= N && pd; k++) {
                WsRef L = qp.lin(k);
                double A[81];
                for (int r = 0; r < 9; r++) for (int c = 0; c <= r; c++) { A[9 * r + c] = L[LIN_Q + sym9(r, c)]; if (A[9 * r + c] != A[9 * r + c]) nan = true; }
                for (int j = 0; j < 9 && pd; j++) {
                    double d = A[10 * j];
                    for (int t = 0; t < j; t++) d -= A[9 * j + t] * A[9 * j + t];
                    if (d <= 0.0) { pd = false; break; }
                    d = sqrt(d);
                    A[10 * j] = d;
                    for (int i = j + 1; i < 9; i++) {
                        double s = A[9 * i + j];
                        for (int t = 0; t < j; t++) s -= A[9 * i + t] * A[9 * j + t];
                        A[9 * i + j] = s / d;
                    }
                }
            }
            // input block: per joint a tridiagonal (Rd, cpl) chain; dVs diagonal
            for (int j = 0; j < NU && pd; j++) {
                double d = 0;
                for (int k = 0; k < N; k++) {
                    double rd = qp.lin(k)[LIN_RD + j];
                    if (rd != rd) nan = true;
                    d = (k == 0 || j == 7) ? rd : rd - qp.dyn.cpl[j] * qp.dyn.cpl[j] / d;
                    if (d <= 0.0) { pd = false; break; }
                }
            }
            if (!pd) { res.status = NON_PD_HESSIAN; done = true; break; }
            if (nan) { res.status = NAN_HESSIAN; done = true; break; }
        }
        QpStats qs = qp.solve();
        res.qp_iters += qs.iters;
        if (qs.ok) {
            for (int k = 0; k <= N; k++) {
                WsRef V = qp.var(k);
                for (int m = 0; m < NX; m++) step[k * HZ + m] = V[V_XI + m];
                for (int j = 0; j < NU; j++) step[k * HZ + NX + j] = (k < N) ? V[V_NU + j] : 0.0;
            }
        } else {
            res.qp_fail++;  // step keeps its previous value (osqp_interface.cpp:479-505)
        }
        // ---- filterLineSearch (osqp_interface.cpp:759-808) ----
        double alpha = 1.0;
        bool accepted = true;  // never reset inside the loop (:767)
        for (int i = 0; i < ls_max; i++) {
            for (int k = 0; k <= N; k++) {
                for (int m = 0; m < NX; m++) trial[k * HZ + m] = guess[k * HZ + m] + alpha * (P.Tx[m] * step[k * HZ + m]);
                for (int j = 0; j < NU; j++) trial[k * HZ + NX + j] = (k < N) ? guess[k * HZ + NX + j] + alpha * (P.Tu[j] * step[k * HZ + NX + j]) : 0.0;
            }
            double o2, g2;
            eval_horizon<false>(P, T, Ts, N, trial, cur_u, rb, rb_stride, rb_stage, nullptr, o2, g2);
            for (int j = 0; j < n_filt; j++)
                if (o2 >= filt[2 * j] && g2 >= filt[2 * j + 1]) { accepted = false; break; }
            if (accepted) {
                int w = 0;
                for (int j = 0; j < n_filt; j++)
                    if (o2 > filt[2 * j] || g2 > filt[2 * j + 1]) { filt[2 * w] = filt[2 * j]; filt[2 * w + 1] = filt[2 * j + 1]; w++; }
                filt[2 * w] = o2; filt[2 * w + 1] = g2;
                n_filt = w + 1;
                break;
            } else alpha *= P.line_search_tau;
        }
        if (accepted && it < 32) res.accept_mask |= (1u << it);
        // ---- take the step (osqp_interface.cpp:549-551) ----
        double inf = 0;
        for (int k = 0; k <= N; k++) {
            for (int m = 0; m < NX; m++) { guess[k * HZ + m] += alpha * (P.Tx[m] * step[k * HZ + m]); inf = fmax(inf, fabs(step[k * HZ + m])); }
            if (k < N) for (int j = 0; j < NU; j++) { guess[k * HZ + NX + j] += alpha * (P.Tu[j] * step[k * HZ + NX + j]); inf = fmax(inf, fabs(step[k * HZ + NX + j])); }
            else for (int j = 0; j < NU; j++) guess[k * HZ + NX + j] = 0.0;
        }
        if (log && log->n < log->max_log) {
            if (log->steps) for (int e = 0; e < (N + 1) * HZ; e++) log->steps[(size_t)log->n * (N + 1) * HZ + e] = step[e];
            log->alphas[log->n] = alpha;
            log->qp_ok[log->n] = qs.ok;
            log->n++;
        }
        if (alpha * inf < P.eps_prim) { res.status = SOLVED; res.iters = it + 1; done = true; break; }
    }
    if (!done) { res.status = MAX_ITER_EXCEEDED; res.iters = max_iter; }
    else if (res.status != SOLVED) res.iters = it;
    return res;
}

// epilogue of runMPC_ (mpc.cpp:140-188): status policy, returned horizon, warm start for the next cycle.
// Returns the reference's bool.
MPCC_HDN bool cycle_epilogue(int N, const SqpResult& r, const double* x0, const WsRef& guess, WarmFlags& fl) {
    if (r.status == SOLVED) {
        fl.valid = 1; fl.failed = 0;
    } else {
        // opt_sol = zero_guess (osqp_interface.cpp:422-428,585-589)
        for (int i = 0; i <= N; i++) {
            for (int e = 0; e < NX; e++) guess[i * HZ + e] = x0[e];
            for (int e = 0; e < NU; e++) guess[i * HZ + NX + e] = 0;
        }
        fl.valid = 0; fl.failed++;
    }
    return r.status == SOLVED || (r.status == MAX_ITER_EXCEEDED && fl.failed < 5);
}

}  // namespace mpcc
