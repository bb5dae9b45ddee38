// TEST INFRASTRUCTURE ONLY.  Compiles the product's device functions (csrc/dev_*.cuh) for the HOST
// so the CPU-only test tier can exercise the exact code the CUDA kernels inline (there is no GPU in
// the build container).  This object is never part of libmpcc_b200.so and is not a fallback path.
#include "../../mpcc_manipulator_b200/csrc/dev_sqp.cuh"
#include "../../mpcc_manipulator_b200/csrc/host/track_fit.h"
#include <cstring>
#include <vector>

using namespace mpcc;

extern "C" {

int emu_params_doubles() { return PARAMS_DOUBLES; }
int emu_track_doubles() { return TRACK_DOUBLES; }
int emu_qp_ws_doubles(int N) { return qp_workspace_doubles(N); }

void emu_fit_track(int n, const double* X, const double* Y, const double* Z, const double* R, double* table) {
    Waypoints w;
    w.X.assign(X, X + n); w.Y.assign(Y, Y + n); w.Z.assign(Z, Z + n); w.R.assign(R, R + 9 * n);
    fit_track(w, *(TrackTable*)table);
}
void emu_kin(const double* q, double* out62) {
    PandaKin k;
    panda_kinematics(q, k);
    std::memcpy(out62, k.p, 3 * 8); std::memcpy(out62 + 3, k.R, 9 * 8); std::memcpy(out62 + 12, k.Jv, 21 * 8); std::memcpy(out62 + 33, k.Jw, 21 * 8);
    out62[54] = panda_manipulability_from(k.Jv, k.Jw);
    panda_dmanipulability(q, out62 + 55);
}
void emu_track_eval(const double* table, double s, double* out21) {
    const TrackTable& t = *(const TrackTable*)table;
    TrackPoint tp;
    track_eval_pos(t, s, tp);
    for (int i = 0; i < 3; i++) { out21[i] = tp.pos[i]; out21[3 + i] = tp.dpos[i]; out21[6 + i] = tp.ddpos[i]; }
    track_eval_rot(t, s, out21 + 9, out21 + 18);
}
double emu_project(const double* table, double max_dist, double s, const double* ee) { return track_project(*(const TrackTable*)table, max_dist, s, ee); }

void emu_stage_lin(const double* params, const double* table, double Ts, int N, int k, const double* x, const double* u, const double* up,
                   const double* un, const double* xn, const double* rb150, double* out212) {
    StageLin sl;
    std::memset(&sl, 0, sizeof(sl));
    RbView rv{rb150, 1};
    stage_eval<true>(*(const Params*)params, *(const TrackTable*)table, Ts, N, k, x, u, up, un, xn, rv, sl);
    std::memcpy(out212, &sl, sizeof(sl));
}
void emu_so3(const double* R9, double* log3, double* exp9) { so3_log(R9, log3); so3_exp(log3, exp9); }

void emu_prologue(const double* params, const double* table, double Ts, int N, double* x0, const double* u0, double* warm, int* valid, int* failed) {
    WarmFlags fl{*valid, *failed};
    cycle_prologue(*(const Params*)params, *(const TrackTable*)table, Ts, N, x0, u0, WsRef{warm, 1}, fl);
    *valid = fl.valid; *failed = fl.failed;
}

int emu_solve_ocp(const double* params, const double* table, double Ts, int N, double* guess, const double* rb, const double* cur_u,
                  int qp_max_iter, double qp_eps, int* status, int* iters, int* qp_iters, double* steps, double* alphas, int* qp_ok, int max_log, int* n_logged) {
    const Params& P = *(const Params*)params;
    const int HN = (N + 1) * HZ;
    std::vector<double> step(HN), trial(HN), filt(2 * ((int)P.max_iter + 2)), ws(qp_workspace_doubles(N));
    SqpLogRef lg{steps, alphas, qp_ok, max_log, 0};
    SqpResult r = sqp_solve(P, *(const TrackTable*)table, Ts, N, WsRef{guess, 1}, WsRef{step.data(), 1}, WsRef{trial.data(), 1}, WsRef{filt.data(), 1},
                            cur_u, rb, 1, RB_DOUBLES, WsRef{ws.data(), 1}, QpOptions{qp_max_iter, qp_eps}, &lg);
    *status = r.status; *iters = r.iters; *qp_iters = r.qp_iters; *n_logged = lg.n;
    return r.status == SOLVED;
}

// epilogue of runMPC_ (status policy, fallback horizon, warm-start flags); returns the reference's bool
int emu_epilogue(int N, int status, int iters, const double* x0, double* guess, int* valid, int* failed) {
    WarmFlags fl{*valid, *failed};
    SqpResult r{status, iters, 0, 0, 0};
    bool ok = cycle_epilogue(N, r, x0, WsRef{guess, 1}, fl);
    *valid = fl.valid; *failed = fl.failed;
    return ok ? 1 : 0;
}

// solve only the QP of the current linearisation; returns the normalised step in horizon layout
int emu_solve_qp(const double* params, const double* table, double Ts, int N, const double* guess, const double* rb, const double* cur_u,
                 int qp_max_iter, double qp_eps, double* step_out, int* iters, double* res3) {
    const Params& P = *(const Params*)params;
    std::vector<double> ws(qp_workspace_doubles(N));
    StageQP qp{P, make_dyn(P, Ts), N, WsRef{ws.data(), 1}, QpOptions{qp_max_iter, qp_eps}};
    double obj, gap;
    WsRef g{(double*)guess, 1};
    eval_horizon<true>(P, *(const TrackTable*)table, Ts, N, g, cur_u, rb, 1, RB_DOUBLES, &qp, obj, gap);
    struct GU { const double* g; double operator()(int i, int kk) const { return g[i * HZ + NX + kk]; } } gu{guess};
    qp.apply_input_bound_quirk(gu);
    QpStats qs = qp.solve();
    for (int k = 0; k <= N; k++) {
        for (int m = 0; m < NX; m++) step_out[k * HZ + m] = qp.var(k)[V_XI + m];
        for (int j = 0; j < NU; j++) step_out[k * HZ + NX + j] = (k < N) ? qp.var(k)[V_NU + j] : 0.0;
    }
    *iters = qs.iters; res3[0] = qs.res_dual; res3[1] = qs.res_prim; res3[2] = qs.gap;
    return qs.ok;
}
}
