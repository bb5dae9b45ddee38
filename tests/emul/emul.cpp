// TEST INFRASTRUCTURE ONLY.  Compiles the product's device functions (csrc/dev_*.cuh) for the HOST
// so the CPU-only test tier can exercise the exact code the CUDA kernels inline (there is no GPU in
// the build container).  This object is never part of libmpcc_b200.so and is not a fallback path.
#include "../../mpcc_manipulator_b200/csrc/dev_sqp.cuh"
#include "../../mpcc_manipulator_b200/csrc/sqp_warp.cuh"
#include "../../mpcc_manipulator_b200/csrc/host/track_fit.h"
#include "../../mpcc_manipulator_b200/csrc/dev_track_fit.cuh"
#include "../../mpcc_manipulator_b200/csrc/mlp_oz_kernel.cuh"
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
// the device track fit (dev_track_fit.cuh) compiled for the host; stride > 1 emulates the kernel's [element][track] scratch layout
void emu_fit_track_device_code(int n, const double* X, const double* Y, const double* Z, const double* R, int stride, double* table) {
    std::vector<double> ws(track_fit_scratch_doubles(n) * stride, -7.0);
    tf_fit_track(n, TArr{(double*)X, 1}, TArr{(double*)Y, 1}, TArr{(double*)Z, 1}, TArr{(double*)R, 1}, TArr{ws.data() + (stride > 1 ? 1 : 0), (size_t)stride}, *(TrackTable*)table);
}
void emu_track_from_knots(const double* s, const double* X, const double* Y, const double* Z, const double* R, double* table) {
    std::vector<double> w((size_t)14 * N_SPLINE);
    tf_table_from_knots(TArr{(double*)s, 1}, TArr{(double*)X, 1}, TArr{(double*)Y, 1}, TArr{(double*)Z, 1}, TArr{(double*)R, 1}, TArr{w.data(), 1}, N_SPLINE, *(TrackTable*)table);
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

// The lanes-per-instance formulation (sqp_warp.cuh) executed phase by phase on the host; reverse = lane order.
// NL = 32: the warp kernel (throughput); NL = 128: the CTA kernel (latency mode).
}  // extern "C"
template <int NL, bool SOC = false>
static int group_solve_ocp(const double* params, const double* table, double Ts, int N, double* guess, const double* rb, const double* cur_u,
                           int qp_max_iter, double qp_eps, int reverse, int* status, int* iters, int* qp_iters, double* steps, double* alphas, int* qp_ok,
                           int max_log, int* n_logged, unsigned* accept_mask) {
    const Params& P = *(const Params*)params;
    const int S = N + 1, HN = S * HZ;
    std::vector<double> gws(warp_ws_doubles(N)), sm(group_smem_doubles<NL>(N), -3.0);
    Lanes<NL> wp; wp.reverse = reverse != 0;
    GroupSqp<NL, SOC> w{P, *(const TrackTable*)table, make_dyn(P, Ts), Ts, N, S, QpOptions{qp_max_iter, qp_eps}, wp};
    w.carve(gws.data(), sm.data());
    for (int e = 0; e < HN; e++) w.GUESS[e] = guess[e];
    SqpLogRef lg{steps, alphas, qp_ok, max_log, 0};
    SqpResult r = w.run(cur_u, rb, 1, RB_DOUBLES, max_log > 0 ? &lg : nullptr);
    for (int e = 0; e < HN; e++) guess[e] = w.GUESS[e];
    *status = r.status; *iters = r.iters; *qp_iters = r.qp_iters; *n_logged = lg.n; *accept_mask = r.accept_mask;
    return r.status == SOLVED;
}
extern "C" {
int emu_warp_solve_ocp(const double* params, const double* table, double Ts, int N, double* guess, const double* rb, const double* cur_u,
                       int qp_max_iter, double qp_eps, int reverse, int* status, int* iters, int* qp_iters, double* steps, double* alphas, int* qp_ok,
                       int max_log, int* n_logged, unsigned* accept_mask) {
    return group_solve_ocp<32>(params, table, Ts, N, guess, rb, cur_u, qp_max_iter, qp_eps, reverse, status, iters, qp_iters, steps, alphas, qp_ok, max_log, n_logged, accept_mask);
}
int emu_cta_solve_ocp(const double* params, const double* table, double Ts, int N, double* guess, const double* rb, const double* cur_u,
                      int qp_max_iter, double qp_eps, int reverse, int* status, int* iters, int* qp_iters, double* steps, double* alphas, int* qp_ok,
                      int max_log, int* n_logged, unsigned* accept_mask) {
    return group_solve_ocp<128>(params, table, Ts, N, guess, rb, cur_u, qp_max_iter, qp_eps, reverse, status, iters, qp_iters, steps, alphas, qp_ok, max_log, n_logged, accept_mask);
}
// the same with the second-order correction compiled in (sqp.json "do_SOC" in params decides at run time)
int emu_warp_solve_ocp_soc(const double* params, const double* table, double Ts, int N, double* guess, const double* rb, const double* cur_u,
                           int qp_max_iter, double qp_eps, int reverse, int* status, int* iters, int* qp_iters, double* steps, double* alphas, int* qp_ok,
                           int max_log, int* n_logged, unsigned* accept_mask) {
    return group_solve_ocp<32, true>(params, table, Ts, N, guess, rb, cur_u, qp_max_iter, qp_eps, reverse, status, iters, qp_iters, steps, alphas, qp_ok, max_log, n_logged, accept_mask);
}
int emu_cta_solve_ocp_soc(const double* params, const double* table, double Ts, int N, double* guess, const double* rb, const double* cur_u,
                          int qp_max_iter, double qp_eps, int reverse, int* status, int* iters, int* qp_iters, double* steps, double* alphas, int* qp_ok,
                          int max_log, int* n_logged, unsigned* accept_mask) {
    return group_solve_ocp<128, true>(params, table, Ts, N, guess, rb, cur_u, qp_max_iter, qp_eps, reverse, status, iters, qp_iters, steps, alphas, qp_ok, max_log, n_logged, accept_mask);
}
// one QP of the warp formulation at the linearisation of `guess`
int emu_warp_solve_qp(const double* params, const double* table, double Ts, int N, const double* guess, const double* rb, const double* cur_u,
                      int qp_max_iter, double qp_eps, int reverse, double* step_out, int* iters, double* res3) {
    const Params& P = *(const Params*)params;
    const int S = N + 1, HN = S * HZ;
    std::vector<double> gws(warp_ws_doubles(N)), sm(warp_smem_doubles(N));
    Warp wp; wp.reverse = reverse != 0;
    WarpSqp w{P, *(const TrackTable*)table, make_dyn(P, Ts), Ts, N, S, QpOptions{qp_max_iter, qp_eps}, wp};
    w.carve(gws.data(), sm.data());
    for (int e = 0; e < HN; e++) { w.GUESS[e] = guess[e]; w.XG[e] = guess[e]; w.XS[e] = 0.0; }
    w.init_scratch();
    double obj, gap;
    bool f1, f2;
    w.eval_horizon<true>(cur_u, rb, 1, RB_DOUBLES, 0.0, true, obj, gap, &f1, &f2);
    for (int c = 0; c < NU * N; c++) {
        const int k = c / NX, m = c % NX, i = c / NU, kk = c % NU;
        const double uv = guess[i * HZ + NX + kk];
        double* L = w.LIN + (size_t)k * WL_SIZE;
        L[WL_XLO + m] = fmax(L[WL_XLO + m], (P.lu[kk] - uv) / P.Tu[kk]);
        L[WL_XHI + m] = fmin(L[WL_XHI + m], (P.uu[kk] - uv) / P.Tu[kk]);
    }
    QpStats qs = w.solve();
    for (int e = 0; e < HN; e++) step_out[e] = w.VAR[e];
    for (int j = 0; j < NU; j++) step_out[N * HZ + NX + j] = 0.0;
    *iters = qs.iters; res3[0] = qs.res_dual; res3[1] = qs.res_prim; res3[2] = qs.gap;
    return qs.ok;
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

// ---- int8-split layer of the MLP kernel (mlp_oz_kernel.cuh): the host packing (pack_mlp_oz_weights) and the scalar arithmetic the kernel uses
//      (oz_col_scales, oz_quantize, oz_digit), with the tensor-core products restated as plain int32 sums ----
extern "C" {
int emu_oz_slices() { return OZ_S; }
// Y[256][ncol] = W[256][256] X[256][ncol] through the split: digits of W from the packed stream, digits of X per column, exact int32 group sums,
// int64 Horner, one rounding.  Also returns the largest |digit| seen (must be <= 64) and the largest |group sum| (must stay below 2^23).
void emu_oz_layer(const double* W, const double* X, int ncol, double* Y, int* max_digit, long long* max_group) {
    std::vector<double> zero30(256 * 30, 0.0), zero21(256 * 21, 0.0), zero256(256 * 256, 0.0), zero9(9 * 256, 0.0), zero64(64 * 256, 0.0), zero1(64, 0.0);
    const double* eW[5] = {zero30.data(), W, zero256.data(), zero256.data(), zero9.data()};
    const double* sW[3] = {zero21.data(), zero64.data(), zero1.data()};
    std::vector<double> dpack(OZ_DPACK_D), rowscale(3 * 256);
    std::vector<uint8_t> qpack((size_t)OZ_CHUNKS_PER_TILE * OZ_CHUNK);
    pack_mlp_oz_weights(eW, sW, dpack.data(), qpack.data(), rowscale.data());
    int md = 0;
    long long mg = 0;
    std::vector<int> b((size_t)OZ_S * 256);
    for (int c = 0; c < ncol; c++) {
        double mx = 0.0;
        for (int k = 0; k < 256; k++) mx = std::fmax(mx, std::fabs(X[(size_t)k * ncol + c]));
        unsigned long long bits;
        std::memcpy(&bits, &mx, 8);
        double sc, cs;
        oz_col_scales((uint32_t)(bits >> 32) & 0x7FFFFFFFu, sc, cs);
        for (int k = 0; k < 256; k++) {
            const long long q = oz_quantize(X[(size_t)k * ncol + c], sc);
            for (int t = 0; t < OZ_S; t++) {
                const int d = oz_digit(q, t);
                b[(size_t)(OZ_S - 1 - t) * 256 + k] = d;
                md = std::max(md, std::abs(d));
            }
        }
        for (int r = 0; r < 256; r++) {
            long long S = 0;
            for (int g = 0; g < OZ_S; g++) {
                long long G = 0;
                for (int i = 0; i <= g; i++) {
                    const int j = g - i;
                    for (int k = 0; k < 256; k++) {
                        const int a = (int)(int8_t)qpack[oz_wq_index(0, r, k, i)];
                        md = std::max(md, std::abs(a));
                        G += (long long)a * b[(size_t)j * 256 + k];
                    }
                }
                mg = std::max(mg, (long long)std::llabs(G));
                S = S * 128 + G;
            }
            Y[(size_t)r * ncol + c] = ((double)S * rowscale[r]) * cs;
        }
    }
    *max_digit = md;
    *max_group = mg;
}
}
