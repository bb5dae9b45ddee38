// TEST INFRASTRUCTURE ONLY -- C ABI over the CPU oracle (oracle/mpcc_oracle.hpp)
// so pytest / bench.py can drive it through ctypes.  Never linked into the product.
#include "mpcc_oracle.hpp"

using namespace orc;

namespace {
struct NN { MLP self, env; };

void fill_params(MPC* m, const double* model7, const double* cost12, const double* bounds48, const double* norm17, const double* sqp9) {
    Param p = {model7[0], model7[1], model7[2], model7[3], model7[4], model7[5], model7[6]};
    CostParam c = {cost12[0], cost12[1], cost12[2], cost12[3], cost12[4], cost12[5], cost12[6], cost12[7], cost12[8], cost12[9], cost12[10], cost12[11]};
    BoundsParam b;
    int k = 0;
    for (int i = 0; i < NX; i++) b.lx[i] = bounds48[k++];
    for (int i = 0; i < NX; i++) b.ux[i] = bounds48[k++];
    for (int i = 0; i < NU; i++) b.lu[i] = bounds48[k++];
    for (int i = 0; i < NU; i++) b.uu[i] = bounds48[k++];
    for (int i = 0; i < DOF; i++) b.ldd[i] = bounds48[k++];
    for (int i = 0; i < DOF; i++) b.udd[i] = bounds48[k++];
    NormParam n;
    for (int i = 0; i < NX; i++) n.Tx[i] = norm17[i];
    for (int i = 0; i < NU; i++) n.Tu[i] = norm17[NX + i];
    SQPParam s = {sqp9[0], sqp9[1], (int)sqp9[2], (int)sqp9[3], sqp9[4] != 0, sqp9[5] != 0, sqp9[6], sqp9[7], sqp9[8]};
    m->param = p;
    m->solver.cost.cost_param = c;
    m->solver.cost.param = p;
    m->solver.constraints.param = p;
    m->solver.bounds.b = b;
    m->solver.bounds.param = p;
    m->solver.norm = n;
    m->solver.sqp = s;
    m->solver.r_ddq = c.r_ddq;
    m->track.max_dist_proj = p.max_dist_proj;
}
void horizon_out(const std::vector<OptVariables>& g, double* h) {
    for (size_t i = 0; i < g.size(); i++) {
        for (int m = 0; m < NX; m++) h[i * 17 + m] = g[i].xk.v[m];
        for (int k = 0; k < NU; k++) h[i * 17 + NX + k] = g[i].uk.v[k];
    }
}
std::vector<OptVariables> horizon_in(const double* h, int N) {
    std::vector<OptVariables> g(N + 1);
    for (int i = 0; i <= N; i++) {
        for (int m = 0; m < NX; m++) g[i].xk.v[m] = h[i * 17 + m];
        for (int k = 0; k < NU; k++) g[i].uk.v[k] = h[i * 17 + NX + k];
    }
    return g;
}
}  // namespace

extern "C" {

void* orc_nn_create(const double* self_w, const double* self_b, const double* env_w, const double* env_b) {
    NN* nn = new NN();
    nn->self.set(DOF, 1, {256, 64}, self_w, self_b);                      // osqp_interface.cpp:35-38
    nn->env.set(DOF + 3, NLINKS, {256, 256, 256, 256}, env_w, env_b);     // osqp_interface.cpp:40-43
    return nn;
}
void orc_nn_destroy(void* nn) { delete (NN*)nn; }

void* orc_mpc_create(int N, double Ts, void* nn, const double* model7, const double* cost12, const double* bounds48,
                     const double* norm17, const double* sqp9) {
    MPC* m = new MPC();
    m->init(N, Ts);
    fill_params(m, model7, cost12, bounds48, norm17, sqp9);
    m->solver.selcolNN = &((NN*)nn)->self;
    m->solver.envcolNN = &((NN*)nn)->env;
    return m;
}
void orc_mpc_destroy(void* m) { delete (MPC*)m; }
void orc_mpc_set_params(void* mp, const double* model7, const double* cost12, const double* bounds48, const double* norm17, const double* sqp9) {
    fill_params((MPC*)mp, model7, cost12, bounds48, norm17, sqp9);
}
// MPC::setParam on a LIVE object (mpc.cpp:204-209 + OsqpInterface::setParam, osqp_interface.cpp:95-100):
//   param_ = Param(file, pv.param); cost_ = Cost(path, pv); constraints_ = Constraints(Ts, path, pv);
//   bounds_ = Bounds(BoundsParam(FILE), Param(file, pv.param));
// normalization_param_, sqp_param_ and the interface's own cost_param_ (r_ddq) are NOT touched.  The caller passes
// model7 / cost12 = file values with the pv.param / pv.cost overrides applied and bounds48 = the file's bounds.
void orc_mpc_set_param_live(void* mp, const double* model7, const double* cost12, const double* bounds48_file) {
    MPC* m = (MPC*)mp;
    const NormParam n = m->solver.norm;
    const SQPParam s = m->solver.sqp;
    const double r_ddq = m->solver.r_ddq;
    double norm17[NX + NU], sqp9[9] = {s.eps_prim, s.eps_dual, (double)s.max_iter, (double)s.line_search_max_iter, s.do_SOC ? 1. : 0., s.use_BFGS ? 1. : 0., s.tau, s.eta, s.rho};
    for (int i = 0; i < NX; i++) norm17[i] = n.Tx[i];
    for (int i = 0; i < NU; i++) norm17[NX + i] = n.Tu[i];
    fill_params(m, model7, cost12, bounds48_file, norm17, sqp9);
    m->solver.r_ddq = r_ddq;
}
void orc_mpc_set_qp_options(void* mp, int max_iter, double eps) { ((MPC*)mp)->solver.qp.max_iter = max_iter; ((MPC*)mp)->solver.qp.eps = eps; }

void orc_mpc_set_track(void* mp, int n, const double* X, const double* Y, const double* Z, const double* R) {
    MPC* m = (MPC*)mp;
    Vec x(X, X + n), y(Y, Y + n), z(Z, Z + n);
    std::vector<Mat3> r(n);
    for (int i = 0; i < n; i++) for (int k = 0; k < 9; k++) r[i].m[k] = R[i * 9 + k];
    m->setTrack(x, y, z, r);
}
void orc_mpc_reset(void* mp) { MPC* m = (MPC*)mp; m->valid_initial_guess = false; m->num_valid_guess_failed = 0; }
double orc_mpc_track_length(void* mp) { return ((MPC*)mp)->track.getLength(); }

// one control cycle; x0 is updated in place (s, vs) like the reference (mpc.cpp:108,115)
int orc_mpc_run(void* mp, double* x0, const double* u0, const double* obs4, double* u_out, double* horizon, int* status, int* iters, double* times5) {
    MPC* m = (MPC*)mp;
    State x;
    Input u;
    for (int i = 0; i < NX; i++) x.v[i] = x0[i];
    for (int i = 0; i < NU; i++) u.v[i] = u0[i];
    MPCReturn r;
    bool ok = m->runMPC_(r, x, u, obs4, obs4[3]);
    for (int i = 0; i < NX; i++) x0[i] = x.v[i];
    for (int i = 0; i < NU; i++) u_out[i] = r.u0.v[i];
    if (horizon) horizon_out(r.mpc_horizon, horizon);
    if (status) *status = (int)r.status;
    if (iters) *iters = (r.status == SOLVED) ? r.sqp_iters + 1 : r.sqp_iters;
    if (times5) { times5[0] = r.compute_time.set_env; times5[1] = r.compute_time.set_qp; times5[2] = r.compute_time.solve_qp; times5[3] = r.compute_time.get_alpha; times5[4] = r.compute_time.total; }
    return ok ? 1 : 0;
}
int orc_mpc_warm_state(void* mp, double* horizon, int* valid, int* failed) {
    MPC* m = (MPC*)mp;
    horizon_out(m->initial_guess, horizon);
    *valid = m->valid_initial_guess; *failed = (int)m->num_valid_guess_failed;
    return 0;
}
// test harness: overwrite the warm start (mpc.h: initial_guess_, valid_initial_guess_, num_valid_guess_failed_) so that a cycle can
// be replayed from exactly the state another implementation started it from
int orc_mpc_set_warm_state(void* mp, const double* horizon, int valid, int failed) {
    MPC* m = (MPC*)mp;
    m->initial_guess = horizon_in(horizon, (int)m->initial_guess.size() - 1);
    m->valid_initial_guess = valid != 0; m->num_valid_guess_failed = (unsigned)failed;
    return 0;
}
// test harness: termination threshold of the dense QP solver (default 1e-9 on every KKT residual)
int orc_mpc_set_qp_eps(void* mp, double eps) { ((MPC*)mp)->solver.qp.eps = eps; return 0; }
// plant step used by the closed-loop harness (integrator.cpp:55-68)
void orc_sim_time_step(const double* x, const double* u, double ts, double* x_next) {
    State s; Input in;
    for (int i = 0; i < NX; i++) s.v[i] = x[i];
    for (int i = 0; i < NU; i++) in.v[i] = u[i];
    State r = Integrator::simTimeStep(s, in, ts);
    for (int i = 0; i < NX; i++) x_next[i] = r.v[i];
}
void orc_rk4(const double* x, const double* u, double ts, double* x_next) {
    State s; Input in;
    for (int i = 0; i < NX; i++) s.v[i] = x[i];
    for (int i = 0; i < NU; i++) in.v[i] = u[i];
    State r = Integrator::RK4(s, in, ts);
    for (int i = 0; i < NX; i++) x_next[i] = r.v[i];
}
void orc_lin_model(double Ts, double* A, double* B, double* g) { Model::getLinModel(Ts, A, B, g); }

// ---- probes -----------------------------------------------------------------
void orc_fk(const double* q, double* p3, double* R9, double* J42) {
    Vec3 p = RobotModel::getEEPosition(q);
    Mat3 R = RobotModel::getEEOrientation(q);
    for (int i = 0; i < 3; i++) p3[i] = p[i];
    for (int i = 0; i < 9; i++) R9[i] = R.m[i];
    RobotModel::getJacobian(q, J42);
}
void orc_manip(const double* q, double* m, double* dm7) { *m = RobotModel::getManipulability(q); RobotModel::getDManipulability(q, dm7); }
void orc_mlp_eval(void* nn, int which, const double* in, double* out, double* jac) {
    const MLP& m = which == 0 ? ((NN*)nn)->self : ((NN*)nn)->env;
    m.eval(in, out, jac);
}
void orc_robot_data(void* nn, const double* q, const double* obs4, double* rb150) {
    RobotData rb;
    rb.update(q, ((NN*)nn)->self);
    rb.updateEnv(obs4, obs4[3], ((NN*)nn)->env);
    rb.to_flat(rb150);
}
// out: pos3, dpos3, ddpos3, R9, dR3  (21 doubles)
void orc_track_eval(void* mp, double s, double* out) {
    const ArcLengthSpline& t = ((MPC*)mp)->track;
    Vec3 p = t.getPosition(s), d = t.getDerivative(s), dd = t.getSecondDerivative(s), dr = t.getOrientationDerivative(s);
    Mat3 R = t.getOrientation(s);
    for (int i = 0; i < 3; i++) { out[i] = p[i]; out[3 + i] = d[i]; out[6 + i] = dd[i]; out[18 + i] = dr[i]; }
    for (int i = 0; i < 9; i++) out[9 + i] = R.m[i];
}
void orc_track_table(void* mp, double* s, double* X, double* Y, double* Z, double* R) {
    const PathData& p = ((MPC*)mp)->track.path_data;
    for (int i = 0; i < p.n_points; i++) {
        s[i] = p.s[i]; X[i] = p.X[i]; Y[i] = p.Y[i]; Z[i] = p.Z[i];
        for (int k = 0; k < 9; k++) R[i * 9 + k] = p.R[i].m[k];
    }
}
double orc_project(void* mp, double s, const double* ee3) { Vec3 e = {ee3[0], ee3[1], ee3[2]}; return ((MPC*)mp)->track.projectOnSpline(s, e); }

void orc_stage_cost(void* mp, const double* x9, const double* u8, const double* rb150, int k, double* obj, double* fx9, double* fu8, double* fxx81, double* fuu64) {
    MPC* m = (MPC*)mp;
    State x; Input u; RobotData rb;
    for (int i = 0; i < NX; i++) x.v[i] = x9[i];
    for (int i = 0; i < NU; i++) u.v[i] = u8[i];
    rb.from_flat(rb150);
    CostGrad g; CostHess h;
    m->solver.cost.getCost(m->solver.track, x, u, rb, k, obj, &g, &h);
    for (int i = 0; i < NX; i++) fx9[i] = g.f_x[i];
    for (int i = 0; i < NU; i++) fu8[i] = g.f_u[i];
    for (int i = 0; i < NX * NX; i++) fxx81[i] = h.f_xx[i];
    for (int i = 0; i < NU * NU; i++) fuu64[i] = h.f_uu[i];
}
void orc_stage_constraints(void* mp, const double* x9, const double* u8, const double* rb150, int k, double* c11, double* cl11, double* cu11, double* cx99, double* cu88) {
    MPC* m = (MPC*)mp;
    State x; Input u; RobotData rb;
    for (int i = 0; i < NX; i++) x.v[i] = x9[i];
    for (int i = 0; i < NU; i++) u.v[i] = u8[i];
    rb.from_flat(rb150);
    ConstraintsInfo ci; ConstraintsJac cj;
    m->solver.constraints.getConstraints(x, u, rb, k, &ci, &cj);
    for (int i = 0; i < NPC; i++) { c11[i] = ci.c_vec[i]; cl11[i] = ci.c_lvec[i]; cu11[i] = ci.c_uvec[i]; }
    for (int i = 0; i < NPC * NX; i++) cx99[i] = cj.c_x[i];
    for (int i = 0; i < NPC * NU; i++) cu88[i] = cj.c_u[i];
}
void orc_qp_dims(int N, int* n_var, int* n_constr) {
    *n_var = (N + 1) * NX + N * NU;
    *n_constr = (N + 1) * NX + (*n_var + N * NU) + (N + 1) * NPC;
}
// dense flat QP exactly as osqp_interface.cpp assembles it
void orc_build_qp(void* mp, const double* guess, const double* rb, const double* cur_u, double* P, double* q, double* A, double* l, double* u, double* c, double* obj) {
    MPC* m = (MPC*)mp;
    Solver& s = m->solver;
    std::vector<OptVariables> g = horizon_in(guess, s.N);
    for (int i = 0; i <= s.N; i++) s.rb[i].from_flat(rb + (size_t)i * RobotData::FLAT);
    for (int i = 0; i < NU; i++) s.current_u.v[i] = cur_u[i];
    Mat H, J;
    Vec grad, cv, lv, uv;
    s.setCost(g, obj, &grad, &H);
    s.setConstraints(g, &J, &cv, &lv, &uv);
    std::copy(H.a.begin(), H.a.end(), P);
    std::copy(grad.begin(), grad.end(), q);
    std::copy(J.a.begin(), J.a.end(), A);
    std::copy(lv.begin(), lv.end(), l);
    std::copy(uv.begin(), uv.end(), u);
    std::copy(cv.begin(), cv.end(), c);
}
int orc_solve_qp_dense(int n, int m, const double* P, const double* q, const double* A, const double* l, const double* u, double* z, int* iters, int max_iter, double eps) {
    Mat Pm(n, n), Am(m, n);
    std::copy(P, P + (size_t)n * n, Pm.a.begin());
    std::copy(A, A + (size_t)m * n, Am.a.begin());
    Vec qv(q, q + n), lv(l, l + m), uv(u, u + m), zv(n, 0.0);
    DenseQP qp;
    qp.max_iter = max_iter; qp.eps = eps;
    QPResult r = qp.solve(Pm, qv, Am, lv, uv, zv);
    if (iters) *iters = r.iters;
    if (r.ok) std::copy(zv.begin(), zv.end(), z);
    return r.ok ? 1 : 0;
}
// SQP loop on a given warm start with frozen RobotData; logs every iteration's QP step.
int orc_solve_ocp(void* mp, double* guess, const double* rb, const double* cur_u, int* status, int* iters,
                  double* steps, double* alphas, int* qp_ok, int max_log, int* n_logged) {
    MPC* m = (MPC*)mp;
    Solver& s = m->solver;
    std::vector<OptVariables> g = horizon_in(guess, s.N);
    s.initial_guess = g;
    s.initial_guess_vec = s.OptvarToVector(g);
    for (int i = 0; i <= s.N; i++) s.rb[i].from_flat(rb + (size_t)i * RobotData::FLAT);
    for (int i = 0; i < NU; i++) s.current_u.v[i] = cur_u[i];
    SqpLog lg;
    s.log = &lg;
    Status st = SOLVED;
    ComputeTime tm;
    std::vector<OptVariables> sol;
    bool ok = s.solveOCP(sol, &st, &tm);
    s.log = nullptr;
    horizon_out(sol, guess);
    *status = (int)st;
    *iters = (st == SOLVED) ? s.sqp_iter + 1 : s.sqp_iter;
    int nl = std::min<int>(max_log, (int)lg.steps.size());
    for (int i = 0; i < nl; i++) {
        std::copy(lg.steps[i].begin(), lg.steps[i].end(), steps + (size_t)i * s.N_var);
        alphas[i] = lg.alphas[i];
        qp_ok[i] = lg.qp_ok[i];
    }
    *n_logged = nl;
    return ok ? 1 : 0;
}
// smallest robustness margin over the filter comparisons of the last solveOCP (see filter_margin)
double orc_mpc_last_filter_margin(void* mp) { return ((MPC*)mp)->solver.min_filter_margin; }
// follow these first-trial accept/reject decisions in the next solveOCP calls (n = 0: natural behaviour)
void orc_mpc_set_forced_decisions(void* mp, const int* accept, int n) { ((MPC*)mp)->solver.forced_accept.assign(accept, accept + n); }
// the line search's own decisions and their robustness margins in the last solveOCP
int orc_mpc_decision_log(void* mp, int* natural, double* margins, int max_n) {
    Solver& s = ((MPC*)mp)->solver;
    int n = std::min<int>(max_n, (int)s.natural_accept.size());
    for (int i = 0; i < n; i++) { natural[i] = s.natural_accept[i]; margins[i] = s.accept_margin[i]; }
    return n;
}
double orc_rbf(double delta, double h) { return getRBF(delta, h); }
void orc_log_exp(const double* R9, double* log9, double* exp_of_log9) {
    Mat3 R; for (int i = 0; i < 9; i++) R.m[i] = R9[i];
    Mat3 L = LogMatrix(R), E = ExpMatrix(L);
    for (int i = 0; i < 9; i++) { log9[i] = L.m[i]; exp_of_log9[i] = E.m[i]; }
}
// natural cubic spline fit/eval probe (cubic_spline.cpp) for the reference's spline property test
void orc_cubic_spline(int n, const double* x, const double* y, int regular, int nq, const double* xq, double* yq, double* dyq, double* ddyq) {
    CubicSpline sp;
    sp.genSpline(Vec(x, x + n), Vec(y, y + n), regular != 0);
    for (int i = 0; i < nq; i++) { yq[i] = sp.getPoint(xq[i]); dyq[i] = sp.getDerivative(xq[i]); ddyq[i] = sp.getSecondDerivative(xq[i]); }
}

}  // extern "C"
