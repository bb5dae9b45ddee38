// Reference-shaped C++ host API over the C ABI of libmpcc_b200.so (include/mpcc_cuda.h).
//
// Mirrors the surface of JunHeonYoon/MPCC_manipulator that sits on the hot path, so that a main.cpp-style caller
// (reference cpp/src/main.cpp:57-114) compiles against it:
//   mpcc::State / Input / PathToJson / ParamValue      reference cpp/include/types.h:33-150  (PODs bit-identical)
//   mpcc::OptVariables, ComputeTime, Status, MPCReturn  osqp_interface.h:48-79, solver_interface.h:28-42, mpc.h:45-56
//   mpcc::MPC                                           mpc.h:58-101   (runMPC, runMPC_, setTrack, getTrackLength, setParam)
//   mpcc::BatchMPC                                      new: B independent MPC objects on one GPU (the batched path)
//   mpcc::ArcLengthSpline, Integrator, Model, Cost, Constraints, RobotModel: per-function evaluators backed by the same
//   device code the cycle uses (mpcc_cuda_eval_*), for parity tests in the style of the reference's gtest suite.
// Eigen is not available in this image: vectors are std::array / std::vector, rotation matrices are row-major 9-arrays.
// No numerics live here; every call goes through the C ABI and fails with std::runtime_error on a non-zero return.
#pragma once
#include "../mpcc_cuda.h"

#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mpcc {

constexpr int NX = MPCC_NX, NU = MPCC_NU, NPC = 11, PANDA_DOF = 7;

struct State {  // types.h:33-71
    double q1, q2, q3, q4, q5, q6, q7, s, vs;
    void setZero() { q1 = q2 = q3 = q4 = q5 = q6 = q7 = s = vs = 0.0; }
    void unwrap(double track_length) { s = s < 0. ? 0. : (s > track_length ? track_length : s); }
    bool operator==(const State& o) const { return q1 == o.q1 && q2 == o.q2 && q3 == o.q3 && q4 == o.q4 && q5 == o.q5 && q6 == o.q6 && q7 == o.q7 && s == o.s && vs == o.vs; }
};
struct Input {  // types.h:73-118
    double dq1, dq2, dq3, dq4, dq5, dq6, dq7, dVs;
    void setZero() { dq1 = dq2 = dq3 = dq4 = dq5 = dq6 = dq7 = dVs = 0.0; }
    bool operator==(const Input& o) const { return dq1 == o.dq1 && dq2 == o.dq2 && dq3 == o.dq3 && dq4 == o.dq4 && dq5 == o.dq5 && dq6 == o.dq6 && dq7 == o.dq7 && dVs == o.dVs; }
};
static_assert(sizeof(State) == NX * sizeof(double) && sizeof(Input) == NU * sizeof(double), "PODs must map onto the C ABI arrays");

struct PathToJson { std::string param_path, cost_path, bounds_path, track_path, normalization_path, sqp_path; };  // types.h:127-134
struct ParamValue { std::map<std::string, double> param, cost, bounds, track, normalization, sqp; };            // types.h:143-150
struct OptVariables { State xk; Input uk; void setZero() { xk.setZero(); uk.setZero(); } };                      // osqp_interface.h:48-62
struct ComputeTime { double set_env, set_qp, solve_qp, get_alpha, total; void setZero() { set_env = set_qp = solve_qp = get_alpha = total = 0; } };
enum Status {  // solver_interface.h:28-42
    SOLVED = 0, MAX_ITER_EXCEEDED, QP_DualInfeasibleInaccurate, QP_PrimalInfeasibleInaccurate, QP_SolvedInaccurate, QP_MaxIterReached,
    QP_PrimalInfeasible, QP_DualInfeasible, Sigint, INVALID_SETTINGS, NAN_HESSIAN, NON_PD_HESSIAN
};
struct MPCReturn {  // mpc.h:45-56 (+ the solver status the reference only prints)
    Input u0;
    std::vector<OptVariables> mpc_horizon;
    ComputeTime compute_time;
    Status status = SOLVED;
    int sqp_iters = 0;
};

inline void check(int rc) { if (rc != MPCC_OK) throw std::runtime_error(std::string("mpcc_cuda: ") + mpcc_cuda_last_error()); }

// Parameter record (the six Params/*.json files) through the library's loader; ParamValue maps become "file.key" overrides.
inline std::vector<double> loadParams(const PathToJson& p, const ParamValue& v = ParamValue()) {
    std::vector<std::string> keys; std::vector<double> vals;
    auto add = [&](const char* f, const std::map<std::string, double>& m) { for (auto& kv : m) { keys.push_back(std::string(f) + "." + kv.first); vals.push_back(kv.second); } };
    add("model", v.param); add("cost", v.cost); add("bounds", v.bounds); add("normalization", v.normalization); add("sqp", v.sqp);
    std::vector<const char*> kp; for (auto& k : keys) kp.push_back(k.c_str());
    std::vector<double> out(MPCC_PARAMS_DOUBLES);
    check(mpcc_load_params_json(p.param_path.c_str(), p.cost_path.c_str(), p.bounds_path.c_str(), p.normalization_path.c_str(), p.sqp_path.c_str(),
                                kp.empty() ? nullptr : kp.data(), vals.empty() ? nullptr : vals.data(), (int32_t)keys.size(), out.data()));
    return out;
}

// Fitted 6-D arc-length path (ArcLengthSpline after gen6DSpline, arc_length_spline.cpp:213-265) as the table the kernels read.
class ArcLengthSpline {
public:
    ArcLengthSpline() : table_(MPCC_TRACK_DOUBLES, 0.0) {}
    void gen6DSpline(const std::vector<double>& X, const std::vector<double>& Y, const std::vector<double>& Z, const std::vector<std::array<double, 9>>& R) {
        check(mpcc_fit_track((int32_t)X.size(), X.data(), Y.data(), Z.data(), R.empty() ? nullptr : R[0].data(), table_.data()));
    }
    void loadJson(const std::string& track_path, const std::array<double, 3>* init_position = nullptr) {
        check(mpcc_load_track_json(track_path.c_str(), init_position ? init_position->data() : nullptr, table_.data()));
    }
    double getLength() const { return table_[MPCC_TRACK_DOUBLES - 3]; }  // TrackTable::length
    const std::vector<double>& table() const { return table_; }
private:
    std::vector<double> table_;
};

// Closed forms of the linear model (model.cpp:31-91, integrator.cpp:29-68): exact for this system.
struct Model {
    static void getLinModel(double Ts, double A[81], double B[72], double g[9]) {
        for (int i = 0; i < 81; i++) A[i] = 0;
        for (int i = 0; i < 72; i++) B[i] = 0;
        for (int i = 0; i < 9; i++) { g[i] = 0; A[10 * i] = 1; }
        A[7 * 9 + 8] = Ts;
        for (int j = 0; j < 7; j++) B[j * 8 + j] = Ts;
        B[7 * 8 + 7] = 0.5 * Ts * Ts; B[8 * 8 + 7] = Ts;
    }
};
struct Integrator {
    static State RK4(const State& x, const Input& u, double ts) {
        State n = x; const double* uu = &u.dq1; double* q = &n.q1;
        for (int j = 0; j < 7; j++) q[j] += ts * uu[j];
        n.s = x.s + ts * x.vs + 0.5 * ts * ts * u.dVs; n.vs = x.vs + ts * u.dVs;
        return n;
    }
};

// B independent mpcc::MPC objects on one GPU.
class BatchMPC {
public:
    BatchMPC(int batch, int horizon, double Ts, const PathToJson& path, const ParamValue& pv = ParamValue(), int device = 0) : B_(batch), N_(horizon), Ts_(Ts), path_(path) {
        mpcc_cuda_config cfg{}; cfg.batch = batch; cfg.horizon = horizon; cfg.Ts = Ts; cfg.device = device;
        check(mpcc_cuda_create(&cfg, &h_));
        ctor_pv_ = pv;
        auto p = loadParams(path_, pv); check(mpcc_cuda_set_params(h_, p.data(), 1));   // MPC(Ts, path, param_value), mpc.cpp:36-52
    }
    ~BatchMPC() { if (h_) mpcc_cuda_destroy(h_); }
    BatchMPC(const BatchMPC&) = delete; BatchMPC& operator=(const BatchMPC&) = delete;
    void loadNetworks(const std::string& self_path, const std::string& env_path) { check(mpcc_cuda_load_nn(h_, self_path.c_str(), env_path.c_str())); }
    // MPC::setParam on a live object (mpc.cpp:204-209, osqp_interface.cpp:95-100): the "param" and "cost" maps take
    // effect (file values + this map), bounds are rebuilt from the file, the normalisation and SQP parameters keep their
    // construction-time values, and the interface's own rddq stays the file's.  Warm starts are kept (main.cpp:103-106).
    void setParam(const ParamValue& pv) {
        ParamValue live; live.param = pv.param; live.cost = pv.cost; live.normalization = ctor_pv_.normalization; live.sqp = ctor_pv_.sqp;
        auto p = loadParams(path_, live); check(mpcc_cuda_set_params(h_, p.data(), 1));
    }
    void setParams(const std::vector<ParamValue>& per_instance) {
        std::vector<double> all; for (auto& pv : per_instance) { auto p = loadParams(path_, pv); all.insert(all.end(), p.begin(), p.end()); }
        check(mpcc_cuda_set_params(h_, all.data(), (int32_t)per_instance.size()));
    }
    void setTrack(const ArcLengthSpline& t) { track_len_.assign(1, t.getLength()); check(mpcc_cuda_set_tracks(h_, t.table().data(), 1, nullptr)); }
    void setTracks(const std::vector<ArcLengthSpline>& ts, const std::vector<int32_t>& track_of_instance) {
        std::vector<double> all; track_len_.clear();
        for (auto& t : ts) { all.insert(all.end(), t.table().begin(), t.table().end()); track_len_.push_back(t.getLength()); }
        check(mpcc_cuda_set_tracks(h_, all.data(), (int32_t)ts.size(), track_of_instance.data()));
    }
    // one control cycle for every instance; x0 is updated in place (s, vs) like MPC::runMPC_ (mpc.cpp:108,115)
    std::vector<bool> runMPC(std::vector<MPCReturn>& ret, std::vector<State>& x0, const std::vector<Input>& u0, const std::vector<std::array<double, 4>>* obs = nullptr) {
        const size_t S = N_ + 1;
        std::vector<double> uo(B_ * NU), hor(B_ * S * MPCC_HZ), tm(B_ * 4);
        std::vector<int32_t> st(B_), it(B_), ok(B_);
        check(mpcc_cuda_run_cycle(h_, &x0[0].q1, &u0[0].dq1, obs ? (*obs)[0].data() : nullptr, uo.data(), hor.data(), st.data(), it.data(), ok.data()));
        check(mpcc_cuda_read_compute_time(h_, tm.data()));
        ret.resize(B_);
        std::vector<bool> res(B_);
        for (int b = 0; b < B_; b++) {
            MPCReturn& r = ret[b];
            for (int j = 0; j < NU; j++) (&r.u0.dq1)[j] = uo[b * NU + j];
            r.mpc_horizon.resize(S);
            for (size_t k = 0; k < S; k++) {
                const double* e = &hor[(b * S + k) * MPCC_HZ];
                for (int m = 0; m < NX; m++) (&r.mpc_horizon[k].xk.q1)[m] = e[m];
                for (int j = 0; j < NU; j++) (&r.mpc_horizon[k].uk.dq1)[j] = e[NX + j];
            }
            r.compute_time = ComputeTime{0.0, tm[4 * b + 1], tm[4 * b + 2], tm[4 * b + 3], tm[4 * b]};
            r.status = (Status)st[b]; r.sqp_iters = it[b]; res[b] = ok[b] != 0;
        }
        return res;
    }
    std::vector<State> simTimeStep(const std::vector<State>& x, const std::vector<Input>& u, double ts) {  // integrator.cpp:55-68
        std::vector<State> xn(B_);
        check(mpcc_cuda_sim_time_step(h_, &x[0].q1, &u[0].dq1, ts, &xn[0].q1));
        return xn;
    }
    // RobotData::update + updateEnv (robot_data.h:55-88) for n joint vectors: 150 doubles each
    std::vector<double> robotData(const std::vector<std::array<double, 7>>& q, const std::vector<std::array<double, 4>>* obs = nullptr) {
        std::vector<double> rb(q.size() * MPCC_RB_DOUBLES);
        check(mpcc_cuda_eval_robot_data(h_, q[0].data(), obs ? (*obs)[0].data() : nullptr, (int32_t)q.size(), rb.data()));
        return rb;
    }
    // Stage linearisation by the device code (Cost::getCost, Constraints::getConstraints, Bounds, dynamics defect) for
    // instance 0's parameters / track 0: the 212 doubles of the normalised QP blocks (layout: csrc/dev_stage.cuh StageLin)
    std::array<double, MPCC_STAGE_LIN_DOUBLES> evalStage(const State& x, const Input& u, const std::array<double, 7>& dq_prev, const std::array<double, 7>& dq_next,
                                                         const State& x_next, const double* rb150, int k) {
        std::array<double, MPCC_STAGE_LIN_DOUBLES> out;
        const int32_t kk = k;
        check(mpcc_cuda_eval_stage(h_, &x.q1, &u.dq1, dq_prev.data(), dq_next.data(), &x_next.q1, rb150, &kk, 1, out.data()));
        return out;
    }
    void reset() { check(mpcc_cuda_reset(h_)); }
    double getTrackLength(int track = 0) const { return track_len_.at(track); }
    int batch() const { return B_; }
    int horizon() const { return N_; }
    mpcc_cuda_handle* handle() { return h_; }
private:
    int B_, N_; double Ts_; PathToJson path_; ParamValue ctor_pv_;
    mpcc_cuda_handle* h_ = nullptr;
    std::vector<double> track_len_;
};

// Panda kinematics through the device code (robot_model.cpp:366-450), one configuration at a time.
class RobotModel {
public:
    explicit RobotModel(BatchMPC* owner) : o_(owner) {}
    std::array<double, 3> getEEPosition(const std::array<double, 7>& q) { auto rb = o_->robotData({q}); return {rb[7], rb[8], rb[9]}; }
    std::array<double, 9> getEEOrientation(const std::array<double, 7>& q) { auto rb = o_->robotData({q}); std::array<double, 9> R; for (int i = 0; i < 9; i++) R[i] = rb[10 + i]; return R; }
    std::array<double, 42> getJacobian(const std::array<double, 7>& q) { auto rb = o_->robotData({q}); std::array<double, 42> J; for (int i = 0; i < 42; i++) J[i] = rb[19 + i]; return J; }  // [Jv; Jw]
    double getManipulability(const std::array<double, 7>& q) { return o_->robotData({q})[61]; }
    std::array<double, 7> getDManipulability(const std::array<double, 7>& q) { auto rb = o_->robotData({q}); std::array<double, 7> d; for (int i = 0; i < 7; i++) d[i] = rb[62 + i]; return d; }
private:
    BatchMPC* o_;
};

// Views of one stage linearisation in the reference's vocabulary (normalised by T_x / T_u, i.e. what the QP sees).
// Cost::getCost (cost.cpp:290-357): value, gradient, Gauss-Newton Hessian;  Constraints::getConstraints
// (constraints.cpp:192-243): the 11 polytopic rows c + C dx + D du <= 0 as gradients, RBF slopes and right-hand sides.
struct CostGrad { std::array<double, 9> f_x; std::array<double, 8> f_u; };
struct CostHess { std::array<double, 81> f_xx; std::array<double, 8> f_uu_diag; };
struct ConstraintsInfo { std::array<double, 11> rhs; std::array<double, 77> grad_h; std::array<double, 11> drbf; double l1_violation; };
class Cost {
public:
    explicit Cost(BatchMPC* owner) : o_(owner) {}
    double getCost(const State& x, const Input& u, const std::array<double, 7>& dq_prev, const std::array<double, 7>& dq_next, const State& x_next,
                   const double* rb150, int k, CostGrad* grad, CostHess* hess) {
        auto v = o_->evalStage(x, u, dq_prev, dq_next, x_next, rb150, k);
        if (grad) { for (int i = 0; i < 9; i++) grad->f_x[i] = v[45 + i]; for (int j = 0; j < 8; j++) grad->f_u[j] = v[62 + j]; }
        if (hess) {
            int q = 0;
            for (int r = 0; r < 9; r++) for (int c = 0; c <= r; c++) { hess->f_xx[9 * r + c] = hess->f_xx[9 * c + r] = v[q++]; }
            for (int j = 0; j < 8; j++) hess->f_uu_diag[j] = v[54 + j];
        }
        return v[210];
    }
private:
    BatchMPC* o_;
};
class Constraints {
public:
    explicit Constraints(BatchMPC* owner) : o_(owner) {}
    ConstraintsInfo getConstraints(const State& x, const Input& u, const std::array<double, 7>& dq_prev, const std::array<double, 7>& dq_next, const State& x_next,
                                   const double* rb150, int k) {
        auto v = o_->evalStage(x, u, dq_prev, dq_next, x_next, rb150, k);
        ConstraintsInfo c;
        for (int i = 0; i < 77; i++) c.grad_h[i] = v[111 + i];
        for (int j = 0; j < 11; j++) { c.drbf[j] = v[188 + j]; c.rhs[j] = v[199 + j]; }
        c.l1_violation = v[211];
        return c;
    }
private:
    BatchMPC* o_;
};

// Drop-in for the reference's single-instance controller (mpc.h:58-101).  N is a run-time argument here
// (reference: compile-time 10, config.h:36).
class MPC {
public:
    MPC(double Ts, const PathToJson& path, const ParamValue& pv = ParamValue(), int horizon = 10, int device = 0)
        : impl_(new BatchMPC(1, horizon, Ts, path, pv, device)), robot_(new RobotModel(impl_.get())) {}
    void loadNetworks(const std::string& self_path, const std::string& env_path) { impl_->loadNetworks(self_path, env_path); }
    bool runMPC(MPCReturn& mpc_return, State& x0, Input& u0) { return run(mpc_return, x0, u0, nullptr); }
    bool runMPC_(MPCReturn& mpc_return, State& x0, Input& u0, const std::array<double, 3>& obs_position, const double& obs_radius) {
        std::vector<std::array<double, 4>> o{{obs_position[0], obs_position[1], obs_position[2], obs_radius}};
        return run(mpc_return, x0, u0, &o);
    }
    void setTrack(const std::vector<double>& X, const std::vector<double>& Y, const std::vector<double>& Z, const std::vector<std::array<double, 9>>& R) {
        track_.gen6DSpline(X, Y, Z, R); impl_->setTrack(track_);
    }
    void setTrack(const ArcLengthSpline& t) { track_ = t; impl_->setTrack(track_); }
    ArcLengthSpline getTrack() { return track_; }
    double getTrackLength() { return track_.getLength(); }
    void setParam(const ParamValue& pv) { impl_->setParam(pv); }
    BatchMPC& batch() { return *impl_; }
private:
    bool run(MPCReturn& r, State& x0, Input& u0, const std::vector<std::array<double, 4>>* obs) {
        std::vector<MPCReturn> rr; std::vector<State> xs{x0}; std::vector<Input> us{u0};
        bool ok = impl_->runMPC(rr, xs, us, obs)[0];
        x0 = xs[0]; r = rr[0];
        return ok;
    }
    std::unique_ptr<BatchMPC> impl_;
    ArcLengthSpline track_;
public:
    std::unique_ptr<RobotModel> robot_;  // public member like the reference's (mpc.h:101)
};

}  // namespace mpcc
