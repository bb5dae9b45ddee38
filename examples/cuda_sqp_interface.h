// cpp/include/Interfaces/cuda_sqp_interface.h -- the ONE file a maintainer adds to JunHeonYoon/MPCC_manipulator to put
// libmpcc_b200.so behind the reference's solver seam (class SolverInterface, cpp/include/Interfaces/solver_interface.h:44-54),
// single instance, reference `MPC` unchanged.  Then cpp/src/MPC/mpc.cpp:31,45 become
//     solver_interface_(new CudaSqpInterface(Ts, path[, param_value]))
// and cpp/CMakeLists.txt links mpcc_b200.  Only C types cross the boundary; Eigen stays on the reference side.
// Compile-checked in this repository against tests/stubs (tests/test_cpp_api.py::test_solver_interface_binding_compiles).
#pragma once
#include "Interfaces/solver_interface.h"
#include "mpcc_cuda.h"   // this repository: include/mpcc_cuda.h

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mpcc {
class CudaSqpInterface : public SolverInterface {
public:
    CudaSqpInterface(double Ts, const PathToJson& path, const ParamValue& pv = ParamValue()) : path_(path), ctor_pv_(pv) {
        mpcc_cuda_config cfg{};
        cfg.batch = 1; cfg.horizon = N; cfg.Ts = Ts; cfg.device = 0;
        check(mpcc_cuda_create(&cfg, &h_));
        // the reference's own weight_k.txt / bias_k.txt (osqp_interface.cpp:35-43 reads the same directories)
        check(mpcc_cuda_load_nn(h_, (pkg_path + "NNmodel/self/parameter").c_str(), (pkg_path + "NNmodel/env/parameter").c_str()));
        upload(pv);   // OsqpInterface(Ts, path, param_value): every map but "bounds" acts (osqp_interface.cpp:50-58)
    }
    ~CudaSqpInterface() override { mpcc_cuda_destroy(h_); }

    void setTrack(const ArcLengthSpline track) override {            // solver_interface.h:46
        // the spline is already fitted: hand over its 100 knots, no second fit / resample pass
        const PathData p = track.getPathData();
        std::vector<double> R(9 * (size_t)p.n_points);
        for (int i = 0; i < p.n_points; i++)
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[9 * i + 3 * r + c] = p.R[i](r, c);   // row-major
        std::vector<double> table(MPCC_TRACK_DOUBLES);
        check(mpcc_track_from_knots(p.s.data(), p.X.data(), p.Y.data(), p.Z.data(), R.data(), table.data()));
        check(mpcc_cuda_set_tracks(h_, table.data(), 1, nullptr));
    }
    void setParam(const ParamValue& pv) override {                   // :47, live semantics of osqp_interface.cpp:95-100
        ParamValue live;
        live.param = pv.param; live.cost = pv.cost;                  // cost_, constraints_, bounds_(file, Param(file, pv.param)) are rebuilt
        live.normalization = ctor_pv_.normalization; live.sqp = ctor_pv_.sqp;   // never updated on a live object
        upload(live);
    }
    void setEnvData(const Eigen::Vector3d& obs, const double& r) override { obs_[0] = obs(0); obs_[1] = obs(1); obs_[2] = obs(2); obs_[3] = r; }   // :48
    void setInitialGuess(const std::vector<OptVariables>& g) override { guess_ = g; }                                                          // :49
    void setCurrentInput(const Input& u) override { std::memcpy(cur_u_, &u, sizeof(u)); }                                                      // :50
    bool solveOCP(std::vector<OptVariables>& opt, Status* status, ComputeTime* t) override {                                                   // :51
        // frozen RobotData at the warm start (robot_data.h:55-88) and the SQP loop, both on the device; no allocation per call
        double q[(N + 1) * 7], obs[(N + 1) * 4], hor[(N + 1) * MPCC_HZ];
        static_assert(sizeof(OptVariables) == MPCC_HZ * sizeof(double), "OptVariables must be 17 doubles");
        for (int k = 0; k <= N; k++) {
            std::memcpy(q + 7 * k, &guess_[k].xk, 7 * sizeof(double));
            std::memcpy(obs + 4 * k, obs_, sizeof(obs_));
            std::memcpy(hor + MPCC_HZ * k, &guess_[k], MPCC_HZ * sizeof(double));
        }
        rb_.resize((size_t)(N + 1) * MPCC_RB_DOUBLES);
        check(mpcc_cuda_eval_robot_data(h_, q, obs, N + 1, rb_.data()));
        int32_t st = 0, it = 0, nl = 0;
        check(mpcc_cuda_solve_ocp(h_, hor, rb_.data(), cur_u_, 1, &st, &it, nullptr, nullptr, 0, &nl));
        double tm[4];
        check(mpcc_cuda_read_compute_time(h_, tm));
        if (t) { t->total = tm[0]; t->set_qp = tm[1]; t->solve_qp = tm[2]; t->get_alpha = tm[3]; t->set_env = 0.0; }
        *status = (Status)st;
        opt.resize(N + 1);
        for (int k = 0; k <= N; k++) std::memcpy(&opt[k], hor + MPCC_HZ * k, MPCC_HZ * sizeof(double));   // zero_guess on failure, osqp_interface.cpp:580-589
        return st == SOLVED;
    }

private:
    static void check(int rc) { if (rc) throw std::runtime_error(mpcc_cuda_last_error()); }
    void upload(const ParamValue& pv) {
        std::vector<std::string> keys; std::vector<double> vals;     // "cost.qC", "model.tol_sing", ...
        auto add = [&](const char* f, const std::map<std::string, double>& m) { for (auto& kv : m) { keys.push_back(std::string(f) + "." + kv.first); vals.push_back(kv.second); } };
        add("model", pv.param); add("cost", pv.cost); add("normalization", pv.normalization); add("sqp", pv.sqp);
        std::vector<const char*> kp; for (auto& k : keys) kp.push_back(k.c_str());
        double params[MPCC_PARAMS_DOUBLES];
        check(mpcc_load_params_json(path_.param_path.c_str(), path_.cost_path.c_str(), path_.bounds_path.c_str(), path_.normalization_path.c_str(),
                                    path_.sqp_path.c_str(), kp.empty() ? nullptr : kp.data(), vals.empty() ? nullptr : vals.data(), (int32_t)kp.size(), params));
        check(mpcc_cuda_set_params(h_, params, 1));
    }
    mpcc_cuda_handle* h_ = nullptr;
    PathToJson path_; ParamValue ctor_pv_;
    double obs_[4] = {3, 3, 3, 0}, cur_u_[8] = {0};
    std::vector<OptVariables> guess_;
    std::vector<double> rb_;
};
}  // namespace mpcc
