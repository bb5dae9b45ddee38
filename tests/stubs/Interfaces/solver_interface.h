// TEST INFRASTRUCTURE ONLY: the smallest stand-in for the reference-side declarations that examples/cuda_sqp_interface.h is
// written against (Eigen is not in this image), so that the binding a maintainer would add is compile-checked here.
// Shapes follow the reference: solver_interface.h:25-54, types.h:33-150, osqp_interface.h:48-79, arc_length_spline.h:49-56,
// config.h:31-38,79.  No behaviour lives here.
#pragma once
#include <iostream>
#include <map>
#include <string>
#include <vector>
namespace Eigen {
struct Vector3d { double v[3]; double operator()(int i) const { return v[i]; } };
struct VectorXd { std::vector<double> v; const double* data() const { return v.data(); } long size() const { return (long)v.size(); } double operator()(int i) const { return v[i]; } };
struct Matrix3d { double m[9]; double operator()(int r, int c) const { return m[3 * c + r]; } };   // column-major like Eigen's default
}
namespace mpcc {
static const int N = 10, NX = 9, NU = 8;
static const std::string pkg_path = "./";
struct State { double q1, q2, q3, q4, q5, q6, q7, s, vs; };
struct Input { double dq1, dq2, dq3, dq4, dq5, dq6, dq7, dVs; };
struct OptVariables { State xk; Input uk; };
struct ComputeTime { double set_env, set_qp, solve_qp, get_alpha, total; };
struct PathToJson { std::string param_path, cost_path, bounds_path, track_path, normalization_path, sqp_path; };
struct ParamValue { std::map<std::string, double> param, cost, bounds, track, normalization, sqp; };
struct PathData { Eigen::VectorXd X, Y, Z; std::vector<Eigen::Matrix3d> R; Eigen::VectorXd s; int n_points; };
class ArcLengthSpline { public: PathData getPathData() const { return path_; } PathData path_; };
enum Status { SOLVED, MAX_ITER_EXCEEDED, QP_DualInfeasibleInaccurate, QP_PrimalInfeasibleInaccurate, QP_SolvedInaccurate, QP_MaxIterReached,
              QP_PrimalInfeasible, QP_DualInfeasible, Sigint, INVALID_SETTINGS, NAN_HESSIAN, NON_PD_HESSIAN };
class SolverInterface {
public:
    virtual void setTrack(const ArcLengthSpline track) = 0;
    virtual void setParam(const ParamValue& param_value) = 0;
    virtual void setEnvData(const Eigen::Vector3d& obs_position, const double& obs_radius) = 0;
    virtual void setInitialGuess(const std::vector<OptVariables>& initial_guess) = 0;
    virtual void setCurrentInput(const Input& cutrent_input) = 0;
    virtual bool solveOCP(std::vector<OptVariables>& opt_sol, Status* status, ComputeTime* mpc_time) = 0;
    virtual ~SolverInterface() {}
};
}  // namespace mpcc
