// Shared POD types (host + device) of the batched MPCC SQP path.
// Names follow the reference (cpp/include/config.h, types.h, Params/params.h).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MPCC_HD __host__ __device__ __forceinline__
#define MPCC_HDN inline __host__ __device__
#define MPCC_HDNI __host__ __device__ __noinline__   // one copy in the kernel image: keeps the instruction footprint of the hot loops small
#else
#define MPCC_HD inline
#define MPCC_HDN inline
#define MPCC_HDNI inline
#endif

namespace mpcc {

constexpr int NX = 9;          // [q1..q7, s, vs]            config.h:31
constexpr int NU = 8;          // [dq1..dq7, dVs]            config.h:32
constexpr int NPC = 11;        // polytopic rows per stage   config.h:34
constexpr int DOF = 7;
constexpr int NLINKS = 9;      // robot_model.h:13
constexpr int N_SPLINE = 100;  // config.h:38
constexpr double INF = 1e30;   // config.h:37
constexpr int MAX_N = 64;      // largest horizon a handle may be created with

// solver_interface.h:28-42, numeric order preserved
enum Status : int32_t {
    SOLVED = 0, MAX_ITER_EXCEEDED, QP_DualInfeasibleInaccurate, QP_PrimalInfeasibleInaccurate,
    QP_SolvedInaccurate, QP_MaxIterReached, QP_PrimalInfeasible, QP_DualInfeasible, Sigint,
    INVALID_SETTINGS, NAN_HESSIAN, NON_PD_HESSIAN
};

// One instance's parameter set = the six reference JSON files flattened.
// (params.h:32-247; the per-key order is the one the C-ABI documents.)
struct Params {
    // model.json
    double max_dist_proj, desired_ee_velocity, s_trust_region, deacc_ratio, tol_sing, tol_selcol, tol_envcol;
    // cost.json
    double q_c, q_c_N_mult, q_l, q_vs, q_ori, q_sing, r_dq, r_ddq, r_dVs, q_c_red_ratio, q_l_inc_ratio, q_ori_red_ratio;
    // bounds.json
    double lx[NX], ux[NX], lu[NU], uu[NU], ldd[DOF], udd[DOF];
    // normalization.json (diagonals of T_x, T_u)
    double Tx[NX], Tu[NU];
    // sqp.json
    double eps_prim, eps_dual;
    double max_iter, line_search_max_iter, do_SOC, use_BFGS;  // stored as doubles; integers in value
    double line_search_tau, line_search_eta, line_search_rho;
    // r_ddq as seen by the solver interface (file value; osqp_interface.cpp:28,57)
    double r_ddq_solver;
};
constexpr int PARAMS_DOUBLES = sizeof(Params) / sizeof(double);

// Fitted track (ArcLengthSpline after fitSpline) as a flat table of doubles.
//   knots s_i, cubic coefficients of X/Y/Z, knot rotations R_i, and per segment the
//   rotation-spline data: w_i = Log(R_i^T R_{i+1})^vee, c_i = 3/h^2, d_i = -2/h^3.
struct TrackTable {
    double s[N_SPLINE];
    double a[3][N_SPLINE], b[3][N_SPLINE], c[3][N_SPLINE], d[3][N_SPLINE];  // b,d use the first 99 entries
    double R[N_SPLINE][9];
    double w[N_SPLINE][3];  // first 99
    double rc[N_SPLINE], rd[N_SPLINE];
    double delta, length;
    double pad[2];
};
constexpr int TRACK_DOUBLES = sizeof(TrackTable) / sizeof(double);

// RobotData (robot_data.h:11-94) as 150 doubles per (instance, stage):
//   q7 | p3 | R9 | Jv21 | Jw21 | manip | dmanip7 | sel | dsel7 | obs_r | env9 | denv63
constexpr int RB_Q = 0, RB_P = 7, RB_R = 10, RB_JV = 19, RB_JW = 40, RB_MANIP = 61, RB_DMANIP = 62, RB_SEL = 69,
              RB_DSEL = 70, RB_OBSR = 77, RB_ENV = 78, RB_DENV = 87, RB_DOUBLES = 150;

// Options of the structured QP solver that replaces the reference's OSQP call.
struct QpOptions { int max_iter; double eps; };
// Interior-point start: slacks t = max(h - G z0, QP_INIT_SLACK), multipliers lam = QP_INIT_SLACK / t (so t lam <= mu0 = QP_INIT_SLACK,
// on the central path wherever the slack is not clamped).  In the solver's normalised variables the step QPs of a warm MPCC
// cycle have slacks of 0.01 .. 1: starting at mu0 = 3e-3 instead of t >= 1, lam = 1 saves ~2 of ~8 interior-point
// iterations (measured on closed-loop QPs of configs C2 and C3 and on widely perturbed starts, no failure in either).
// Sweep on the bench workload (k_sqp_warp, ms): 1e-2 8.15 | 3e-3 7.96 | 1e-3 8.78 | 5e-4 10.6.
#ifndef MPCC_QP_INIT_SLACK
#define MPCC_QP_INIT_SLACK 3e-3
#endif
constexpr double QP_INIT_SLACK = MPCC_QP_INIT_SLACK;
// Fraction to the boundary of the interior-point step: tau = max(0.995, 1 - QP_TAU_GAIN mu).  With the constant 0.995 every iteration shrinks
// mu and the dual residual by exactly 200 once the steps are full (3e-3 -> 7.6e-5 -> 6.5e-7 -> 3.3e-9 -> 1.6e-11: four iterations per warm QP);
// letting tau follow 1 - mu (the usual rule: IPOPT's tau = max(tau_min, 1 - mu)) makes the tail superlinear: three iterations.  Gain 30: tau leaves 0.995 only below mu = 1.7e-4, so the hard QPs of a start-up transient (whose mu
// stays large for many iterations) behave as before (measured on the host build: gain 1 costs them 40 % more iterations, gain 30 none, gain 1000 does
// not reach three iterations).  MPCC_QP_TAU_GAIN <= 0: the constant.
#ifndef MPCC_QP_TAU_GAIN
#define MPCC_QP_TAU_GAIN 30.0
#endif
constexpr double QP_TAU_GAIN = MPCC_QP_TAU_GAIN;
MPCC_HD double qp_step_tau(double mu) { const double t = 1.0 - QP_TAU_GAIN * mu; return (QP_TAU_GAIN > 0.0 && t > 0.995) ? t : 0.995; }

}  // namespace mpcc
