// ============================================================================
// TEST INFRASTRUCTURE ONLY.  CPU oracle: a dependency-free C++17 restatement of
// the per-control-cycle SQP solve of JunHeonYoon/MPCC_manipulator
// (mpcc::MPC::runMPC and everything below it).  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it,
// and only as the checker.  The product library never links or calls this.
//
// Every function cites the reference file:line it follows (paths relative to
// the reference's cpp/ directory).  Differences from the reference, all forced:
//   * horizon N is a runtime value (reference: compile-time 10, config.h:36);
//   * RBDL is replaced by the explicit Panda chain it encodes
//     (robot_model.cpp:68-319), Eigen by oracle/linalg.hpp;
//   * OSQP (un-vendored, un-pinned; osqp_interface.cpp:592-656) is replaced by
//     a generic dense primal-dual interior-point QP solver run to ~1e-9
//     residuals with no wall-clock limit.  QP/SQP parity is therefore
//     "unpinned" by the reference (it has no test on that path either);
//   * duals (lambda_) are not tracked: they only feed BFGS (off by default).
// Pinning status: kinematics pinned by the constants the reference embeds
// (see tests/test_oracle_pins.py); MLP/cost/constraints by the reference's own
// property tests; QP/SQP: parity unpinned.
// ============================================================================
#pragma once
#include "linalg.hpp"
#include <map>
#include <limits>
#include <cstdio>
#include <chrono>

namespace orc {

constexpr int NX = 9;    // config.h:31
constexpr int NU = 8;    // config.h:32
constexpr int NPC = 11;  // config.h:34
constexpr int DOF = 7;
constexpr int NLINKS = 9;      // robot_model.h:13
constexpr int N_SPLINE = 100;  // config.h:38
constexpr double INF = 1e30;   // config.h:37

// solver_interface.h:28-42 (numeric order preserved)
enum Status {
    SOLVED = 0, MAX_ITER_EXCEEDED, QP_DualInfeasibleInaccurate, QP_PrimalInfeasibleInaccurate,
    QP_SolvedInaccurate, QP_MaxIterReached, QP_PrimalInfeasible, QP_DualInfeasible, Sigint,
    INVALID_SETTINGS, NAN_HESSIAN, NON_PD_HESSIAN
};

// ---- Params (params.h:32-247; params.cpp) ----------------------------------
struct Param { double max_dist_proj, desired_ee_velocity, s_trust_region, deacc_ratio, tol_sing, tol_selcol, tol_envcol; };
struct CostParam { double q_c, q_c_N_mult, q_l, q_vs, q_ori, q_sing, r_dq, r_ddq, r_dVs, q_c_red_ratio, q_l_inc_ratio, q_ori_red_ratio; };
struct BoundsParam { double lx[NX], ux[NX], lu[NU], uu[NU], ldd[DOF], udd[DOF]; };
struct NormParam { double Tx[NX], Tu[NU]; };
struct SQPParam { double eps_prim, eps_dual; int max_iter, line_search_max_iter; bool do_SOC, use_BFGS; double tau, eta, rho; };

struct State { double v[NX]; double& s() { return v[7]; } double& vs() { return v[8]; } double s() const { return v[7]; } double vs() const { return v[8]; } };
struct Input { double v[NU]; };
struct OptVariables { State xk; Input uk; };

// ---- Robot kinematics (robot_model.cpp:68-319,366-450) ---------------------
struct RobotModel {
    // joint placements r_i (robot_model.cpp:170-182) and parent->child rotations
    // E_i (robot_model.cpp:184-263); child orientation in parent = E_i^T (RBDL).
    static void chain(const double* q, Vec3 p[8], Mat3 R[8], Vec3& p_ee, Mat3& R_ee) {
        static const double r[8][3] = {{0, 0, 0}, {0, 0, 0.333}, {0, 0, 0}, {0, -0.316, 0}, {0.0825, 0, 0},
                                       {-0.0825, 0.384, 0}, {0, 0, 0}, {0.088, 0, 0}};
        static const Mat3 Ea = {{1, 0, 0, 0, 0, -1, 0, 1, 0}};
        static const Mat3 Eb = {{1, 0, 0, 0, 0, 1, 0, -1, 0}};
        static const Mat3 Ei = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
        const Mat3* E[8] = {&Ei, &Ei, &Ea, &Eb, &Eb, &Ea, &Eb, &Eb};
        p[0] = {0, 0, 0};
        R[0] = Mat3::Identity();
        for (int i = 1; i <= 7; i++) {
            Vec3 ri = {r[i][0], r[i][1], r[i][2]};
            p[i] = add(p[i - 1], mul(R[i - 1], ri));
            double c = std::cos(q[i - 1]), s = std::sin(q[i - 1]);
            Mat3 Rz = {{c, -s, 0, s, c, 0, 0, 0, 1}};
            R[i] = mul(mul(R[i - 1], transpose(*E[i])), Rz);
        }
        // link7 -> hand (fixed, literal 0.707107: robot_model.cpp:238-242) -> hand_tcp (fixed)
        static const Mat3 E8 = {{0.707107, -0.707107, 0, 0.707107, 0.707107, 0, 0, 0, 1}};
        R_ee = mul(R[7], transpose(E8));
        Vec3 rh = {0, 0, 0.107}, rt = {0, 0, 0.1034};
        p_ee = add(add(p[7], mul(R[7], rh)), mul(R_ee, rt));
    }
    static Vec3 getEEPosition(const double* q) { Vec3 p[8]; Mat3 R[8]; Vec3 pe; Mat3 Re; chain(q, p, R, pe, Re); return pe; }
    static Mat3 getEEOrientation(const double* q) { Vec3 p[8]; Mat3 R[8]; Vec3 pe; Mat3 Re; chain(q, p, R, pe, Re); return Re; }
    // J = [Jv; Jw], 6x7 row-major (robot_model.cpp:366-377)
    static void getJacobian(const double* q, double J[42]) {
        Vec3 p[8]; Mat3 R[8]; Vec3 pe; Mat3 Re;
        chain(q, p, R, pe, Re);
        for (int i = 1; i <= 7; i++) {
            Vec3 z = {R[i](0, 2), R[i](1, 2), R[i](2, 2)};
            Vec3 lin = cross(z, sub(pe, p[i]));
            for (int a = 0; a < 3; a++) { J[a * 7 + (i - 1)] = lin[a]; J[(3 + a) * 7 + (i - 1)] = z[a]; }
        }
    }
    // robot_model.cpp:431-435
    static double getManipulability(const double* q) {
        double J[42], JJt[36];
        getJacobian(q, J);
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) {
                double s = 0;
                for (int k = 0; k < 7; k++) s += J[a * 7 + k] * J[b * 7 + k];
                JJt[a * 6 + b] = s;
            }
        return std::sqrt(det_lu(JJt, 6));
    }
    // robot_model.cpp:437-450 (central differences, delta = 1e-4)
    static void getDManipulability(const double* q, double d[7]) {
        const double delta = 1e-4;
        for (int i = 0; i < 7; i++) {
            double qp[7], qm[7];
            for (int k = 0; k < 7; k++) { qp[k] = q[k]; qm[k] = q[k]; }
            qp[i] = q[i] + delta;
            qm[i] = q[i] - delta;
            double m1 = getManipulability(qp), m2 = getManipulability(qm);
            d[i] = (m1 - m2) / (2 * delta);
        }
    }
};

// ---- MLP distance fields (SelfCollisionModel.cpp:75-250; Env clone) --------
struct MLP {
    int n_input = 0, n_output = 0, n_layer = 0;
    std::vector<int> n_hidden;
    bool is_nerf = true;
    std::vector<Mat> weight;   // weight[l] is (out x in), row per output neuron (:25-31)
    std::vector<Vec> bias;
    // setNeuralNetwork (:131-138) with weights handed over instead of read from text
    void set(int n_in, int n_out, const std::vector<int>& hidden, const double* w_flat, const double* b_flat) {
        n_input = n_in; n_output = n_out; n_hidden = hidden; n_layer = (int)hidden.size() + 1;
        weight.resize(n_layer); bias.resize(n_layer);
        size_t wo = 0, bo = 0;
        for (int i = 0; i < n_layer; i++) {
            int in = (i == 0) ? (is_nerf ? 3 * n_in : n_in) : hidden[i - 1];
            int out = (i == n_layer - 1) ? n_out : hidden[i];
            weight[i].setZero(out, in);
            bias[i].assign(out, 0.0);
            for (int k = 0; k < out * in; k++) weight[i].a[k] = w_flat[wo + k];
            for (int k = 0; k < out; k++) bias[i][k] = b_flat[bo + k];
            wo += (size_t)out * in; bo += out;
        }
    }
    // calculateMlpOutput (:140-250): output (n_output) and d output / d input (n_output x n_input),
    // as written: dense encoding Jacobian, all n_input tangent columns.
    void eval(const double* input, double* output, double* out_jac) const {
        const int ni = n_input, ne = 3 * n_input;
        Vec in_nerf(ne);
        for (int i = 0; i < ni; i++) { in_nerf[i] = input[i]; in_nerf[ni + i] = std::sin(input[i]); in_nerf[2 * ni + i] = std::cos(input[i]); }
        Vec hidden, hidden_next;
        Mat temp, temp_next, hd;
        for (int layer = 0; layer < n_layer; layer++) {
            const Mat& W = weight[layer];
            if (layer == 0) {
                hidden.assign(W.r, 0.0);
                hd.setZero(W.r, ne);
                for (int h = 0; h < W.r; h++) {
                    double s = 0;
                    for (int k = 0; k < ne; k++) s += W(h, k) * in_nerf[k];
                    s += bias[0][h];
                    double dr = (s > 0) ? 1.0 : 0.0;  // ReLU_derivative (SelfCollisionModel.h:66-69)
                    for (int k = 0; k < ne; k++) hd(h, k) = dr * W(h, k);
                    hidden[h] = std::max(0.0, s);
                }
                Mat nerf_jac(ne, ni);  // (:179-184)
                for (int i = 0; i < ni; i++) {
                    nerf_jac(i, i) = 1.0;
                    nerf_jac(ni + i, i) = std::cos(input[i]);
                    nerf_jac(2 * ni + i, i) = -std::sin(input[i]);
                }
                temp.setZero(W.r, ni);
                for (int h = 0; h < W.r; h++)
                    for (int k = 0; k < ne; k++) {
                        double w = hd(h, k);
                        const double* nj = nerf_jac.row(k);
                        double* t = temp.row(h);
                        for (int c = 0; c < ni; c++) t[c] += w * nj[c];
                    }
            } else if (layer == n_layer - 1) {
                for (int o = 0; o < W.r; o++) {
                    double s = 0;
                    for (int k = 0; k < W.c; k++) s += W(o, k) * hidden[k];
                    output[o] = s + bias[layer][o];
                    double* oj = out_jac + (size_t)o * ni;
                    for (int c = 0; c < ni; c++) oj[c] = 0;
                    for (int k = 0; k < W.c; k++) {
                        double w = W(o, k);
                        const double* t = temp.row(k);
                        for (int c = 0; c < ni; c++) oj[c] += w * t[c];
                    }
                }
            } else {
                hidden_next.assign(W.r, 0.0);
                temp_next.setZero(W.r, ni);
                for (int h = 0; h < W.r; h++) {
                    double s = 0;
                    for (int k = 0; k < W.c; k++) s += W(h, k) * hidden[k];
                    s += bias[layer][h];
                    hidden_next[h] = std::max(0.0, s);
                    // hidden_derivative.row(h) = ReLU'(s) * W.row(h), then the dense product
                    // with the running Jacobian (:207-219) -- inactive rows are multiplied too
                    double dr = (s > 0) ? 1.0 : 0.0;
                    double* tn = temp_next.row(h);
                    for (int k = 0; k < W.c; k++) {
                        double w = dr * W(h, k);
                        const double* t = temp.row(k);
                        for (int c = 0; c < ni; c++) tn[c] += w * t[c];
                    }
                }
                hidden.swap(hidden_next);
                std::swap(temp, temp_next);
            }
        }
    }
};

// ---- RobotData (robot_data.h:11-94) ----------------------------------------
struct RobotData {
    double q[DOF];
    Vec3 EE_position;
    Mat3 EE_orientation;
    double J[42];  // [Jv; Jw] row-major 6x7
    double manipul, d_manipul[DOF];
    double sel_min_dist, d_sel_min_dist[DOF];
    double obs_radius;
    double env_min_dist[NLINKS], d_env_min_dist[NLINKS * DOF];
    bool is_data_valid = false, is_env_data_valid = false;
    const double* Jv() const { return J; }
    const double* Jw() const { return J + 21; }
    void update(const double* q_in, const MLP& selcol) {  // robot_data.h:55-71
        for (int i = 0; i < DOF; i++) q[i] = q_in[i];
        EE_position = RobotModel::getEEPosition(q);
        EE_orientation = RobotModel::getEEOrientation(q);
        RobotModel::getJacobian(q, J);
        manipul = RobotModel::getManipulability(q);
        RobotModel::getDManipulability(q, d_manipul);
        double out[1], jac[DOF];
        selcol.eval(q, out, jac);
        sel_min_dist = out[0];
        for (int i = 0; i < DOF; i++) d_sel_min_dist[i] = jac[i];
        is_data_valid = true;
    }
    void updateEnv(const double obs[3], double radius, const MLP& envcol) {  // robot_data.h:74-88
        obs_radius = radius;
        double in[DOF + 3], out[NLINKS], jac[NLINKS * (DOF + 3)];
        for (int i = 0; i < DOF; i++) in[i] = q[i];
        for (int i = 0; i < 3; i++) in[DOF + i] = obs[i];
        envcol.eval(in, out, jac);
        for (int l = 0; l < NLINKS; l++) {
            env_min_dist[l] = out[l];
            for (int j = 0; j < DOF; j++) d_env_min_dist[l * DOF + j] = jac[l * (DOF + 3) + j];
        }
        is_env_data_valid = true;
    }
    // flat export used by the parity tests (150 doubles)
    static constexpr int FLAT = 150;
    void to_flat(double* o) const {
        int k = 0;
        for (int i = 0; i < 7; i++) o[k++] = q[i];
        for (int i = 0; i < 3; i++) o[k++] = EE_position[i];
        for (int i = 0; i < 9; i++) o[k++] = EE_orientation.m[i];
        for (int i = 0; i < 42; i++) o[k++] = J[i];
        o[k++] = manipul;
        for (int i = 0; i < 7; i++) o[k++] = d_manipul[i];
        o[k++] = sel_min_dist;
        for (int i = 0; i < 7; i++) o[k++] = d_sel_min_dist[i];
        o[k++] = obs_radius;
        for (int i = 0; i < 9; i++) o[k++] = env_min_dist[i];
        for (int i = 0; i < 63; i++) o[k++] = d_env_min_dist[i];
    }
    void from_flat(const double* o) {
        int k = 0;
        for (int i = 0; i < 7; i++) q[i] = o[k++];
        for (int i = 0; i < 3; i++) EE_position[i] = o[k++];
        for (int i = 0; i < 9; i++) EE_orientation.m[i] = o[k++];
        for (int i = 0; i < 42; i++) J[i] = o[k++];
        manipul = o[k++];
        for (int i = 0; i < 7; i++) d_manipul[i] = o[k++];
        sel_min_dist = o[k++];
        for (int i = 0; i < 7; i++) d_sel_min_dist[i] = o[k++];
        obs_radius = o[k++];
        for (int i = 0; i < 9; i++) env_min_dist[i] = o[k++];
        for (int i = 0; i < 63; i++) d_env_min_dist[i] = o[k++];
        is_data_valid = is_env_data_valid = true;
    }
};

// ---- SO(3) helpers (cubic_spline_rot.cpp:25-95) ----------------------------
inline Mat3 getSkewMatrix(const Vec3& v) {
    Mat3 r = Mat3::Zero();
    r(0, 1) = -v[2]; r(0, 2) = v[1]; r(1, 0) = v[2]; r(1, 2) = -v[0]; r(2, 0) = -v[1]; r(2, 1) = v[0];
    return r;
}
inline Vec3 getInverseSkewVector(const Mat3& R) { return {R(2, 1), R(0, 2), R(1, 0)}; }

// Unit eigenvector of a symmetric 3x3 for the eigenvalue closest to 1 (used only
// in the rotation-by-pi branch).  Eigen's SelfAdjointEigenSolver returns it with
// an implementation-defined sign; here the sign makes the largest component
// positive.  That branch is therefore sign-unpinned (documented in DESIGN.md).
inline Vec3 eigvec_for_one(const Mat3& R) {
    // (R + I)/2 = v v^T for a rotation by pi; take the column with the largest diagonal.
    double d[3] = {(R(0, 0) + 1) / 2, (R(1, 1) + 1) / 2, (R(2, 2) + 1) / 2};
    int c = 0;
    if (d[1] > d[c]) c = 1;
    if (d[2] > d[c]) c = 2;
    Vec3 v = {(R(0, c) + (c == 0)) / 2, (R(1, c) + (c == 1)) / 2, (R(2, c) + (c == 2)) / 2};
    double n = norm(v);
    v = scale(v, 1.0 / n);
    int big = 0;
    if (std::fabs(v[1]) > std::fabs(v[big])) big = 1;
    if (std::fabs(v[2]) > std::fabs(v[big])) big = 2;
    if (v[big] < 0) v = scale(v, -1.0);
    return v;
}

inline Mat3 LogMatrix(const Mat3& R) {  // cubic_spline_rot.cpp:44-79
    double tr = trace(R);
    if (std::fabs(tr + 1.0) < 1e-6) {
        Vec3 v = eigvec_for_one(R);
        return scale(getSkewMatrix(v), -M_PI);
    } else if (std::fabs(tr - 3.0) < 1e-6) {
        return Mat3::Zero();
    }
    double th = std::acos((tr - 1.0) / 2.0);
    return scale(sub(R, transpose(R)), 1.0 / 2.0 * th / std::sin(th));
}
inline Mat3 ExpMatrix(const Mat3& sk) {  // cubic_spline_rot.cpp:81-95
    Vec3 v = getInverseSkewVector(sk);
    double vn = norm(v);
    if (vn <= 1e-8) {
        // "1/2" is integer division in the reference (:92) -> the quadratic term vanishes
        return add(Mat3::Identity(), scale(sk, std::cos(vn)));
    }
    Mat3 sk2 = mul(sk, sk);
    return add(add(Mat3::Identity(), scale(sk, std::sin(vn) / vn)), scale(sk2, (1 - std::cos(vn)) / (vn * vn)));
}

// ---- CubicSpline (cubic_spline.cpp) -----------------------------------------
struct CubicSpline {
    Vec x_data, y_data, a, b, c, d;
    int n_points = 0;
    bool is_regular = false;
    double delta_x = 0;
    std::map<double, int> x_map;
    void genSpline(const Vec& x_in, const Vec& y_in, bool regular) {  // :160-183
        x_data = x_in; y_data = y_in; n_points = (int)x_in.size(); is_regular = regular;
        x_map.clear();
        if (regular) delta_x = x_in[1] - x_in[0];
        else { delta_x = 0; for (int i = 0; i < n_points; i++) x_map[x_in[i]] = i; }
        compSplineParams();
    }
    void compSplineParams() {  // :65-124
        const int n = n_points;
        a = y_data; b.assign(n - 1, 0); c.assign(n, 0); d.assign(n - 1, 0);
        Vec mu(n - 1, 0), h(n - 1, 0), alpha(n - 1, 0), l(n, 0), z(n, 0);
        for (int i = 0; i < n - 1; i++) h[i] = x_data[i + 1] - x_data[i];
        for (int i = 1; i < n - 1; i++) alpha[i] = 3.0 / h[i] * (a[i + 1] - a[i]) - 3.0 / h[i - 1] * (a[i] - a[i - 1]);
        l[0] = 1.0; mu[0] = 0.0; z[0] = 0.0;
        for (int i = 1; i < n - 1; i++) {
            l[i] = 2.0 * (x_data[i + 1] - x_data[i - 1]) - h[i - 1] * mu[i - 1];
            mu[i] = h[i] / l[i];
            z[i] = (alpha[i] - h[i - 1] * z[i - 1]) / l[i];
        }
        l[n - 1] = 1.0; z[n - 1] = 0.0; c[n - 1] = 0.0;
        for (int i = n - 2; i >= 0; i--) {
            c[i] = z[i] - mu[i] * c[i + 1];
            b[i] = (a[i + 1] - a[i]) / h[i] - (h[i] * (c[i + 1] + 2.0 * c[i])) / 3.0;
            d[i] = (c[i + 1] - c[i]) / (3.0 * h[i]);
        }
    }
    int getIndex(double x) const {  // :126-153
        if (x == x_data[n_points - 1]) return n_points - 1;
        if (is_regular) return int(std::floor(x / delta_x));
        auto it = x_map.upper_bound(x);
        if (it == x_map.end()) return -1;
        return it->second - 1;
    }
    double unwrapInput(double x) const { return std::max(0., std::min(x, x_data[n_points - 1])); }  // :155-160
    double getPoint(double x) const {  // :185-206
        x = unwrapInput(x);
        int index = getIndex(x);
        double dx = x - x_data[index];
        if (index == n_points - 1) return y_data[n_points - 1];
        return a[index] + b[index] * dx + c[index] * dx * dx + d[index] * (dx * (dx * dx));
    }
    double getDerivative(double x) const {  // :208-226
        x = unwrapInput(x);
        int index = getIndex(x);
        double dx = x - x_data[index];
        if (index == n_points - 1) return 0.;
        return b[index] + 2.0 * c[index] * dx + 3.0 * d[index] * (dx * dx);
    }
    double getSecondDerivative(double x) const {  // :228-246
        x = unwrapInput(x);
        int index = getIndex(x);
        double dx = x - x_data[index];
        if (index == n_points - 1) return 2.0 * c[index];
        return 2.0 * c[index] + 6.0 * d[index] * dx;
    }
};

// ---- CubicSplineRot (cubic_spline_rot.cpp:97-259) ---------------------------
struct CubicSplineRot {
    Vec x_data, c, d;
    std::vector<Mat3> R_data;
    int n_points = 0;
    bool is_regular = false;
    double delta_x = 0;
    std::map<double, int> x_map;
    void genSpline(const Vec& x_in, const std::vector<Mat3>& R_in, bool regular) {  // :191-214
        x_data = x_in; R_data = R_in; n_points = (int)x_in.size(); is_regular = regular;
        x_map.clear();
        if (regular) delta_x = x_in[1] - x_in[0];
        else { delta_x = 0; for (int i = 0; i < n_points; i++) x_map[x_in[i]] = i; }
        c.assign(n_points - 1, 0); d.assign(n_points - 1, 0);
        for (int i = 0; i < n_points - 1; i++) {  // :139-156
            double h = x_data[i + 1] - x_data[i];
            c[i] = 3.0 / std::pow(h, 2);
            d[i] = -2.0 / std::pow(h, 3);
        }
    }
    int getIndex(double x) const {  // :158-182
        if (x == x_data[n_points - 1]) return n_points - 1;
        if (is_regular) return int(std::floor(x / delta_x));
        auto it = x_map.upper_bound(x);
        if (it == x_map.end()) return -1;
        return it->second - 1;
    }
    double unwrapInput(double x) const { return std::max(0., std::min(x, x_data[n_points - 1])); }
    Mat3 getPoint(double x) const {  // :216-238
        x = unwrapInput(x);
        int index = getIndex(x);
        if (index == n_points - 1) return R_data[n_points - 1];
        double dx = x - x_data[index], dx2 = dx * dx, dx3 = dx * dx2;
        Mat3 log_RR = LogMatrix(mul(transpose(R_data[index]), R_data[index + 1]));
        return mul(R_data[index], ExpMatrix(scale(log_RR, c[index] * dx2 + d[index] * dx3)));
    }
    Vec3 getDerivative(double x) const {  // :240-259
        x = unwrapInput(x);
        int index = getIndex(x);
        if (index == n_points - 1) return {0, 0, 0};
        double dx = x - x_data[index], dx2 = dx * dx;
        Vec3 Log_RR = getInverseSkewVector(LogMatrix(mul(transpose(R_data[index]), R_data[index + 1])));
        return scale(Log_RR, 2.0 * c[index] * dx + 3.0 * d[index] * dx2);
    }
};

// ---- ArcLengthSpline (arc_length_spline.cpp) --------------------------------
struct PathData { Vec X, Y, Z, s; std::vector<Mat3> R; int n_points = 0; };

// Eigen's setLinSpaced(size, low, high) for |high| >= |low| (low + i*step, last = high).
inline Vec linspaced(int n, double low, double high) {
    Vec v(n);
    double step = (high - low) / (n - 1);
    for (int i = 0; i < n; i++) v[i] = (i == n - 1) ? high : low + i * step;
    return v;
}

struct ArcLengthSpline {
    PathData path_data;
    CubicSpline spline_x, spline_y, spline_z;
    CubicSplineRot spline_r;
    double max_dist_proj = 0.03;  // param_.max_dist_proj

    static Vec compArcLength(const Vec& X, const Vec& Y, const Vec& Z) {  // :66-87
        int n = (int)X.size();
        Vec s(n, 0.0);
        for (int i = 0; i < n - 1; i++) {
            double dx = X[i + 1] - X[i], dy = Y[i + 1] - Y[i], dz = Z[i + 1] - Z[i];
            s[i + 1] = s[i] + std::sqrt(dx * dx + dy * dy + dz * dz);
        }
        return s;
    }
    static PathData resamplePath(const CubicSpline& sx, const CubicSpline& sy, const CubicSpline& sz,
                                 const CubicSplineRot& sr, double total) {  // :89-118
        PathData p;
        p.n_points = N_SPLINE;
        p.s = linspaced(N_SPLINE, 0, total);
        p.X.assign(N_SPLINE, 0); p.Y.assign(N_SPLINE, 0); p.Z.assign(N_SPLINE, 0); p.R.resize(N_SPLINE);
        for (int i = 0; i < N_SPLINE; i++) {
            p.X[i] = sx.getPoint(p.s[i]); p.Y[i] = sy.getPoint(p.s[i]); p.Z[i] = sz.getPoint(p.s[i]);
            p.R[i] = sr.getPoint(p.s[i]);
        }
        return p;
    }
    void gen6DSpline(const Vec& X, const Vec& Y, const Vec& Z, const std::vector<Mat3>& R) { fitSpline(X, Y, Z, R); }  // :255-265
    void fitSpline(const Vec& X, const Vec& Y, const Vec& Z, const std::vector<Mat3>& R) {  // :213-253
        Vec s_approx = compArcLength(X, Y, Z);
        double total = s_approx.back();
        CubicSpline fx, fy, fz, gx, gy, gz;
        CubicSplineRot fr, gr;
        fx.genSpline(s_approx, X, false); fy.genSpline(s_approx, Y, false); fz.genSpline(s_approx, Z, false);
        fr.genSpline(s_approx, R, false);
        PathData first = resamplePath(fx, fy, fz, fr, total);
        s_approx = compArcLength(first.X, first.Y, first.Z);
        total = s_approx.back();
        gx.genSpline(s_approx, first.X, false); gy.genSpline(s_approx, first.Y, false); gz.genSpline(s_approx, first.Z, false);
        gr.genSpline(s_approx, first.R, false);
        PathData second = resamplePath(gx, gy, gz, gr, total);
        path_data = second;  // setRegularData (:50-64)
        spline_x.genSpline(path_data.s, path_data.X, true);
        spline_y.genSpline(path_data.s, path_data.Y, true);
        spline_z.genSpline(path_data.s, path_data.Z, true);
        spline_r.genSpline(path_data.s, path_data.R, true);
    }
    Vec3 getPosition(double s) const { return {spline_x.getPoint(s), spline_y.getPoint(s), spline_z.getPoint(s)}; }                    // :267-275
    Mat3 getOrientation(double s) const { return spline_r.getPoint(s); }                                                             // :277-283
    Vec3 getDerivative(double s) const { return {spline_x.getDerivative(s), spline_y.getDerivative(s), spline_z.getDerivative(s)}; }   // :285-293
    Vec3 getOrientationDerivative(double s) const { return spline_r.getDerivative(s); }                                              // :295-301
    Vec3 getSecondDerivative(double s) const { return {spline_x.getSecondDerivative(s), spline_y.getSecondDerivative(s), spline_z.getSecondDerivative(s)}; }  // :303-311
    double getLength() const { return path_data.s[path_data.n_points - 1]; }                                                         // :313-316
    double unwrapInput(double x) const { return std::max(0., std::min(x, getLength())); }

    double projectOnSpline(double s, const Vec3& ee_pos) const {  // :318-379
        double s_guess = s;
        Vec3 pos_path = getPosition(s_guess);
        double s_opt = s_guess;
        double dist = norm(sub(ee_pos, pos_path));
        const int n = path_data.n_points;
        if (dist >= max_dist_proj) {
            int min_all = 0, min_valid = -1;
            double best_all = std::numeric_limits<double>::infinity(), best_valid = std::numeric_limits<double>::infinity();
            for (int i = 0; i < n; i++) {
                double dx = path_data.X[i] - ee_pos[0], dy = path_data.Y[i] - ee_pos[1], dz = path_data.Z[i] - ee_pos[2];
                double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best_all) { best_all = d2; min_all = i; }
                bool valid = std::fabs(path_data.s[i] - s_guess) <= max_dist_proj;
                if (valid && d2 < best_valid) { best_valid = d2; min_valid = i; }
            }
            s_opt = (min_valid < 0) ? path_data.s[min_all] : path_data.s[min_valid];
        }
        if (s_opt >= path_data.s[n - 1]) return path_data.s[n - 1];
        double s_old = s_opt;
        for (int i = 0; i < 20; i++) {
            pos_path = getPosition(s_opt);
            Vec3 ds = getDerivative(s_opt), dds = getSecondDerivative(s_opt);
            Vec3 diff = sub(pos_path, ee_pos);
            double jac = 2.0 * diff[0] * ds[0] + 2.0 * diff[1] * ds[1] + 2.0 * diff[2] * ds[2];
            double hessian = 2.0 * ds[0] * ds[0] + 2.0 * diff[0] * dds[0] + 2.0 * ds[1] * ds[1] + 2.0 * diff[1] * dds[1] +
                             2.0 * ds[2] * ds[2] + 2.0 * diff[2] * dds[2];
            s_opt -= jac / hessian;
            s_opt = unwrapInput(s_opt);
            if (std::fabs(s_old - s_opt) <= 1e-5) return s_opt;
            s_old = s_opt;
        }
        return s_guess;
    }
};

// ---- Model / Integrator (model.cpp:31-124; integrator.cpp:29-68) ------------
struct Model {
    static void getF(const State& x, const Input& u, double f[NX]) {  // model.cpp:31-45
        for (int i = 0; i < 7; i++) f[i] = u.v[i];
        f[7] = x.v[8];
        f[8] = u.v[7];
    }
    // getLinModel (:117-124): ZOH discretisation of the constant, nilpotent pair
    // (A_c, B_c) of getModelJacobian (:47-65).  expm of the 18x18 matrix (:67-91)
    // truncates exactly after the quadratic term: A_d = I + Ts e_s e_vs^T,
    // B_d(q,dq) = Ts I, B_d(vs,dVs) = Ts, B_d(s,dVs) = Ts^2/2, g_d = 0.
    static void getLinModel(double Ts, double A[NX * NX], double B[NX * NU], double g[NX]) {
        for (int i = 0; i < NX * NX; i++) A[i] = 0;
        for (int i = 0; i < NX * NU; i++) B[i] = 0;
        for (int i = 0; i < NX; i++) { A[i * NX + i] = 1.0; g[i] = 0; }
        A[7 * NX + 8] = Ts;
        for (int i = 0; i < 7; i++) B[i * NU + i] = Ts;
        B[8 * NU + 7] = Ts;
        B[7 * NU + 7] = 0.5 * Ts * Ts;
    }
};
struct Integrator {
    static State RK4(const State& x, const Input& u, double ts) {  // integrator.cpp:29-43
        double k1[NX], k2[NX], k3[NX], k4[NX];
        State t;
        Model::getF(x, u, k1);
        for (int i = 0; i < NX; i++) t.v[i] = x.v[i] + ts / 2. * k1[i];
        Model::getF(t, u, k2);
        for (int i = 0; i < NX; i++) t.v[i] = x.v[i] + ts / 2. * k2[i];
        Model::getF(t, u, k3);
        for (int i = 0; i < NX; i++) t.v[i] = x.v[i] + ts * k3[i];
        Model::getF(t, u, k4);
        State r;
        for (int i = 0; i < NX; i++) r.v[i] = x.v[i] + ts * (k1[i] / 6. + k2[i] / 3. + k3[i] / 3. + k4[i] / 6.);
        return r;
    }
    static State simTimeStep(const State& x, const Input& u, double ts) {  // integrator.cpp:55-68; fine step 1 ms (integrator.h:53)
        const double fine = 0.001;
        State xn = x;
        const int steps = (int)(ts / fine);
        for (int i = 0; i < steps; i++) xn = RK4(xn, u, fine);
        return xn;
    }
};

// ---- Cost (cost.cpp) --------------------------------------------------------
struct CostGrad { double f_x[NX], f_u[NU]; void setZero() { for (double& v : f_x) v = 0; for (double& v : f_u) v = 0; } };
struct CostHess {
    double f_xx[NX * NX], f_uu[NU * NU], f_xu[NX * NU];
    void setZero() { for (double& v : f_xx) v = 0; for (double& v : f_uu) v = 0; for (double& v : f_xu) v = 0; }
};

struct Cost {
    CostParam cost_param;
    Param param;
    double contouring_cost_, lag_cost_, heading_cost_;
    int N;

    static double CubicSplineFn(double x, double x_0, double x_f, double y_0, double y_f) {  // :36-43
        double t = (x - x_0) / (x_f - x_0);
        double t2 = std::pow(t, 2), t3 = std::pow(t, 3);
        return y_0 + (y_f - y_0) * (3 * t2 - 2 * t3);
    }
    struct ErrorInfo { Vec3 contouring_error, lag_error; double d_contouring_error[3 * NX], d_lag_error[3 * NX]; };

    ErrorInfo getErrorInfo(const ArcLengthSpline& track, const State& x, const RobotData& rb) const {  // :82-117
        ErrorInfo e;
        const double s = x.s();
        Vec3 pos_ref = track.getPosition(s), dpos = track.getDerivative(s), ddpos = track.getSecondDerivative(s);
        Vec3 Tangent = dpos;
        Vec3 Normal = {ddpos[0], ddpos[1], ddpos[1]};  // ddz_ref = ddpos_ref(1) (:65)
        Vec3 total_error = sub(rb.EE_position, pos_ref);
        Vec3 lag_error = scale(Tangent, dot(Tangent, total_error));
        Vec3 contouring_error = sub(total_error, lag_error);
        double d_total[3 * NX];
        for (double& v : d_total) v = 0;
        for (int a = 0; a < 3; a++) {
            for (int j = 0; j < DOF; j++) d_total[a * NX + j] = rb.Jv()[a * 7 + j];
            d_total[a * NX + 7] = -Tangent[a];
        }
        // d_lag = (T T^T) d_total + (T e^T + |e_l| I) d_Tangent, d_Tangent only has the s column (:104-109)
        double M[9];
        double ln = norm(lag_error);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) M[a * 3 + b] = Tangent[a] * total_error[b] + (a == b ? ln : 0.0);
        for (int a = 0; a < 3; a++)
            for (int c = 0; c < NX; c++) {
                double v = 0;
                for (int b = 0; b < 3; b++) v += Tangent[a] * Tangent[b] * d_total[b * NX + c];
                if (c == 7) for (int b = 0; b < 3; b++) v += M[a * 3 + b] * Normal[b];
                e.d_lag_error[a * NX + c] = v;
                e.d_contouring_error[a * NX + c] = d_total[a * NX + c] - v;
            }
        e.contouring_error = contouring_error;
        e.lag_error = lag_error;
        return e;
    }
    void getContouringCost(const ArcLengthSpline& track, const State& x, const RobotData& rb, int k, double* obj, CostGrad* grad, CostHess* hess) const {  // :119-162
        ErrorInfo e = getErrorInfo(track, x, rb);
        double cc0 = k < N ? contouring_cost_ : cost_param.q_c_N_mult * contouring_cost_;
        double cc1 = lag_cost_;
        double s_max = track.getLength();
        double desired = (x.s() < s_max * param.deacc_ratio) ? param.desired_ee_velocity
                                                             : -param.desired_ee_velocity / (s_max * param.deacc_ratio) * (x.s() - s_max);
        if (obj) *obj = cc0 * dot(e.contouring_error, e.contouring_error) + cc1 * dot(e.lag_error, e.lag_error) + cost_param.q_vs * std::pow(x.vs() - desired, 2);
        if (grad) {
            grad->setZero();
            for (int c = 0; c < NX; c++) {
                double g = 0;
                for (int a = 0; a < 3; a++) g += 2.0 * cc0 * e.d_contouring_error[a * NX + c] * e.contouring_error[a] + 2.0 * cc1 * e.d_lag_error[a * NX + c] * e.lag_error[a];
                grad->f_x[c] = g;
            }
            grad->f_x[8] += 2.0 * cost_param.q_vs * (x.vs() - desired);
        }
        if (hess) {
            hess->setZero();
            for (int r = 0; r < NX; r++)
                for (int c = 0; c < NX; c++) {
                    double h = 0;
                    for (int a = 0; a < 3; a++) h += 2.0 * cc0 * e.d_contouring_error[a * NX + r] * e.d_contouring_error[a * NX + c] + 2.0 * cc1 * e.d_lag_error[a * NX + r] * e.d_lag_error[a * NX + c];
                    hess->f_xx[r * NX + c] = h;
                }
            hess->f_xx[8 * NX + 8] += 2.0 * cost_param.q_vs;
        }
    }
    void getHeadingCost(const ArcLengthSpline& track, const State& x, const RobotData& rb, double* obj, CostGrad* grad, CostHess* hess) const {  // :164-207
        Mat3 ref_R = track.getOrientation(x.s());
        Vec3 dR_ref = track.getOrientationDerivative(x.s());
        const Mat3& cur_R = rb.EE_orientation;
        Mat3 R_bar = mul(transpose(ref_R), cur_R);
        Vec3 Log_R_bar = getInverseSkewVector(LogMatrix(R_bar));
        double d_Log[3 * NX];
        for (double& v : d_Log) v = 0;
        if (obj) *obj = heading_cost_ * dot(Log_R_bar, Log_R_bar);
        if (grad || hess) {
            Mat3 J_r_inv;
            double n = norm(Log_R_bar);
            if (n < 1e-8) J_r_inv = Mat3::Identity();
            else {
                Mat3 sk = getSkewMatrix(Log_R_bar);
                // "+" where the closed form has "-" (:188), reproduced
                double coef = 1. / dot(Log_R_bar, Log_R_bar) + (1. + std::cos(n)) / (2. * n * std::sin(n));
                J_r_inv = add(add(Mat3::Identity(), scale(sk, 1. / 2.)), scale(mul(sk, sk), coef));
            }
            Mat3 M = mul(J_r_inv, transpose(cur_R));
            for (int a = 0; a < 3; a++) {
                for (int j = 0; j < DOF; j++) {
                    double v = 0;
                    for (int b = 0; b < 3; b++) v += M(a, b) * rb.Jw()[b * 7 + j];
                    d_Log[a * NX + j] = v;
                }
                double v = 0;
                for (int b = 0; b < 3; b++) v += M(a, b) * dR_ref[b];
                d_Log[a * NX + 7] = -v;
            }
        }
        if (grad) {
            grad->setZero();
            for (int c = 0; c < NX; c++) {
                double g = 0;
                for (int a = 0; a < 3; a++) g += d_Log[a * NX + c] * Log_R_bar[a];
                grad->f_x[c] = 2.0 * heading_cost_ * g;
            }
        }
        if (hess) {
            hess->setZero();
            for (int r = 0; r < NX; r++)
                for (int c = 0; c < NX; c++) {
                    double h = 0;
                    for (int a = 0; a < 3; a++) h += d_Log[a * NX + r] * d_Log[a * NX + c];
                    hess->f_xx[r * NX + c] = 2.0 * heading_cost_ * h;
                }
        }
    }
    void getInputCost(const State&, const Input& u, int k, double* obj, CostGrad* grad, CostHess* hess) const {  // :209-270
        if (obj) {
            *obj = 0;
            if (k != N) {
                double sq = 0;
                for (int i = 0; i < DOF; i++) sq += u.v[i] * u.v[i];
                *obj = cost_param.r_dq * sq + cost_param.r_dVs * std::pow(u.v[7], 2);
            }
        }
        if (grad) {
            grad->setZero();
            if (k != N) {
                for (int i = 0; i < DOF; i++) grad->f_u[i] = 2.0 * cost_param.r_dq * u.v[i];
                grad->f_u[7] = 2.0 * cost_param.r_dVs * u.v[7];
            }
        }
        if (hess) {
            hess->setZero();
            if (k != N) {
                for (int i = 0; i < DOF; i++) hess->f_uu[i * NU + i] = 2.0 * cost_param.r_dq;
                hess->f_uu[7 * NU + 7] = 2.0 * cost_param.r_dVs;
            }
        }
    }
    void getSingularityCost(const RobotData& rb, double* obj, CostGrad* grad, CostHess* hess) const {  // :272-288
        if (obj) *obj = -cost_param.q_sing * rb.manipul;
        if (grad) { grad->setZero(); for (int i = 0; i < DOF; i++) grad->f_x[i] = -cost_param.q_sing * rb.d_manipul[i]; }
        if (hess) hess->setZero();
    }
    void getCost(const ArcLengthSpline& track, const State& x, const Input& u, const RobotData& rb, int k, double* obj, CostGrad* grad, CostHess* hess) {  // :290-357
        double ratio = std::min(rb.sel_min_dist / (param.tol_selcol * 2.0), rb.manipul / (param.tol_sing * 2.0));
        if (ratio <= 1.0) {
            contouring_cost_ = cost_param.q_c * CubicSplineFn(ratio, 0.5, 1.0, cost_param.q_c_red_ratio, 1.0);
            lag_cost_ = cost_param.q_l * CubicSplineFn(ratio, 0.5, 1.0, cost_param.q_l_inc_ratio, 1.0);
            heading_cost_ = cost_param.q_ori * CubicSplineFn(ratio, 0.5, 1.0, cost_param.q_ori_red_ratio, 1.0);
        } else {
            contouring_cost_ = cost_param.q_c; lag_cost_ = cost_param.q_l; heading_cost_ = cost_param.q_ori;
        }
        double oc = 0, oh = 0, oi = 0, os = 0;
        CostGrad gc, gh, gi, gs;
        CostHess hc, hh, hi, hs;
        getContouringCost(track, x, rb, k, &oc, grad ? &gc : nullptr, hess ? &hc : nullptr);
        getHeadingCost(track, x, rb, &oh, grad ? &gh : nullptr, hess ? &hh : nullptr);
        getInputCost(x, u, k, &oi, grad ? &gi : nullptr, hess ? &hi : nullptr);
        getSingularityCost(rb, &os, grad ? &gs : nullptr, hess ? &hs : nullptr);
        if (obj) *obj = oc + oh + oi + os;
        if (grad) {
            for (int i = 0; i < NX; i++) grad->f_x[i] = gc.f_x[i] + gh.f_x[i] + gi.f_x[i] + gs.f_x[i];
            for (int i = 0; i < NU; i++) grad->f_u[i] = gc.f_u[i] + gh.f_u[i] + gi.f_u[i] + gs.f_u[i];
        }
        if (hess) {
            for (int i = 0; i < NX * NX; i++) hess->f_xx[i] = hc.f_xx[i] + hh.f_xx[i] + hi.f_xx[i] + hs.f_xx[i];
            for (int i = 0; i < NU * NU; i++) hess->f_uu[i] = hc.f_uu[i] + hh.f_uu[i] + hi.f_uu[i] + hs.f_uu[i];
            for (int i = 0; i < NX * NU; i++) hess->f_xu[i] = hc.f_xu[i] + hh.f_xu[i] + hi.f_xu[i] + hs.f_xu[i];
            for (int i = 0; i < NX; i++) hess->f_xx[i * NX + i] += 1e-6;
            for (int i = 0; i < NU; i++) hess->f_uu[i * NU + i] += 1e-6;
        }
    }
};

// ---- Constraints (constraints.cpp) ------------------------------------------
inline double getRBF(double delta, double h) {  // :34-43
    if (h >= delta) return -std::log(h + 1);
    return -std::log(delta + 1) - 1 / (delta + 1) * (h - delta) + 1 / (2 * std::pow(delta + 1, 2)) * std::pow(h - delta, 2);
}
inline double getDRBF(double delta, double h) {  // :52-61
    if (h >= delta) return -1 / (h + 1);
    return -1 / (delta + 1) + 1 / (std::pow(delta + 1, 2)) * (h - delta);
}
struct ConstraintsInfo { double c_vec[NPC], c_lvec[NPC], c_uvec[NPC]; };
struct ConstraintsJac { double c_x[NPC * NX], c_u[NPC * NU]; };

struct Constraints {
    Param param;
    int N;
    // one scalar row: -grad_h . dq + RBF(h) <= 0 (:70-147 pattern)
    void row(double h, const double* d_h, const Input& u, int k, int j, ConstraintsInfo* ci, ConstraintsJac* cj) const {
        const double delta = -0.5;
        if (ci) {
            ci->c_vec[j] = 0; ci->c_lvec[j] = 0; ci->c_uvec[j] = 0;
            if (k != N) {
                double dotp = 0;
                for (int i = 0; i < DOF; i++) dotp += d_h[i] * u.v[i];
                ci->c_lvec[j] = -INF; ci->c_uvec[j] = 0.0;
                ci->c_vec[j] = -dotp + getRBF(delta, h);
            }
        }
        if (cj) {
            for (int i = 0; i < NX; i++) cj->c_x[j * NX + i] = 0;
            for (int i = 0; i < NU; i++) cj->c_u[j * NU + i] = 0;
            if (k != N) {
                double d_RBF = getDRBF(delta, h);
                for (int i = 0; i < DOF; i++) { cj->c_x[j * NX + i] = d_RBF * d_h[i]; cj->c_u[j * NU + i] = -d_h[i]; }
            }
        }
    }
    void getConstraints(const State&, const Input& u, const RobotData& rb, int k, ConstraintsInfo* ci, ConstraintsJac* cj) const {  // :192-243
        double d[DOF];
        // self collision (:70-108), [cm] -> [m]
        for (int i = 0; i < DOF; i++) d[i] = 0.01 * rb.d_sel_min_dist[i];
        row(0.01 * rb.sel_min_dist - param.tol_selcol * 0.01, d, u, k, 0, ci, cj);
        // singularity (:110-147)
        row(rb.manipul - param.tol_sing, rb.d_manipul, u, k, 1, ci, cj);
        // env collision (:149-190)
        for (int l = 0; l < NLINKS; l++) {
            for (int i = 0; i < DOF; i++) d[i] = 0.01 * rb.d_env_min_dist[l * DOF + i];
            double md = 0.01 * (rb.env_min_dist[l] - rb.obs_radius * 1.2);
            row(md - 0.01 * param.tol_envcol, d, u, k, 2 + l, ci, cj);
        }
    }
};

// ---- Bounds (bounds.cpp:85-128) ---------------------------------------------
struct Bounds {
    BoundsParam b;
    Param param;
    void getBoundsLX(const State& x, double* o) const { for (int i = 0; i < NX; i++) o[i] = b.lx[i]; o[7] = std::max(x.s() - param.s_trust_region, 0.); }
    void getBoundsUX(const State& x, double L, double* o) const { for (int i = 0; i < NX; i++) o[i] = b.ux[i]; o[7] = std::min(x.s() + param.s_trust_region, L); }
};

// ---- generic dense convex QP solver (replaces OSQP; see file header) --------
// min 1/2 z'Pz + q'z  s.t.  l <= A z <= u.   Primal-dual interior point
// (Mehrotra predictor-corrector).  Generic presolve only: all-zero rows are
// checked and dropped; rows with l == u are equalities; |bound| >= 1e20 is
// infinite.  No knowledge of the OCP structure is used.
struct QPResult { bool ok = false; int iters = 0; double res_prim = 0, res_dual = 0, gap = 0; };

struct DenseQP {
    int max_iter = 100;
    double eps = 1e-9;

    QPResult solve(const Mat& P, const Vec& q, const Mat& A, const Vec& l, const Vec& u, Vec& z_out) const {
        const int n = P.r, mtot = A.r;
        QPResult res;
        // --- classify rows -----------------------------------------------------
        std::vector<int> eq_rows;                   // equality rows
        struct Ineq { int row; double sign; double h; };  // sign*a z <= h
        std::vector<Ineq> ineq;
        // sparse row storage
        std::vector<std::vector<std::pair<int, double>>> rows(mtot);
        for (int i = 0; i < mtot; i++) {
            const double* ai = A.row(i);
            for (int j = 0; j < n; j++) if (ai[j] != 0.0) rows[i].push_back({j, ai[j]});
            bool lo_inf = l[i] <= -1e20, up_inf = u[i] >= 1e20;
            if (rows[i].empty()) {
                if ((!lo_inf && l[i] > 1e-9) || (!up_inf && u[i] < -1e-9)) return res;  // 0 not in [l,u]: infeasible
                continue;
            }
            if (!lo_inf && !up_inf && l[i] == u[i]) { eq_rows.push_back(i); continue; }
            if (!lo_inf && !up_inf && l[i] > u[i]) return res;  // trivially infeasible
            if (!up_inf) ineq.push_back({i, 1.0, u[i]});
            if (!lo_inf) ineq.push_back({i, -1.0, -l[i]});
        }
        const int ne = (int)eq_rows.size(), m = (int)ineq.size();
        Vec z(n, 0.0), y(ne, 0.0), t(m), lam(m);
        auto Gz = [&](const Vec& zz, int k) {
            double s = 0;
            for (auto& e : rows[ineq[k].row]) s += e.second * zz[e.first];
            return ineq[k].sign * s;
        };
        for (int k = 0; k < m; k++) { t[k] = std::max(ineq[k].h - Gz(z, k), 1.0); lam[k] = 1.0; }

        Mat H(n, n), Y(ne, n), S(ne, ne);
        Vec rd(n), re(ne), rp(m), rc(m), rhs(n), w(m), dz(n), dy(ne), dt(m), dlam(m), dz_a(n), dt_a(m), dlam_a(m), tmp(ne);
        bool factored_ok = true;

        auto residuals = [&]() {
            for (int i = 0; i < n; i++) {
                const double* Pi = P.row(i);
                double s = q[i];
                for (int j = 0; j < n; j++) s += Pi[j] * z[j];
                rd[i] = s;
            }
            for (int k = 0; k < ne; k++)
                for (auto& e : rows[eq_rows[k]]) rd[e.first] += e.second * y[k];
            for (int k = 0; k < m; k++)
                for (auto& e : rows[ineq[k].row]) rd[e.first] += ineq[k].sign * e.second * lam[k];
            for (int k = 0; k < ne; k++) {
                double s = 0;
                for (auto& e : rows[eq_rows[k]]) s += e.second * z[e.first];
                re[k] = s - l[eq_rows[k]];
            }
            for (int k = 0; k < m; k++) rp[k] = Gz(z, k) + t[k] - ineq[k].h;
        };
        // Solve the reduced KKT system for a given complementarity residual rc:
        //   (P + G'WG) dz + Ae' dy = -rd - G'[(lam.*rp - rc)/t],  Ae dz = -re
        auto kkt_solve = [&](const Vec& rcv, Vec& dzv, Vec& dyv, Vec& dtv, Vec& dlamv) {
            for (int i = 0; i < n; i++) rhs[i] = -rd[i];
            for (int k = 0; k < m; k++) {
                double v = (lam[k] * rp[k] - rcv[k]) / t[k];
                for (auto& e : rows[ineq[k].row]) rhs[e.first] -= ineq[k].sign * e.second * v;
            }
            // dz0 = H^-1 rhs
            dzv = rhs;
            cholesky_solve(H.a.data(), n, n, dzv.data());
            if (ne > 0) {
                // S dy = Ae dz0 + re ; dz = dz0 - H^-1 Ae' dy
                for (int k = 0; k < ne; k++) {
                    double s = re[k];
                    for (auto& e : rows[eq_rows[k]]) s += e.second * dzv[e.first];
                    dyv[k] = s;
                }
                cholesky_solve(S.a.data(), ne, ne, dyv.data());
                Vec corr(n, 0.0);
                for (int k = 0; k < ne; k++)
                    for (auto& e : rows[eq_rows[k]]) corr[e.first] += e.second * dyv[k];
                cholesky_solve(H.a.data(), n, n, corr.data());
                for (int i = 0; i < n; i++) dzv[i] -= corr[i];
            }
            for (int k = 0; k < m; k++) {
                double g = 0;
                for (auto& e : rows[ineq[k].row]) g += e.second * dzv[e.first];
                g *= ineq[k].sign;
                dtv[k] = -rp[k] - g;
                dlamv[k] = (-rcv[k] - lam[k] * dtv[k]) / t[k];
            }
        };
        auto max_step = [&](const Vec& dtv, const Vec& dlamv) {
            double a = 1.0;
            for (int k = 0; k < m; k++) {
                if (dtv[k] < 0) a = std::min(a, -t[k] / dtv[k]);
                if (dlamv[k] < 0) a = std::min(a, -lam[k] / dlamv[k]);
            }
            return a;
        };

        double qn = 0;
        for (int i = 0; i < n; i++) qn = std::max(qn, std::fabs(q[i]));
        for (int it = 0; it < max_iter; it++) {
            residuals();
            double mu = 0;
            for (int k = 0; k < m; k++) mu += t[k] * lam[k];
            mu = (m > 0) ? mu / m : 0.0;
            double nrd = 0, nre = 0, nrp = 0;
            for (int i = 0; i < n; i++) nrd = std::max(nrd, std::fabs(rd[i]));
            for (int k = 0; k < ne; k++) nre = std::max(nre, std::fabs(re[k]));
            for (int k = 0; k < m; k++) nrp = std::max(nrp, std::fabs(rp[k]));
            res.iters = it; res.res_dual = nrd; res.res_prim = std::max(nre, nrp); res.gap = mu;
            if (nrd <= eps * (1.0 + qn) && nre <= eps && nrp <= eps && mu <= eps) { res.ok = true; break; }
            if (!(nrd == nrd) || !(mu == mu)) break;  // NaN
            // --- factor H = P + G' W G and S = Ae H^-1 Ae' ---------------------
            H = P;
            for (int k = 0; k < m; k++) {
                double wk = lam[k] / t[k];
                auto& rk = rows[ineq[k].row];
                for (auto& e1 : rk) {
                    double* Hr = H.row(e1.first);
                    double f = wk * e1.second;
                    for (auto& e2 : rk) Hr[e2.first] += f * e2.second;
                }
            }
            factored_ok = cholesky_inplace(H.a.data(), n, n);
            if (!factored_ok) break;
            if (ne > 0) {
                // Y = L^-1 Ae'  (stored as ne x n, row k = L^-1 a_k)
                for (int k = 0; k < ne; k++) {
                    double* yk = Y.row(k);
                    for (int j = 0; j < n; j++) yk[j] = 0;
                    for (auto& e : rows[eq_rows[k]]) yk[e.first] = e.second;
                    cholesky_forward(H.a.data(), n, n, yk);
                }
                for (int a = 0; a < ne; a++)
                    for (int b = 0; b <= a; b++) {
                        const double* ya = Y.row(a);
                        const double* yb = Y.row(b);
                        double s = 0;
                        for (int j = 0; j < n; j++) s += ya[j] * yb[j];
                        S(a, b) = s; S(b, a) = s;
                    }
                if (!cholesky_inplace(S.a.data(), ne, ne)) break;
            }
            // --- predictor -----------------------------------------------------
            for (int k = 0; k < m; k++) rc[k] = lam[k] * t[k];
            kkt_solve(rc, dz_a, dy, dt_a, dlam_a);
            double a_aff = max_step(dt_a, dlam_a);
            double mu_aff = 0;
            for (int k = 0; k < m; k++) mu_aff += (t[k] + a_aff * dt_a[k]) * (lam[k] + a_aff * dlam_a[k]);
            mu_aff = (m > 0) ? mu_aff / m : 0.0;
            double sigma = (mu > 0) ? std::pow(mu_aff / mu, 3) : 0.0;
            // --- corrector -----------------------------------------------------
            for (int k = 0; k < m; k++) rc[k] = lam[k] * t[k] + dt_a[k] * dlam_a[k] - sigma * mu;
            kkt_solve(rc, dz, dy, dt, dlam);
            double a = std::min(1.0, 0.995 * max_step(dt, dlam));
            for (int i = 0; i < n; i++) z[i] += a * dz[i];
            for (int k = 0; k < ne; k++) y[k] += a * dy[k];
            for (int k = 0; k < m; k++) { t[k] += a * dt[k]; lam[k] += a * dlam[k]; }
            res.iters = it + 1;
        }
        if (res.ok) z_out = z;
        return res;
    }
};

// ---- OsqpInterface restated (osqp_interface.cpp) ----------------------------
struct FilterData { double obj, gap_vio; };
struct ComputeTime { double set_env = 0, set_qp = 0, solve_qp = 0, get_alpha = 0, total = 0; };

struct SqpLog {  // per-iteration record for the parity tests
    std::vector<Vec> steps;       // normalised QP step of each SQP iteration (N_var)
    std::vector<double> alphas;
    std::vector<int> qp_ok, qp_iters;
};
// Robustness of one filter comparison "trial dominated by entry" (osqp_interface.cpp:778-783): the smallest
// perturbation of (obj, gap) that flips it; obj scaled by 1+|obj|.  The parity tests use it to recognise
// accept/reject decisions that hinge on solver noise (both sides of the gap comparison are ~1e-9 numbers).
inline double filter_margin(double o2, double g2, double fo, double fg) {
    double so = 1.0 + std::fabs(fo);
    double d_o = (o2 - fo) / so, d_g = g2 - fg;
    if (d_o >= 0 && d_g >= 0) return std::min(d_o, d_g);
    return std::max(d_o < 0 ? -d_o : 0.0, d_g < 0 ? -d_g : 0.0);
}

struct Solver {
    int N;
    double Ts;
    Cost cost;
    Constraints constraints;
    Bounds bounds;
    NormParam norm;
    SQPParam sqp;
    double r_ddq;  // cost_param_.r_ddq of the interface (file value only, osqp_interface.cpp:28,57)
    const MLP* selcolNN = nullptr;
    const MLP* envcolNN = nullptr;
    ArcLengthSpline track;
    DenseQP qp;

    int N_var, N_eq, N_ineqb, N_ineqp, N_constr;  // osqp_interface.h:113-117
    std::vector<RobotData> rb;
    Input current_u;
    std::vector<OptVariables> initial_guess;
    Vec initial_guess_vec, step;
    std::vector<FilterData> filter;
    int sqp_iter = 0;
    SqpLog* log = nullptr;
    double min_filter_margin = 1e300;  // over every filter comparison of the last solveOCP (test instrumentation)
    // Test instrumentation for "parity modulo certified filter ties": per SQP iteration of the last solveOCP, the
    // line search's own first-trial decision (1 = accepted) and the robustness margin of that decision; and an
    // optional list of decisions to FOLLOW instead (so the oracle can be replayed along the device's branch).
    std::vector<int> natural_accept;
    std::vector<double> accept_margin;
    std::vector<int> forced_accept;
    int n_soc_fail = 0;  // failed second-order-correction QPs of the last solveOCP

    void init(int N_) {
        N = N_;
        N_var = (N + 1) * NX + N * NU; N_eq = (N + 1) * NX; N_ineqb = N_var + N * NU; N_ineqp = (N + 1) * NPC;
        N_constr = N_eq + N_ineqb + N_ineqp;
        cost.N = N; constraints.N = N;
        rb.assign(N + 1, RobotData());
        initial_guess.assign(N + 1, OptVariables());
    }
    void setTrack(const ArcLengthSpline& t) { track = t; }  // :90-93
    void setEnvData(const double obs[3], double radius) { for (int i = 0; i <= N; i++) rb[i].updateEnv(obs, radius, *envcolNN); }  // :102-108
    void setInitialGuess(const std::vector<OptVariables>& g) {  // :110-122
        initial_guess = g;
        initial_guess_vec = OptvarToVector(g);
        for (int i = 0; i <= N; i++) rb[i].update(g[i].xk.v, *selcolNN);
    }
    void setCurrentInput(const Input& u) { current_u = u; }  // :124-127

    Vec OptvarToVector(const std::vector<OptVariables>& g) const {  // :847-857
        Vec v(N_var, 0.0);
        for (int i = 0; i <= N; i++) {
            for (int m = 0; m < NX; m++) v[NX * i + m] = g[i].xk.v[m];
            if (i != N) for (int k = 0; k < NU; k++) v[NX * (N + 1) + NU * i + k] = g[i].uk.v[k];
        }
        return v;
    }
    std::vector<OptVariables> vectorToOptvar(const Vec& v) const {  // :835-845 (resize value-initialises: uk[N] = 0)
        std::vector<OptVariables> g(N + 1);
        for (int i = 0; i <= N; i++) {
            for (int m = 0; m < NX; m++) g[i].xk.v[m] = v[NX * i + m];
            for (int k = 0; k < NU; k++) g[i].uk.v[k] = (i != N) ? v[NX * (N + 1) + NU * i + k] : 0.0;
        }
        return g;
    }
    Vec deNormalizeStep(const Vec& s) const {  // :859-869
        Vec d(N_var, 0.0);
        for (int i = 0; i <= N; i++) {
            for (int m = 0; m < NX; m++) d[NX * i + m] = norm.Tx[m] * s[NX * i + m];
            if (i != N) for (int k = 0; k < NU; k++) d[NX * (N + 1) + NU * i + k] = norm.Tu[k] * s[NX * (N + 1) + NU * i + k];
        }
        return d;
    }

    void setCost(const std::vector<OptVariables>& g, double* obj, Vec* grad_obj, Mat* hess_obj) {  // :129-219
        if (obj) *obj = 0;
        if (grad_obj) grad_obj->assign(N_var, 0.0);
        if (hess_obj) hess_obj->setZero(N_var, N_var);
        const int U0 = NX * (N + 1);
        for (int i = 0; i <= N; i++) {
            double obj_k;
            CostGrad gk;
            CostHess hk;
            cost.getCost(track, g[i].xk, g[i].uk, rb[i], i, &obj_k, grad_obj ? &gk : nullptr, hess_obj ? &hk : nullptr);
            if (obj) *obj += obj_k;
            if (grad_obj) for (int m = 0; m < NX; m++) (*grad_obj)[NX * i + m] = norm.Tx[m] * gk.f_x[m];
            if (hess_obj)
                for (int a = 0; a < NX; a++)
                    for (int b = 0; b < NX; b++) (*hess_obj)(NX * i + a, NX * i + b) = norm.Tx[a] * hk.f_xx[a * NX + b] * norm.Tx[b];
            if (i != N) {
                if (grad_obj) for (int k = 0; k < NU; k++) (*grad_obj)[U0 + NU * i + k] = norm.Tu[k] * gk.f_u[k];
                if (hess_obj) {
                    for (int a = 0; a < NU; a++)
                        for (int b = 0; b < NU; b++) (*hess_obj)(U0 + NU * i + a, U0 + NU * i + b) = norm.Tu[a] * hk.f_uu[a * NU + b] * norm.Tu[b];
                    for (int a = 0; a < NX; a++)
                        for (int b = 0; b < NU; b++) {
                            double v = norm.Tx[a] * hk.f_xu[a * NU + b] * norm.Tu[b];
                            (*hess_obj)(NX * i + a, U0 + NU * i + b) = v;
                            (*hess_obj)(U0 + NU * i + b, NX * i + a) = v;
                        }
                }
                // joint-acceleration cost (:167-217)
                if (obj && i != N - 1) {
                    double sq = 0;
                    for (int j = 0; j < DOF; j++) { double dv = g[i + 1].uk.v[j] - g[i].uk.v[j]; sq += dv * dv; }
                    *obj += r_ddq * sq;
                }
                if (grad_obj) {
                    for (int j = 0; j < DOF; j++) {
                        double gg;
                        if (i == 0) gg = 2. * r_ddq * (g[i].uk.v[j] - g[i + 1].uk.v[j]);
                        else if (i == N - 1) gg = 2. * r_ddq * (g[i].uk.v[j] - g[i - 1].uk.v[j]);
                        else gg = 2. * r_ddq * (2. * g[i].uk.v[j] - g[i + 1].uk.v[j] - g[i - 1].uk.v[j]);
                        (*grad_obj)[U0 + NU * i + j] += norm.Tu[j] * gg;
                    }
                }
                if (hess_obj) {
                    for (int j = 0; j < DOF; j++) {
                        double ii = (i == 0 || i == N - 1) ? 2. * r_ddq : 4. * r_ddq;
                        (*hess_obj)(U0 + NU * i + j, U0 + NU * i + j) += norm.Tu[j] * ii * norm.Tu[j];
                        if (i != N - 1) {
                            double ij = -2. * r_ddq;
                            (*hess_obj)(U0 + NU * i + j, U0 + NU * (i + 1) + j) += norm.Tu[j] * ij * norm.Tu[j];
                            (*hess_obj)(U0 + NU * (i + 1) + j, U0 + NU * i + j) += norm.Tu[j] * ij * norm.Tu[j];
                        }
                    }
                }
            }
        }
    }

    // setDynamics/setBounds/setPolytopicConstraints/setConstraints (:221-389) written into the
    // row blocks [eq ; box-x ; box-u ; accel ; polytopic]
    void setConstraints(const std::vector<OptVariables>& g, Mat* jac, Vec* constr, Vec* l, Vec* u) {
        if (jac) jac->setZero(N_constr, N_var);
        if (constr) constr->assign(N_constr, 0.0);
        if (l) l->assign(N_constr, 0.0);
        if (u) u->assign(N_constr, 0.0);
        const int U0 = NX * (N + 1);
        double A[NX * NX], B[NX * NU], gd[NX];
        // --- dynamics (:221-252) ---
        for (int i = 0; i <= N; i++) {
            if (i == 0) {
                if (jac) for (int m = 0; m < NX; m++) (*jac)(m, m) = 1.0;
            } else {
                Model::getLinModel(Ts, A, B, gd);
                if (jac) {
                    for (int a = 0; a < NX; a++) {
                        for (int b = 0; b < NX; b++) (*jac)(NX * i + a, NX * (i - 1) + b) = -(1.0 / norm.Tx[a]) * A[a * NX + b] * norm.Tx[b];
                        (*jac)(NX * i + a, NX * i + a) = 1.0;
                        for (int b = 0; b < NU; b++) (*jac)(NX * i + a, U0 + NU * (i - 1) + b) = -(1.0 / norm.Tx[a]) * B[a * NU + b] * norm.Tu[b];
                    }
                }
                if (constr) {
                    for (int a = 0; a < NX; a++) {
                        double pred = gd[a];
                        for (int b = 0; b < NX; b++) pred += A[a * NX + b] * g[i - 1].xk.v[b];
                        for (int b = 0; b < NU; b++) pred += B[a * NU + b] * g[i - 1].uk.v[b];
                        (*constr)[NX * i + a] = (1.0 / norm.Tx[a]) * (g[i].xk.v[a] - pred);
                    }
                }
            }
        }
        // --- bounds (:254-300) ---
        const int B0 = N_eq;
        double lx[NX], ux[NX];
        for (int i = 0; i <= N; i++) {
            bounds.getBoundsLX(g[i].xk, lx);
            bounds.getBoundsUX(g[i].xk, track.getLength(), ux);
            for (int m = 0; m < NX; m++) {
                if (jac) (*jac)(B0 + NX * i + m, NX * i + m) = norm.Tx[m];
                if (constr) (*constr)[B0 + NX * i + m] = g[i].xk.v[m];
                if (l) (*l)[B0 + NX * i + m] = lx[m];
                if (u) (*u)[B0 + NX * i + m] = ux[m];
            }
            if (i != N) {
                for (int k = 0; k < NU; k++) {
                    // input-bound coefficient lands on column NU*i + k (state block), :273
                    if (jac) (*jac)(B0 + U0 + NU * i + k, NU * i + k) = norm.Tu[k];
                    if (constr) (*constr)[B0 + U0 + NU * i + k] = g[i].uk.v[k];
                    if (l) (*l)[B0 + U0 + NU * i + k] = bounds.b.lu[k];
                    if (u) (*u)[B0 + U0 + NU * i + k] = bounds.b.uu[k];
                }
                const int D0 = B0 + U0 + NU * N;
                for (int j = 0; j < DOF; j++) {
                    if (i == 0) {
                        if (jac) (*jac)(D0 + NU * i + j, U0 + NU * i + j) = 1. / Ts * norm.Tu[j];
                        if (constr) (*constr)[D0 + NU * i + j] = 1. / Ts * g[i].uk.v[j];
                        if (l) (*l)[D0 + NU * i + j] = bounds.b.ldd[j] + 1. / Ts * current_u.v[j];
                        if (u) (*u)[D0 + NU * i + j] = bounds.b.udd[j] + 1. / Ts * current_u.v[j];
                    } else {
                        if (jac) {
                            (*jac)(D0 + NU * i + j, U0 + NU * i + j) = 1. / Ts * norm.Tu[j];
                            (*jac)(D0 + NU * i + j, U0 + NU * (i - 1) + j) = -1. / Ts * norm.Tu[j];
                        }
                        if (constr) (*constr)[D0 + NU * i + j] = 1. / Ts * (g[i].uk.v[j] - g[i - 1].uk.v[j]);
                        if (l) (*l)[D0 + NU * i + j] = bounds.b.ldd[j];
                        if (u) (*u)[D0 + NU * i + j] = bounds.b.udd[j];
                    }
                }
            }
        }
        // --- polytopic (:302-344) ---
        const int P0 = N_eq + N_ineqb;
        for (int i = 0; i <= N; i++) {
            ConstraintsInfo ci;
            ConstraintsJac cj;
            constraints.getConstraints(g[i].xk, g[i].uk, rb[i], i, &ci, jac ? &cj : nullptr);
            for (int j = 0; j < NPC; j++) {
                if (jac) {
                    for (int m = 0; m < NX; m++) (*jac)(P0 + NPC * i + j, NX * i + m) = cj.c_x[j * NX + m] * norm.Tx[m];
                    if (i != N) for (int k = 0; k < NU; k++) (*jac)(P0 + NPC * i + j, U0 + NU * i + k) = cj.c_u[j * NU + k] * norm.Tu[k];
                }
                if (constr) (*constr)[P0 + NPC * i + j] = ci.c_vec[j];
                if (l) (*l)[P0 + NPC * i + j] = ci.c_lvec[j];
                if (u) (*u)[P0 + NPC * i + j] = ci.c_uvec[j];
            }
        }
    }
    static double constraint_norm(const Vec& c, const Vec& l, const Vec& u) {  // :824-833
        double s = 0;
        for (size_t i = 0; i < c.size(); i++) s += std::max(l[i] - c[i], 0.0);
        for (size_t i = 0; i < c.size(); i++) s += std::max(c[i] - u[i], 0.0);
        return s;
    }
    double filterLineSearch(const std::vector<OptVariables>& g, const Vec& st) {  // :759-808
        FilterData upd;
        Vec c, l, u;
        bool is_alpha_accepted = true;  // never reset inside the loop (:767)
        double alpha = 1.0;
        Vec base = OptvarToVector(g), dn = deNormalizeStep(st);
        for (int i = 0; i < sqp.line_search_max_iter; i++) {
            Vec v(N_var);
            for (int k = 0; k < N_var; k++) v[k] = base[k] + alpha * dn[k];
            std::vector<OptVariables> ug = vectorToOptvar(v);
            setCost(ug, &upd.obj, nullptr, nullptr);
            setConstraints(ug, nullptr, &c, &l, &u);
            upd.gap_vio = constraint_norm(c, l, u);
            for (size_t j = 0; j < filter.size(); j++)
                min_filter_margin = std::min(min_filter_margin, filter_margin(upd.obj, upd.gap_vio, filter[j].obj, filter[j].gap_vio));
            for (size_t j = 0; j < filter.size(); j++)
                if (upd.obj >= filter[j].obj && upd.gap_vio >= filter[j].gap_vio) { is_alpha_accepted = false; break; }
            if (i == 0) {
                double mg = 1e300;
                for (size_t j = 0; j < filter.size(); j++) mg = std::min(mg, filter_margin(upd.obj, upd.gap_vio, filter[j].obj, filter[j].gap_vio));
                natural_accept.push_back(is_alpha_accepted ? 1 : 0);
                accept_margin.push_back(mg);
                size_t it = natural_accept.size() - 1;
                if (it < forced_accept.size()) is_alpha_accepted = forced_accept[it] != 0;
            }
            if (is_alpha_accepted) {
                std::vector<FilterData> nf;
                for (size_t j = 0; j < filter.size(); j++)
                    if (upd.obj > filter[j].obj || upd.gap_vio > filter[j].gap_vio) nf.push_back(filter[j]);
                nf.push_back(upd);
                filter = nf;
                break;
            } else alpha *= sqp.tau;
        }
        return alpha;
    }
    static bool isNan(const Mat& H) { for (double v : H.a) if (v != v) return true; return false; }

    bool solveOCP(std::vector<OptVariables>& opt_sol, Status* status, ComputeTime* time) {  // :398-590
        auto t0 = std::chrono::high_resolution_clock::now();
        step.assign(N_var, 0.0);
        filter.clear();
        min_filter_margin = 1e300;
        natural_accept.clear(); accept_margin.clear();
        n_soc_fail = 0;
        *time = ComputeTime();
        std::vector<OptVariables> zero_guess(N + 1);
        for (int i = 0; i <= N; i++) { zero_guess[i].xk = initial_guess[0].xk; for (double& v : zero_guess[i].uk.v) v = 0; }
        Mat Hess, jac;
        Vec grad_obj, constr, l, u;
        double obj;
        for (sqp_iter = 0; sqp_iter < sqp.max_iter; sqp_iter++) {
            auto a0 = std::chrono::high_resolution_clock::now();
            setCost(initial_guess, &obj, &grad_obj, &Hess);
            setConstraints(initial_guess, &jac, &constr, &l, &u);
            {  // isPosdef (:810-817): Eigen::LLT reports NumericalIssue only at a pivot x <= 0 (Eigen/src/Cholesky/LLT.h,
               // llt_inplace::unblocked: `if (x <= RealScalar(0)) return k;`), so a NaN pivot passes and the NaN test below decides
                Mat tmp = Hess;
                if (!llt_no_numerical_issue(tmp.a.data(), N_var, N_var)) { *status = NON_PD_HESSIAN; break; }
            }
            if (isNan(Hess)) { *status = NAN_HESSIAN; break; }
            auto a1 = std::chrono::high_resolution_clock::now();
            Vec lq(N_constr), uq(N_constr), z;
            for (int i = 0; i < N_constr; i++) { lq[i] = l[i] - constr[i]; uq[i] = u[i] - constr[i]; }
            QPResult r = qp.solve(Hess, grad_obj, jac, lq, uq, z);
            if (r.ok) step = z;
            else *status = QP_MaxIterReached;  // no break: step keeps its previous value (:479-505)
            int qp_iters_total = r.iters;
            if (sqp.do_SOC) {
                // SecondOrderCorrection (:658-681): the same P, q and A, bounds shifted by d = c(x (+) step) - A step, where (+) adds the
                // NORMALISED step to the unnormalised iterate (:661, no deNormalizeStep); solveQP overwrites step only on success (:646-647),
                // a failure sets the status and the loop goes on (:506-533)
                Vec v = OptvarToVector(initial_guess);
                for (int k = 0; k < N_var; k++) v[k] += step[k];
                std::vector<OptVariables> ug = vectorToOptvar(v);
                Vec cs, ls, us;
                setConstraints(ug, nullptr, &cs, &ls, &us);
                Vec l2(N_constr), u2(N_constr), z2;
                for (int i = 0; i < N_constr; i++) {
                    const double* ai = jac.row(i);
                    double as = 0;
                    for (int k = 0; k < N_var; k++) as += ai[k] * step[k];
                    const double d = cs[i] - as;
                    l2[i] = ls[i] - d; u2[i] = us[i] - d;
                }
                QPResult r2 = qp.solve(Hess, grad_obj, jac, l2, u2, z2);
                if (r2.ok) step = z2;
                else *status = QP_MaxIterReached;
                qp_iters_total += r2.iters;
                n_soc_fail += r2.ok ? 0 : 1;
            }
            auto a2 = std::chrono::high_resolution_clock::now();
            double alpha = filterLineSearch(initial_guess, step);
            auto a3 = std::chrono::high_resolution_clock::now();
            Vec dn = deNormalizeStep(step);
            for (int k = 0; k < N_var; k++) initial_guess_vec[k] += alpha * dn[k];
            initial_guess = vectorToOptvar(initial_guess_vec);
            double inf = 0;
            for (double v : step) inf = std::max(inf, std::fabs(v));
            double primal_step_norm = alpha * inf;
            if (log) { log->steps.push_back(step); log->alphas.push_back(alpha); log->qp_ok.push_back(r.ok); log->qp_iters.push_back(qp_iters_total); }
            time->set_qp += std::chrono::duration<double>(a1 - a0).count();
            time->solve_qp += std::chrono::duration<double>(a2 - a1).count();
            time->get_alpha += std::chrono::duration<double>(a3 - a2).count();
            if (primal_step_norm < sqp.eps_prim) { *status = SOLVED; break; }
        }
        if (sqp_iter == sqp.max_iter) *status = MAX_ITER_EXCEEDED;
        time->total = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        if (*status == SOLVED) { opt_sol = initial_guess; return true; }
        opt_sol = zero_guess;
        return false;
    }
};

// ---- MPC (mpc.cpp) ----------------------------------------------------------
struct MPCReturn { Input u0; std::vector<OptVariables> mpc_horizon; ComputeTime compute_time; Status status; int sqp_iters; };

struct MPC {
    int N;
    double Ts;
    Param param;
    Solver solver;
    ArcLengthSpline track;
    bool valid_initial_guess = false;
    unsigned num_valid_guess_failed = 0;
    std::vector<OptVariables> initial_guess;

    void init(int N_, double Ts_) { N = N_; Ts = Ts_; solver.init(N_); solver.Ts = Ts_; initial_guess.assign(N + 1, OptVariables()); }
    void setTrack(const Vec& X, const Vec& Y, const Vec& Z, const std::vector<Mat3>& R) {  // :192-197
        track.max_dist_proj = param.max_dist_proj;
        track.gen6DSpline(X, Y, Z, R);
        solver.setTrack(track);
        valid_initial_guess = false;
    }
    void unwrapInitialGuess() { double L = track.getLength(); for (int i = 1; i <= N; i++) initial_guess[i].xk.s() = std::min(initial_guess[i].xk.s(), L); }  // :70-77
    void updateInitialGuess(const State& x0) {  // :54-68
        for (int i = 1; i < N; i++) initial_guess[i - 1] = initial_guess[i];
        initial_guess[0].xk = x0;
        initial_guess[N - 1].xk = initial_guess[N - 2].xk;
        initial_guess[N - 1].uk = initial_guess[N - 2].uk;
        initial_guess[N].xk = Integrator::RK4(initial_guess[N - 1].xk, initial_guess[N - 1].uk, Ts);
        for (double& v : initial_guess[N].uk.v) v = 0;
        unwrapInitialGuess();
    }
    void generateNewInitialGuess(const State& x0) {  // :79-89
        for (int i = 0; i <= N; i++) { initial_guess[i].xk = x0; for (double& v : initial_guess[i].uk.v) v = 0; }
        unwrapInitialGuess();
        valid_initial_guess = true;
    }
    bool runMPC_(MPCReturn& ret, State& x0, const Input& u0, const double obs[3], double obs_radius) {  // :104-190
        auto t0 = std::chrono::high_resolution_clock::now();
        double last_s = x0.s();
        x0.s() = track.projectOnSpline(last_s, RobotModel::getEEPosition(x0.v));
        double J[42];
        RobotModel::getJacobian(x0.v, J);
        Vec3 ee_vel = {0, 0, 0};
        for (int a = 0; a < 3; a++) for (int j = 0; j < DOF; j++) ee_vel[a] += J[a * 7 + j] * u0.v[j];
        x0.vs() = dot(ee_vel, track.getDerivative(x0.s()));
        if (std::fabs(last_s - x0.s()) > param.max_dist_proj) { valid_initial_guess = false; num_valid_guess_failed++; }
        if (valid_initial_guess) updateInitialGuess(x0);
        else generateNewInitialGuess(x0);
        solver.setCurrentInput(u0);
        solver.setInitialGuess(initial_guess);
        auto e0 = std::chrono::high_resolution_clock::now();
        solver.setEnvData(obs, obs_radius);
        auto e1 = std::chrono::high_resolution_clock::now();
        Status st = SOLVED;  // reference leaves it uninitialised (mpc.cpp:133); every path below assigns it
        ComputeTime tm;
        solver.solveOCP(initial_guess, &st, &tm);
        if (st == SOLVED) { valid_initial_guess = true; num_valid_guess_failed = 0; }
        else { valid_initial_guess = false; num_valid_guess_failed++; }
        ret.u0 = initial_guess[0].uk;
        ret.mpc_horizon = initial_guess;
        ret.compute_time = tm;
        ret.compute_time.total = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        ret.compute_time.set_env = std::chrono::duration<double>(e1 - e0).count();
        ret.status = st;
        ret.sqp_iters = solver.sqp_iter;
        return st == SOLVED || (st == MAX_ITER_EXCEEDED && num_valid_guess_failed < 5);
    }
    bool runMPC(MPCReturn& ret, State& x0, const Input& u0) {  // :91-101
        double dummy[3] = {3, 3, 3};
        return runMPC_(ret, x0, u0, dummy, 0.);
    }
};

}  // namespace orc
