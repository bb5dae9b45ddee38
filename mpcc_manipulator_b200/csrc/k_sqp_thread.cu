// One-thread-per-instance SQP kernels (sqp_kernel = 1): the direct device compilation of dev_sqp.cuh / dev_qp.cuh.
#include "cycle_args.h"

namespace mpcc {

// SQP loop + epilogue (osqp_interface.cpp:398-590, mpc.cpp:140-188): one thread per instance
__global__ void k_sqp_thread(CycleArgs a) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    const size_t B = (size_t)a.B, NS = B * a.S;
    WsRef guess{a.warm + b, B}, step{a.step + b, B}, trial{a.trial + b, B}, filt{a.filt + b, B}, ws{a.ws + b, B};
    double cur_u[NU], x0[NX];
    for (int i = 0; i < NU; i++) cur_u[i] = a.u0[b * NU + i];
    for (int i = 0; i < NX; i++) x0[i] = a.x0[b * NX + i];
    SqpResult r = sqp_solve(P, T, a.Ts, a.N, guess, step, trial, filt, cur_u, a.rb + (size_t)b * a.S, NS, 1, ws, a.qp, nullptr);
    WarmFlags fl = a.flags[b];
    bool ok = cycle_epilogue(a.N, r, x0, guess, fl);
    a.flags[b] = fl;
    a.status[b] = r.status; a.iters[b] = r.iters; a.ok[b] = ok ? 1 : 0; a.qp_iters[b] = r.qp_iters; a.qp_fail[b] = r.qp_fail; a.accept_mask[b] = (int32_t)r.accept_mask;
    for (int j = 0; j < NU; j++) a.u_out[b * NU + j] = guess[NX + j];
    double* h = a.horizon + (size_t)b * a.S * HZ;
    for (int e = 0; e < a.S * HZ; e++) h[e] = guess[e];
}

// solveOCP on given warm starts and RobotData (both AoS), logging the SQP iterations
__global__ void k_solve_ocp(CycleArgs a, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                            int32_t* qp_ok, int max_log, int32_t* n_logged) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    const size_t B = (size_t)a.B;
    const int HN = a.S * HZ;
    WsRef g{guess + (size_t)b * HN, 1}, step{a.step + b, B}, trial{a.trial + b, B}, filt{a.filt + b, B}, ws{a.ws + b, B};
    SqpLogRef lg{steps ? steps + (size_t)b * max_log * HN : nullptr, alphas + (size_t)b * max_log, qp_ok + (size_t)b * max_log, max_log, 0};
    SqpResult r = sqp_solve(P, T, a.Ts, a.N, g, step, trial, filt, cur_u + b * NU, rb + (size_t)b * a.S * RB_DOUBLES, 1, RB_DOUBLES, ws, a.qp,
                            max_log > 0 ? &lg : nullptr);
    a.status[b] = r.status; a.iters[b] = r.iters;
    n_logged[b] = lg.n;
}

void launch_sqp_thread(const CycleArgs& a, cudaStream_t s) { k_sqp_thread<<<(a.B + 31) / 32, 32, 0, s>>>(a); }
void launch_solve_ocp_thread(const CycleArgs& a, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                             int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s) {
    k_solve_ocp<<<(n + 31) / 32, 32, 0, s>>>(a, guess, rb, cur_u, n, steps, alphas, qp_ok, max_log, n_logged);
}

}  // namespace mpcc
