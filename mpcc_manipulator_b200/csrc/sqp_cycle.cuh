// One control cycle's SQP solve + the epilogue of runMPC_ for ONE instance, executed by a group of NL lanes
// (NL = 32: a warp of k_sqp_warp; NL = 128: the CTA of k_sqp_cta).  Shared by both kernel families.
#pragma once
#include "cycle_args.h"
#include "sqp_warp.cuh"

namespace mpcc {

// b: instance; sm: this group's shared memory (group_smem_doubles<NL>(N) doubles); lane in [0, NL)
template <int NL, bool SOC = false>
__device__ __forceinline__ void sqp_group_cycle(const CycleArgs& a, double* wws, size_t ws_per, double* sm, int b, int lane) {
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    const size_t B = (size_t)a.B, NS = B * a.S;
    const int HN = a.S * HZ;
    GroupSqp<NL, SOC> w{P, T, make_dyn(P, a.Ts), a.Ts, a.N, a.S, a.qp, Lanes<NL>{lane, nullptr}};
    w.carve(wws + (size_t)b * ws_per, sm);
    for (int e = lane; e < HN; e += NL) w.GUESS[e] = a.warm[(size_t)e * B + b];
    w.W.sync();
    double cur_u[NU], x0[NX];
    for (int i = 0; i < NU; i++) cur_u[i] = a.u0[b * NU + i];
    for (int i = 0; i < NX; i++) x0[i] = a.x0[b * NX + i];
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    SqpResult r = w.run(cur_u, a.rb + (size_t)b * a.S, NS, 1, nullptr);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    // epilogue of runMPC_ (mpc.cpp:140-188)
    WarmFlags fl = a.flags[b];
    if (r.status == SOLVED) { fl.valid = 1; fl.failed = 0; }
    else {
        // solveOCP hands back zero_guess (osqp_interface.cpp:422-428,587): every stage the current state, inputs zero
        for (int e = lane; e < HN; e += NL) { const int rr = e % HZ; w.GUESS[e] = (rr < NX) ? x0[rr] : 0.0; }
        fl.valid = 0; fl.failed++;
    }
    w.W.sync();
    const bool ok = r.status == SOLVED || (r.status == MAX_ITER_EXCEEDED && fl.failed < 5);
    double* h = a.horizon + (size_t)b * HN;
    double* hh = a.horizon_host ? a.horizon_host + (size_t)b * HN : nullptr;
    for (int e = lane; e < HN; e += NL) { const double v = w.GUESS[e]; a.warm[(size_t)e * B + b] = v; h[e] = v; if (hh) hh[e] = v; }
    if (lane < NU) a.u_out[b * NU + lane] = w.GUESS[NX + lane];
    if (lane == 0) {
        a.flags[b] = fl;
        a.status[b] = r.status; a.iters[b] = r.iters; a.ok[b] = ok ? 1 : 0; a.qp_iters[b] = r.qp_iters; a.qp_fail[b] = r.qp_fail;
        a.accept_mask[b] = (int32_t)r.accept_mask;
        a.hist[b] = (int32_t)(((unsigned)a.hist[b] << 8) | (unsigned)min(r.iters, 255));
        a.sqp_ns[4 * b] = t1 - t0; a.sqp_ns[4 * b + 1] = (long long)w.tm_set_qp; a.sqp_ns[4 * b + 2] = (long long)w.tm_solve_qp; a.sqp_ns[4 * b + 3] = (long long)w.tm_get_alpha;
    }
}

// SolverInterface::solveOCP probe for ONE instance: AoS guess / RobotData, optional iteration log
template <int NL, bool SOC = false>
__device__ __forceinline__ void solve_ocp_group(const CycleArgs& a, double* wws, size_t ws_per, double* sm, int b, int lane, double* guess, const double* rb,
                                                const double* cur_u_all, double* steps, double* alphas, int32_t* qp_ok, int max_log, int32_t* n_logged) {
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    const int HN = a.S * HZ;
    GroupSqp<NL, SOC> w{P, T, make_dyn(P, a.Ts), a.Ts, a.N, a.S, a.qp, Lanes<NL>{lane, nullptr}};
    w.carve(wws + (size_t)b * ws_per, sm);
    for (int e = lane; e < HN; e += NL) w.GUESS[e] = guess[(size_t)b * HN + e];
    w.W.sync();
    double cur_u[NU];
    for (int i = 0; i < NU; i++) cur_u[i] = cur_u_all[b * NU + i];
    SqpLogRef lg{steps ? steps + (size_t)b * max_log * HN : nullptr, alphas + (size_t)b * max_log, qp_ok + (size_t)b * max_log, max_log, 0};
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    SqpResult r = w.run(cur_u, rb + (size_t)b * a.S * RB_DOUBLES, 1, RB_DOUBLES, max_log > 0 ? &lg : nullptr);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    // solveOCP returns zero_guess unless SOLVED (osqp_interface.cpp:580-589)
    if (r.status != SOLVED) {
        for (int e = lane; e < HN; e += NL) { const int rr = e % HZ; w.GUESS[e] = (rr < NX) ? guess[(size_t)b * HN + rr] : 0.0; }
        w.W.sync();
    }
    for (int e = lane; e < HN; e += NL) guess[(size_t)b * HN + e] = w.GUESS[e];
    if (lane == 0) {
        a.status[b] = r.status; a.iters[b] = r.iters; a.qp_iters[b] = r.qp_iters; a.qp_fail[b] = r.qp_fail; n_logged[b] = lg.n;
        a.sqp_ns[4 * b] = t1 - t0; a.sqp_ns[4 * b + 1] = (long long)w.tm_set_qp; a.sqp_ns[4 * b + 2] = (long long)w.tm_solve_qp; a.sqp_ns[4 * b + 3] = (long long)w.tm_get_alpha;
    }
}

}  // namespace mpcc
