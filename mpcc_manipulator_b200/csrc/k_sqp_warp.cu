// Warp-per-instance SQP kernels (default): see sqp_warp.cuh.
#include "cycle_args.h"
#include "sqp_warp.cuh"
#include <cstdlib>

namespace mpcc {

// Launch order: instances that took many SQP iterations in one of the last four cycles go first, so that the long
// tail of the batch starts at time zero (longest-processing-time-first; affects scheduling only, never results).
// The first order[B] of them (>= 15 iterations lately: in practice the MAX_ITER "stragglers", which keep coming back
// for a while) are solved by a second, small launch of the same kernel on a high-priority stream whose CTAs ask for so
// much shared memory that each gets an SM of its own: a straggler's ~100 latency-bound evaluations and its first QP
// run ~1.5x faster without nine other warps on the SM, and it is the stragglers that bound the kernel in those cycles.
__global__ void k_order(const int32_t* __restrict__ hist, int32_t* __restrict__ order, int B, int32_t* hint) {
    __shared__ int cnt[16], base[16], n2;
    if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) n2 = 0;
    __syncthreads();
    auto key = [&](int b) {
        const unsigned h = (unsigned)hist[b];
        unsigned m = h & 255u;
        m = max(m, (h >> 8) & 255u); m = max(m, (h >> 16) & 255u); m = max(m, (h >> 24) & 255u);
        return 15 - (int)min(m, 15u);  // bucket 0 = longest
    };
    for (int b = threadIdx.x; b < B; b += blockDim.x) { atomicAdd(&cnt[key(b)], 1); if (((unsigned)hist[b] & 255u) >= 2u) atomicAdd(&n2, 1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (hint) { hint[0] = n2; hint[1] = cnt[0]; }  // pinned host memory: how transient the batch is (read, possibly a cycle late, by the launcher)
        order[B] = cnt[0]; int s = 0; for (int i = 0; i < 16; i++) { base[i] = s; s += cnt[i]; } }  // order[B]: instances with >= 15 iterations lately
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) order[atomicAdd(&base[key(b)], 1)] = b;
}

// SQP loop + epilogue, one WARP per instance (sqp_warp.cuh)
#ifndef MPCC_SQPW_WARPS
#define MPCC_SQPW_WARPS 2
#endif
#ifndef MPCC_SQPW_MINB
#define MPCC_SQPW_MINB 5
#endif
constexpr int SQPW_WARPS = MPCC_SQPW_WARPS;  // warps (instances) per CTA
#ifdef MPCC_SQPW_MAXNREG
#define SQPW_BOUNDS __maxnreg__(MPCC_SQPW_MAXNREG)
#else
#define SQPW_BOUNDS __launch_bounds__(SQPW_WARPS * 32, MPCC_SQPW_MINB)
#endif
constexpr int SQPW_EXCL_CTAS = 8;            // CTAs of the exclusive launch (one SM each)
// excl: 1 = exclusive launch (the first slots only), 0 = main launch (skips them), -1 = single launch (everything)
extern __shared__ __align__(16) double sqpw_smem[];
__device__ __forceinline__ void sqp_warp_cycle(const CycleArgs& a, double* wws, size_t ws_per, size_t sm_per, int excl) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SQPW_WARPS + wid;
    if (slot >= a.B) return;  // whole warps leave together
    if (excl >= 0) {
        const int n_excl = min(a.order[a.B], SQPW_EXCL_CTAS * SQPW_WARPS);
        if ((excl == 1) != (slot < n_excl)) return;
    }
    const int b = a.order[slot];
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    const size_t B = (size_t)a.B, NS = B * a.S;
    const int HN = a.S * HZ;
    WarpSqp w{P, T, make_dyn(P, a.Ts), a.Ts, a.N, a.S, a.qp, Warp{lane}};
    w.carve(wws + (size_t)b * ws_per, sqpw_smem + (size_t)wid * sm_per);
    for (int e = lane; e < HN; e += 32) w.GUESS[e] = a.warm[(size_t)e * B + b];
    __syncwarp();
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
        for (int e = lane; e < HN; e += 32) { const int rr = e % HZ; w.GUESS[e] = (rr < NX) ? x0[0] : 0.0; }
        __syncwarp();
        for (int k = lane; k < a.S; k += 32) for (int m = 0; m < NX; m++) w.GUESS[k * HZ + m] = x0[m];
        fl.valid = 0; fl.failed++;
    }
    __syncwarp();
    const bool ok = r.status == SOLVED || (r.status == MAX_ITER_EXCEEDED && fl.failed < 5);
    double* h = a.horizon + (size_t)b * HN;
    for (int e = lane; e < HN; e += 32) { const double v = w.GUESS[e]; a.warm[(size_t)e * B + b] = v; h[e] = v; }
    if (lane < NU) a.u_out[b * NU + lane] = w.GUESS[NX + lane];
    if (lane == 0) {
        a.flags[b] = fl;
        a.status[b] = r.status; a.iters[b] = r.iters; a.ok[b] = ok ? 1 : 0; a.qp_iters[b] = r.qp_iters; a.qp_fail[b] = r.qp_fail;
        a.accept_mask[b] = (int32_t)r.accept_mask;
        a.hist[b] = (int32_t)(((unsigned)a.hist[b] << 8) | (unsigned)min(r.iters, 255));
        a.sqp_ns[4 * b] = t1 - t0; a.sqp_ns[4 * b + 1] = (long long)w.tm_set_qp; a.sqp_ns[4 * b + 2] = (long long)w.tm_solve_qp; a.sqp_ns[4 * b + 3] = (long long)w.tm_get_alpha;
    }
}
// Two builds of the same code.  k_sqp_warp: 5 CTAs per SM (168 registers, 2.8 KB of spills) -- a steady-state batch (every
// instance one QP) wants resident warps: 5.0 ms against 5.45.  k_sqp_warp_r255: the full register file (255 registers, 0.8 KB
// of spills, 4 CTAs per SM) -- whatever runs long as a single warp (stragglers, multi-iteration instances of a start-up
// transient) is bound by its own dependent latency and the spills are on that path: 1.5 ms less in those cycles.  The
// exclusive launch always uses the second build, the main launch picks per cycle from the history (launch_sqp_warp).
__global__ void SQPW_BOUNDS k_sqp_warp(CycleArgs a, double* wws, size_t ws_per, size_t sm_per, int excl) { sqp_warp_cycle(a, wws, ws_per, sm_per, excl); }
__global__ void __launch_bounds__(SQPW_WARPS * 32, 1) k_sqp_warp_r255(CycleArgs a, double* wws, size_t ws_per, size_t sm_per, int excl) { sqp_warp_cycle(a, wws, ws_per, sm_per, excl); }
// solveOCP probe, one warp per instance: AoS guess / RobotData, optional iteration log
__global__ void SQPW_BOUNDS k_solve_ocp_warp(CycleArgs a, double* wws, size_t ws_per, size_t sm_per, double* guess, const double* rb,
                                                                    const double* cur_u_all, int n, double* steps, double* alphas, int32_t* qp_ok, int max_log, int32_t* n_logged) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * SQPW_WARPS + wid;
    if (b >= n) return;
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    const int HN = a.S * HZ;
    WarpSqp w{P, T, make_dyn(P, a.Ts), a.Ts, a.N, a.S, a.qp, Warp{lane}};
    w.carve(wws + (size_t)b * ws_per, sqpw_smem + (size_t)wid * sm_per);
    for (int e = lane; e < HN; e += 32) w.GUESS[e] = guess[(size_t)b * HN + e];
    __syncwarp();
    double cur_u[NU];
    for (int i = 0; i < NU; i++) cur_u[i] = cur_u_all[b * NU + i];
    SqpLogRef lg{steps ? steps + (size_t)b * max_log * HN : nullptr, alphas + (size_t)b * max_log, qp_ok + (size_t)b * max_log, max_log, 0};
    SqpResult r = w.run(cur_u, rb + (size_t)b * a.S * RB_DOUBLES, 1, RB_DOUBLES, max_log > 0 ? &lg : nullptr);
    for (int e = lane; e < HN; e += 32) guess[(size_t)b * HN + e] = w.GUESS[e];
    if (lane == 0) { a.status[b] = r.status; a.iters[b] = r.iters; n_logged[b] = lg.n; }
}


size_t sqp_warp_ws_doubles(int N) { return warp_ws_doubles(N); }
// MPCC_SQPW_SMEM_PAD (bytes, diagnostic): extra dynamic shared memory per CTA, to lower the number of resident CTAs per SM
// in residency experiments without rebuilding
static size_t sqp_warp_smem_pad() {
    static const size_t pad = [] { const char* e = getenv("MPCC_SQPW_SMEM_PAD"); return e ? (size_t)atol(e) : (size_t)0; }();
    return pad;
}
size_t sqp_warp_smem_bytes(int N) { return SQPW_WARPS * warp_smem_doubles(N) * sizeof(double) + sqp_warp_smem_pad(); }
// shared memory request of the exclusive launch: leaves no room for a CTA of the main launch on the same SM
static size_t sqp_warp_excl_smem_bytes(int N) {
    const size_t normal = sqp_warp_smem_bytes(N), want = (size_t)227 * 1024 - normal;  // 227 KB = per-SM limit on sm_100
    return want > normal ? want : normal;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the FUNCTION on the device, not to a handle: handles of different
// horizons live side by side (mpcc::MPC at N = 10 next to an N = 20 batch), so the limits are set once to the worst case
// (the launches themselves ask only for what their horizon needs, which is what decides the residency).
cudaError_t configure_sqp_warp(int) {
    const int main_limit = (int)sqp_warp_smem_bytes(MAX_N), excl_limit = 227 * 1024;
    cudaError_t e = cudaFuncSetAttribute(k_sqp_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, main_limit);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sqp_warp_r255, cudaFuncAttributeMaxDynamicSharedMemorySize, excl_limit);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_solve_ocp_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, main_limit);
}
void launch_sqp_warp(const CycleArgs& a, double* wws, cudaStream_t s, cudaStream_t aux, cudaEvent_t ev_pre, cudaEvent_t ev_order, cudaEvent_t ev_aux, int32_t* hint) {
    static const bool no_excl = getenv("MPCC_SQPW_NO_EXCL") != nullptr;  // diagnostic (as mpcc_cuda_config.reserved bit 0): everything in the main launch
    if (no_excl) aux = nullptr;
    const int grid = (a.B + SQPW_WARPS - 1) / SQPW_WARPS;
    // hint (pinned host memory, written by k_order, read here without synchronising -- it may be a cycle old, it only
    // selects a build): [0] instances with >= 2 SQP iterations in their last cycle, [1] instances with >= 15 lately
    // (the recent long runners themselves go to the exclusive launch, which always uses the 255-register build)
    const bool transient = hint && ((volatile int32_t*)hint)[0] * 32 >= a.B;
    auto main_launch = [&](int excl) {
        if (transient) k_sqp_warp_r255<<<grid, SQPW_WARPS * 32, sqp_warp_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), warp_smem_doubles(a.N), excl);
        else k_sqp_warp<<<grid, SQPW_WARPS * 32, sqp_warp_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), warp_smem_doubles(a.N), excl);
    };
    if (!aux) {
        k_order<<<1, 1024, 0, s>>>(a.hist, a.order, a.B, hint);
        main_launch(-1);
        return;
    }
    // The exclusive launch must get its SMs BEFORE the main launch fills the machine (a 180 KB CTA never fits next to
    // resident 44 KB CTAs): it follows k_order in the auxiliary stream, so it starts the moment the order exists, while
    // the main launch waits for the same order across streams.
    cudaEventRecord(ev_pre, s);                    // RobotData and the previous cycle's history are complete
    cudaStreamWaitEvent(aux, ev_pre, 0);
    k_order<<<1, 1024, 0, aux>>>(a.hist, a.order, a.B, hint);
    cudaEventRecord(ev_order, aux);
    const int gx = grid < SQPW_EXCL_CTAS ? grid : SQPW_EXCL_CTAS;
    k_sqp_warp_r255<<<gx, SQPW_WARPS * 32, sqp_warp_excl_smem_bytes(a.N), aux>>>(a, wws, warp_ws_doubles(a.N), warp_smem_doubles(a.N), 1);
    cudaEventRecord(ev_aux, aux);
    cudaStreamWaitEvent(s, ev_order, 0);
    main_launch(0);
    cudaStreamWaitEvent(s, ev_aux, 0);
}
void launch_solve_ocp_warp(const CycleArgs& a, double* wws, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                           int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s) {
    k_solve_ocp_warp<<<(n + SQPW_WARPS - 1) / SQPW_WARPS, SQPW_WARPS * 32, sqp_warp_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), warp_smem_doubles(a.N), guess, rb,
                                                                                                         cur_u, n, steps, alphas, qp_ok, max_log, n_logged);
}

}  // namespace mpcc
