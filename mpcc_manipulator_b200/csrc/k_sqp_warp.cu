// Warp-per-instance SQP kernels (default): see sqp_warp.cuh.
#include "cycle_args.h"
#include "sqp_cycle.cuh"
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
    // warp index and instance number go through a shuffle: provably warp-uniform for the compiler (uniform registers for everything derived
    // from them -- the workspace pointers --, no convergence barriers around warp-uniform branches)
    const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SQPW_WARPS + wid;
    if (slot >= a.B) return;  // whole warps leave together
    if (excl >= 0) {
        const int n_excl = min(a.order[a.B], SQPW_EXCL_CTAS * SQPW_WARPS);
        if ((excl == 1) != (slot < n_excl)) return;
    }
    sqp_group_cycle<32>(a, wws, ws_per, sqpw_smem + (size_t)wid * sm_per, __shfl_sync(0xffffffffu, a.order[slot], 0), lane);
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
    solve_ocp_group<32>(a, wws, ws_per, sqpw_smem + (size_t)wid * sm_per, b, lane, guess, rb, cur_u_all, steps, alphas, qp_ok, max_log, n_logged);
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
