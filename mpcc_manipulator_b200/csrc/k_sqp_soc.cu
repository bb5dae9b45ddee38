// SQP kernels with the reference's optional second-order correction compiled in (sqp.json "do_SOC", default false;
// osqp_interface.cpp:506-533,658-681): GroupSqp<NL, true> of sqp_warp.cuh.  A handle uses them as soon as one of its parameter sets
// switches the correction on; an instance whose own set has it off runs the plain loop inside them.  They live in their own translation
// unit so that the default kernels (k_sqp_warp.cu, k_sqp_cta.cu) stay byte for byte the code they were tuned as.  One launch per cycle
// (no exclusive-SM launch): with the correction every SQP iteration solves two QPs, the path is not the throughput configuration.
#include "cycle_args.h"
#include "sqp_cycle.cuh"

namespace mpcc {

constexpr int SOC_WARPS = 2;    // warps (instances) per CTA of the warp-per-instance kernel
constexpr int SOC_CTA_NL = 128;
extern __shared__ __align__(16) double sqps_smem[];

__global__ void __launch_bounds__(SOC_WARPS * 32, 1) k_sqp_warp_soc(CycleArgs a, double* wws, size_t ws_per, size_t sm_per) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SOC_WARPS + wid;
    if (slot >= a.B) return;  // whole warps leave together
    sqp_group_cycle<32, true>(a, wws, ws_per, sqps_smem + (size_t)wid * sm_per, slot, lane);
}
__global__ void __launch_bounds__(SOC_WARPS * 32, 1) k_solve_ocp_warp_soc(CycleArgs a, double* wws, size_t ws_per, size_t sm_per, double* guess, const double* rb,
                                                                         const double* cur_u_all, int n, double* steps, double* alphas, int32_t* qp_ok, int max_log, int32_t* n_logged) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * SOC_WARPS + wid;
    if (b >= n) return;
    solve_ocp_group<32, true>(a, wws, ws_per, sqps_smem + (size_t)wid * sm_per, b, lane, guess, rb, cur_u_all, steps, alphas, qp_ok, max_log, n_logged);
}
__global__ void __launch_bounds__(SOC_CTA_NL, 1) k_sqp_cta_soc(CycleArgs a, double* wws, size_t ws_per) {
    sqp_group_cycle<SOC_CTA_NL, true>(a, wws, ws_per, sqps_smem, (int)blockIdx.x, (int)threadIdx.x);
}
__global__ void __launch_bounds__(SOC_CTA_NL, 1) k_solve_ocp_cta_soc(CycleArgs a, double* wws, size_t ws_per, double* guess, const double* rb, const double* cur_u_all,
                                                                    double* steps, double* alphas, int32_t* qp_ok, int max_log, int32_t* n_logged) {
    solve_ocp_group<SOC_CTA_NL, true>(a, wws, ws_per, sqps_smem, (int)blockIdx.x, (int)threadIdx.x, guess, rb, cur_u_all, steps, alphas, qp_ok, max_log, n_logged);
}

static size_t soc_warp_smem_bytes(int N) { return SOC_WARPS * warp_smem_doubles(N) * sizeof(double); }
static size_t soc_cta_smem_bytes(int N) { return group_smem_doubles<SOC_CTA_NL>(N) * sizeof(double); }
// the attribute belongs to the function, not to a handle: worst case once (see configure_sqp_warp)
cudaError_t configure_sqp_soc() {
    cudaError_t e = cudaFuncSetAttribute(k_sqp_warp_soc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)soc_warp_smem_bytes(MAX_N));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_solve_ocp_warp_soc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)soc_warp_smem_bytes(MAX_N));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sqp_cta_soc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)soc_cta_smem_bytes(MAX_N));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_solve_ocp_cta_soc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)soc_cta_smem_bytes(MAX_N));
}
void launch_sqp_soc(const CycleArgs& a, double* wws, bool cta, cudaStream_t s) {
    if (cta) k_sqp_cta_soc<<<a.B, SOC_CTA_NL, soc_cta_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N));
    else k_sqp_warp_soc<<<(a.B + SOC_WARPS - 1) / SOC_WARPS, SOC_WARPS * 32, soc_warp_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), warp_smem_doubles(a.N));
}
void launch_solve_ocp_soc(const CycleArgs& a, double* wws, bool cta, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                          int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s) {
    if (cta) k_solve_ocp_cta_soc<<<n, SOC_CTA_NL, soc_cta_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), guess, rb, cur_u, steps, alphas, qp_ok, max_log, n_logged);
    else k_solve_ocp_warp_soc<<<(n + SOC_WARPS - 1) / SOC_WARPS, SOC_WARPS * 32, soc_warp_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), warp_smem_doubles(a.N), guess, rb,
                                                                                                              cur_u, n, steps, alphas, qp_ok, max_log, n_logged);
}

}  // namespace mpcc
