// CTA-per-instance SQP kernels: the LATENCY path (BASELINE configs[4]: a batch that does not fill the machine).
// Same code as the warp kernel (sqp_warp.cuh is a template on the lane count): 128 lanes work on one instance, so the parts of
// a cycle that are parallel over constraints / stages / entries (linearisation with lane = stage, the five per-constraint
// passes and two gradients of an interior-point iteration, the tile copies) run 4 wide, while the inherently sequential
// Riccati recursion stays on warp 0.  One CTA per instance; the CTA's shared memory (~100 KB at N = 40) lets at most two
// share an SM, and a batch of up to 2 x 148 instances is resident at once.
#include "cycle_args.h"
#include "sqp_cycle.cuh"

namespace mpcc {

constexpr int CTA_NL = 128;
extern __shared__ __align__(16) double sqpc_smem[];

__global__ void __launch_bounds__(CTA_NL, 1) k_sqp_cta(CycleArgs a, double* wws, size_t ws_per) {
    sqp_group_cycle<CTA_NL>(a, wws, ws_per, sqpc_smem, (int)blockIdx.x, (int)threadIdx.x);
}
__global__ void __launch_bounds__(CTA_NL, 1) k_solve_ocp_cta(CycleArgs a, double* wws, size_t ws_per, double* guess, const double* rb, const double* cur_u_all,
                                                             double* steps, double* alphas, int32_t* qp_ok, int max_log, int32_t* n_logged) {
    solve_ocp_group<CTA_NL>(a, wws, ws_per, sqpc_smem, (int)blockIdx.x, (int)threadIdx.x, guess, rb, cur_u_all, steps, alphas, qp_ok, max_log, n_logged);
}

size_t sqp_cta_smem_bytes(int N) { return group_smem_doubles<CTA_NL>(N) * sizeof(double); }
cudaError_t configure_sqp_cta() {
    // the attribute belongs to the function, not to a handle: worst case once (see configure_sqp_warp)
    cudaError_t e = cudaFuncSetAttribute(k_sqp_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sqp_cta_smem_bytes(MAX_N));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_solve_ocp_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sqp_cta_smem_bytes(MAX_N));
}
void launch_sqp_cta(const CycleArgs& a, double* wws, cudaStream_t s) {
    k_sqp_cta<<<a.B, CTA_NL, sqp_cta_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N));
}
void launch_solve_ocp_cta(const CycleArgs& a, double* wws, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                          int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s) {
    k_solve_ocp_cta<<<n, CTA_NL, sqp_cta_smem_bytes(a.N), s>>>(a, wws, warp_ws_doubles(a.N), guess, rb, cur_u, steps, alphas, qp_ok, max_log, n_logged);
}

}  // namespace mpcc
