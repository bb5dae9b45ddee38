// Kernel argument block of one control cycle and the launchers of the SQP kernels (each kernel family lives in
// its own translation unit so that the library builds in parallel).
#pragma once
#include "mpcc_types.h"
#include "dev_sqp.cuh"
#include <cuda_runtime.h>

namespace mpcc {

constexpr int MAX_SQP_ITER = 128;  // largest sqp.max_iter a handle accepts (filter capacity)
constexpr int FILT_DOUBLES = 2 * (MAX_SQP_ITER + 2);

struct CycleArgs {
    int B, N, S;
    double Ts;
    const Params* params; int params_per_instance;
    const TrackTable* tracks; const int32_t* track_id;
    double* x0; const double* u0; const double* obs;  // [B][9], [B][8], [B][4] (AoS)
    double* warm;      // [S*17][B]  (SoA) warm start == SQP iterate
    double* step;      // [S*17][B]
    double* trial;     // [S*17][B]
    double* filt;      // [FILT_DOUBLES][B]
    double* ws;        // [S*STAGE_WS][B]
    WarmFlags* flags;  // [B]
    double* qs;        // [7][B*S]
    double* rb;        // [150][B*S]
    double* u_out;     // [B][8]
    double* horizon;   // [B][S][17]
    double* horizon_host;  // optional second destination of the horizon: the caller's PINNED host buffer, written by the SQP kernel as each instance finishes
                           // (the 11.7 MB of mpc_horizon then cross PCIe under the kernel instead of in a copy behind it); nullptr: none
    int32_t* status; int32_t* iters; int32_t* ok; int32_t* qp_iters; int32_t* qp_fail; int32_t* accept_mask;
    long long* sqp_ns;  // per-instance duration of the SQP loop (ComputeTime::total analogue)
    int32_t* hist;      // per-instance history of SQP iteration counts (4 cycles, one byte each): scheduling hint only
    int32_t* order;     // launch order of the instances (longest expected first)
    QpOptions qp;
};


// one warp per instance (k_sqp_warp.cu)
size_t sqp_warp_ws_doubles(int N);    // global workspace per instance
size_t sqp_warp_smem_bytes(int N);    // dynamic shared memory per CTA
cudaError_t configure_sqp_warp(int N);
// aux != nullptr: the instances with many recent SQP iterations go to an exclusive-SM launch on `aux` (see k_sqp_warp.cu)
// hint: two pinned host ints written by k_order (how transient the batch is), or nullptr
void launch_sqp_warp(const CycleArgs& a, double* wws, cudaStream_t s, cudaStream_t aux, cudaEvent_t ev_pre, cudaEvent_t ev_order, cudaEvent_t ev_aux, int32_t* hint);
void launch_solve_ocp_warp(const CycleArgs& a, double* wws, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                           int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s);

// track ingestion on the device, one thread per track (k_track_fit.cu); scratch: track_fit_scratch_doubles(n) * n_tracks doubles
void launch_fit_tracks(int n_tracks, int n, const double* X, const double* Y, const double* Z, const double* R, double* scratch, TrackTable* out, cudaStream_t s);
// one CTA (128 lanes) per instance: the latency path (k_sqp_cta.cu)
size_t sqp_cta_smem_bytes(int N);
cudaError_t configure_sqp_cta();
void launch_sqp_cta(const CycleArgs& a, double* wws, cudaStream_t s);
void launch_solve_ocp_cta(const CycleArgs& a, double* wws, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                          int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s);

// the same two kernel families with the second-order correction compiled in (k_sqp_soc.cu); cta: one CTA per instance
cudaError_t configure_sqp_soc();
void launch_sqp_soc(const CycleArgs& a, double* wws, bool cta, cudaStream_t s);
void launch_solve_ocp_soc(const CycleArgs& a, double* wws, bool cta, double* guess, const double* rb, const double* cur_u, int n, double* steps, double* alphas,
                          int32_t* qp_ok, int max_log, int32_t* n_logged, cudaStream_t s);

}  // namespace mpcc
