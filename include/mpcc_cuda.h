/* C ABI of the B200-native batched MPCC SQP path (libmpcc_b200.so).
 *
 * This is the drop-in boundary for the hot path of JunHeonYoon/MPCC_manipulator: one call to
 * mpcc_cuda_run_cycle() performs, for every instance of the batch, what one call to
 *     bool mpcc::MPC::runMPC_(MPCReturn&, State& x0, Input& u0, const Eigen::Vector3d& obs, const double& r)
 * (reference cpp/src/MPC/mpc.cpp:104-190, declared cpp/include/MPC/mpc.h:58-74) does for one, i.e. the
 * calls MPC makes through the SolverInterface seam (cpp/include/Interfaces/solver_interface.h:44-54):
 * setCurrentInput, setInitialGuess, setEnvData, solveOCP.  Plain pointers and sizes only; every
 * function returns 0 on success and a non-zero mpcc_cuda_error otherwise (no exceptions cross the
 * boundary); mpcc_cuda_last_error() returns a description of the last failure on this thread.
 *
 * Layouts (all IEEE double unless noted, row-major, instance-major):
 *   state  x   [9]  = q1..q7, s, vs            (reference types.h:33-71)
 *   input  u   [8]  = dq1..dq7, dVs            (reference types.h:73-118)
 *   horizon    [N+1][17] = per stage [x(9), u(8)]   (std::vector<OptVariables>, osqp_interface.h:48-62)
 *   obstacle   [4]  = x, y, z, radius          (runMPC_ arguments)
 *   params     [MPCC_PARAMS_DOUBLES]           (the six Params/ *.json files flattened, see below)
 *   track      [MPCC_TRACK_DOUBLES]            (fitted ArcLengthSpline table from mpcc_fit_track)
 *   robot data [150] = q7|p3|R9|Jv21|Jw21|manip|dmanip7|sel|dsel7|obs_r|env9|denv63  (robot_data.h:11-94)
 */
#ifndef MPCC_CUDA_H
#define MPCC_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCC_NX 9
#define MPCC_NU 8
#define MPCC_HZ 17
#define MPCC_RB_DOUBLES 150
#define MPCC_STAGE_LIN_DOUBLES 212
/* params record: model7 | cost12 | lx9 ux9 lu8 uu8 ldd7 udd7 | Tx9 Tu8 | sqp9 | r_ddq_solver
 *   model7 = max_dist_proj, desired_ee_velocity, s_trust_region, deaccelerate_ratio, tol_sing, tol_selcol, tol_envcol
 *   cost12 = qC, qCNmult, qL, qVs, qOri, qSing, rdq, rddq, rdVs, qC_reduction_ratio, qL_increase_ratio, qOri_reduction_ratio
 *   sqp9   = eps_prim, eps_dual, max_iter, line_search_max_iter, do_SOC, use_BFGS, line_search_tau, _eta, _rho
 *   r_ddq_solver = the cost FILE's rddq (what OsqpInterface::cost_param_ holds, osqp_interface.cpp:28,57) */
#define MPCC_PARAMS_DOUBLES 94
#define MPCC_TRACK_DOUBLES 2704

typedef enum {
    MPCC_OK = 0,
    MPCC_ERR_INVALID = 1,   /* bad argument / configuration */
    MPCC_ERR_CUDA = 2,      /* CUDA runtime failure */
    MPCC_ERR_STATE = 3,     /* call order: params / nn / track missing */
    MPCC_ERR_IO = 4         /* file could not be read or parsed */
} mpcc_cuda_error;

/* numeric order of mpcc::Status (reference solver_interface.h:28-42) */
typedef enum {
    MPCC_SOLVED = 0, MPCC_MAX_ITER_EXCEEDED, MPCC_QP_DualInfeasibleInaccurate, MPCC_QP_PrimalInfeasibleInaccurate,
    MPCC_QP_SolvedInaccurate, MPCC_QP_MaxIterReached, MPCC_QP_PrimalInfeasible, MPCC_QP_DualInfeasible, MPCC_Sigint,
    MPCC_INVALID_SETTINGS, MPCC_NAN_HESSIAN, MPCC_NON_PD_HESSIAN
} mpcc_status;

typedef struct {
    int32_t batch;        /* instances on this GPU */
    int32_t horizon;      /* N (reference: compile-time 10, config.h:36); 2 <= N <= 64 */
    double Ts;            /* control period (config.json "Ts") */
    int32_t device;       /* CUDA device ordinal */
    int32_t qp_max_iter;  /* interior-point iteration cap of the structured QP solver (0 -> 60) */
    double qp_eps;        /* its residual tolerance (0 -> 1e-9) */
    int32_t sqp_kernel;   /* must be 0 (one SQP kernel ships: a warp per instance) */
    int32_t reserved;     /* diagnostics: bit 0 = no exclusive-SM launch for recent long runners (scheduling only, same results);
                           * bit 1 = always the warp-per-instance SQP kernel, bit 2 = always the CTA-per-instance one
                           * (default: CTA per instance when batch <= 2 x SMs -- the latency path --, else warp per instance);
                           * bit 3 = collision networks entirely in fp64 (k_mlp: mma.sync.m8n8k4.f64) instead of the default kernel, which runs their
                           * three 256 x 256 layers as exact int8 digit products on tcgen05 (k_mlp_oz; results within 1e-13 of each other);
                           * bit 4 = print the per-phase cycle counts of that kernel's CTA 0 to stderr after mpcc_cuda_eval_robot_data;
                           * bit 8 = mpcc_cuda_run_cycle always copies mpc_horizon behind the kernel (default: a PINNED `horizon` buffer is written
                           * by the SQP kernel directly, under its own run time; a pageable one is copied) */
} mpcc_cuda_config;

typedef struct mpcc_cuda_handle mpcc_cuda_handle;

const char* mpcc_cuda_last_error(void);

int mpcc_cuda_create(const mpcc_cuda_config* cfg, mpcc_cuda_handle** out);
int mpcc_cuda_destroy(mpcc_cuda_handle* h);

/* Networks (SelCollNNmodel / EnvCollNNmodel::setNeuralNetwork, osqp_interface.cpp:35-43): layer-after-layer,
 * row per output neuron.  self: 256x21, 64x256, 1x64; env: 256x30, 256x256 (x3), 9x256; biases likewise. */
int mpcc_cuda_upload_nn(mpcc_cuda_handle* h, const double* self_w, const double* self_b, const double* env_w, const double* env_b);
/* either this repo's packed file (*.f64) or the reference's directory of weight_k.txt / bias_k.txt */
int mpcc_cuda_load_nn(mpcc_cuda_handle* h, const char* self_path, const char* env_path);

/* n_sets == 1: one parameter set for the whole batch; n_sets == batch: one per instance.
 * sqp.do_SOC (osqp_interface.cpp:506-533,658-681) is honoured per instance: a handle with at least one such set runs the SQP kernels that have the
 * second-order correction compiled in.  sqp.use_BFGS != 0 is rejected (MPCC_ERR_INVALID); sqp.max_iter must be in [1, 128]. */
int mpcc_cuda_set_params(mpcc_cuda_handle* h, const double* params, int32_t n_sets);
/* the reference's PathToJson (types.h:127-134); overrides: n_over (key, value) pairs with keys written
 * "file.key", e.g. "cost.qC" (the reference's ParamValue maps, types.h:143-150) */
int mpcc_load_params_json(const char* model_path, const char* cost_path, const char* bounds_path, const char* normalization_path,
                          const char* sqp_path, const char* const* over_keys, const double* over_vals, int32_t n_over, double* params_out);

/* Track ingestion (Track::Track + ArcLengthSpline::gen6DSpline): waypoints -> table. R: n x 9 rotation matrices */
int mpcc_fit_track(int32_t n, const double* X, const double* Y, const double* Z, const double* R, double* table_out);
/* n_tracks fits at once on n_threads host threads (0: all cores): X, Y, Z [n_tracks][n], R [n_tracks][n][9] -> tables
 * [n_tracks][MPCC_TRACK_DOUBLES].  For heterogeneous batches (one track per instance). */
int mpcc_fit_tracks(int32_t n_tracks, int32_t n, const double* X, const double* Y, const double* Z, const double* R, double* tables_out, int32_t n_threads);
int mpcc_load_track_json(const char* track_path, const double* init_position3 /* nullable */, double* table_out);
/* tables: n_tracks x MPCC_TRACK_DOUBLES; track_of_instance: batch indices or NULL (all instances use track 0).
 * Invalidates every warm start (valid_initial_guess_ = false; the failure counter keeps its value: MPC::setTrack, mpc.cpp:192-197). */
int mpcc_cuda_set_tracks(mpcc_cuda_handle* h, const double* tables, int32_t n_tracks, const int32_t* track_of_instance);

/* Track ingestion on the DEVICE (ArcLengthSpline::fitSpline, arc_length_spline.cpp:213-253, one thread per track): host
 * waypoints X, Y, Z [n_tracks][n], R [n_tracks][n][9] in; the fitted tables become the handle's tracks (equivalent to
 * mpcc_fit_tracks + mpcc_cuda_set_tracks).  mpcc_cuda_get_tracks copies the first n_tracks installed tables back to the host. */
int mpcc_cuda_fit_tracks(mpcc_cuda_handle* h, int32_t n_tracks, int32_t n, const double* X, const double* Y, const double* Z, const double* R,
                         const int32_t* track_of_instance);
int mpcc_cuda_get_tracks(mpcc_cuda_handle* h, double* tables_out, int32_t n_tracks);
/* Table from the 100 knots of an already fitted ArcLengthSpline (s, X, Y, Z [100], R [100][9]): no fit / resample pass. */
int mpcc_track_from_knots(const double* s, const double* X, const double* Y, const double* Z, const double* R, double* table_out);

/* forget all warm starts (valid_initial_guess_ = false, num_valid_guess_failed_ = 0) */
int mpcc_cuda_reset(mpcc_cuda_handle* h);

/* One control cycle for the whole batch, HOST buffers: copies x0/u0/obs in, runs the cycle, copies the
 * results out and returns after the stream is idle.  x0 is updated in place (s and vs), as runMPC does
 * (mpc.cpp:108,115).  obs may be NULL: the reference's dummy obstacle (3,3,3), radius 0 (mpc.cpp:97-100).
 * Outputs (any may be NULL): u_out [B][8] (MPCReturn::u0), horizon [B][N+1][17] (MPCReturn::mpc_horizon),
 * status [B] (mpcc_status), sqp_iters [B], ok [B] (runMPC's bool). */
int mpcc_cuda_run_cycle(mpcc_cuda_handle* h, double* x0, const double* u0, const double* obs,
                        double* u_out, double* horizon, int32_t* status, int32_t* sqp_iters, int32_t* ok);

/* Same cycle on DEVICE buffers, enqueued on the handle's stream without synchronising; results stay on the
 * device until mpcc_cuda_read_results().  d_obs may be NULL. */
int mpcc_cuda_run_cycle_device(mpcc_cuda_handle* h, double* d_x0, const double* d_u0, const double* d_obs);
int mpcc_cuda_read_results(mpcc_cuda_handle* h, double* u_out, double* horizon, int32_t* status, int32_t* sqp_iters, int32_t* ok);
/* device pointers of the result arrays (same shapes as above) and the handle's stream (a cudaStream_t) */
int mpcc_cuda_result_pointers(mpcc_cuda_handle* h, double** d_u_out, double** d_horizon, int32_t** d_status, int32_t** d_iters, int32_t** d_ok);
void* mpcc_cuda_stream(mpcc_cuda_handle* h);
int mpcc_cuda_synchronize(mpcc_cuda_handle* h);

/* Warm-start state of every instance (MPC::initial_guess_, valid_initial_guess_, num_valid_guess_failed_). */
int mpcc_cuda_get_warm_state(mpcc_cuda_handle* h, double* horizon, int32_t* valid, int32_t* failed);
int mpcc_cuda_set_warm_state(mpcc_cuda_handle* h, const double* horizon, const int32_t* valid, const int32_t* failed);

/* Closed-loop plant step for the whole batch: Integrator::simTimeStep (integrator.cpp:55-68). Host buffers. */
int mpcc_cuda_sim_time_step(mpcc_cuda_handle* h, const double* x, const double* u, double ts, double* x_next);

/* Same plant step on DEVICE buffers ([B][9], [B][8] -> [B][9]), enqueued on the handle's stream. */
int mpcc_cuda_sim_time_step_device(mpcc_cuda_handle* h, const double* d_x, const double* d_u, double ts, double* d_x_next);

/* Instrumentation (the reference's ComputeTime, osqp_interface.h:71-79, per kernel instead of per phase):
 * with profiling on, every cycle records CUDA events between its kernels on the handle's stream;
 * ms4 = durations of [prologue, kinematics, networks, SQP] of the last cycle (waits for that cycle). */
int mpcc_cuda_set_profiling(mpcc_cuda_handle* h, int32_t on);
int mpcc_cuda_get_kernel_times(mpcc_cuda_handle* h, double* ms4);
/* measured FP64 throughput of a device (the roofline denominator of this path): best of 5 bursts of DFMA chains and of
 * mma.sync.m8n8k4.f64 chains (one shared pipe; the larger figure is returned) */
int mpcc_cuda_fp64_peak(int32_t device, double* tflops);

/* ---- per-function evaluators (each runs the same device code the cycle uses; n <= batch*(N+1)) ---- */
/* RobotData::update + updateEnv for n joint vectors: q [n][7], obs [n][4] (NULL -> dummy) -> rb [n][150] */
int mpcc_cuda_eval_robot_data(mpcc_cuda_handle* h, const double* q, const double* obs, int32_t n, double* rb_out);
/* Stage linearisation of instance 0's parameter set / track 0 for n stages:
 * x [n][9], u [n][8], u_prev [n][7], u_next [n][7], x_next [n][9], rb [n][150], k [n] -> lin [n][212]
 * (Cost::getCost, Constraints::getConstraints, Bounds, dynamics defect in the normalised QP blocks) */
int mpcc_cuda_eval_stage(mpcc_cuda_handle* h, const double* x, const double* u, const double* u_prev, const double* u_next,
                         const double* x_next, const double* rb, const int32_t* k, int32_t n, double* lin_out);
/* Track evaluation on track 0: s [n] -> out [n][21] = pos3, dpos3, ddpos3, R9, dR3 */
int mpcc_cuda_eval_track(mpcc_cuda_handle* h, const double* s, int32_t n, double* out);
/* SolverInterface::solveOCP on given warm starts with given (frozen) RobotData, for the first n instances:
 * guess [n][N+1][17] in/out, rb [n][N+1][150], cur_u [n][8]; status/iters [n]; optional log of the first
 * max_log SQP iterations: steps [n][max_log][N+1][17] (normalised QP steps), alphas [n][max_log], n_logged [n] */
int mpcc_cuda_solve_ocp(mpcc_cuda_handle* h, double* guess, const double* rb, const double* cur_u, int32_t n,
                        int32_t* status, int32_t* iters, double* steps, double* alphas, int32_t max_log, int32_t* n_logged);

/* Per-instance ComputeTime of the last cycle's solveOCP (osqp_interface.h:71-79) in seconds, device global timer:
 * seconds [B][4] = total, set_qp, solve_qp, get_alpha.  Filled by the warp-per-instance kernel. */
int mpcc_cuda_read_compute_time(mpcc_cuda_handle* h, double* seconds);

/* Per-instance QP counters of the last cycle: interior-point iterations summed over the cycle's QPs, failed QPs. */
int mpcc_cuda_read_qp_counters(mpcc_cuda_handle* h, int32_t* qp_iters, int32_t* qp_fail);

/* Line-search decisions of the last cycle, per instance: bit i = the filter accepted the first trial of SQP
 * iteration i (filterLineSearch, osqp_interface.cpp:759-808; i < 32).  Diagnostic used by the parity tests to
 * replay the oracle along the same branch when a decision hinges on solver noise. */
int mpcc_cuda_read_decisions(mpcc_cuda_handle* h, int32_t* accept_mask);

/* ---- multi-GPU (one process and one handle per GPU; the caller shards the batch: instances are independent) ----
 * The path's only exchange is the gather of the per-instance results.  The library owns a NCCL communicator
 * (bound at run time, libnccl.so.2): rank 0 calls mpcc_cuda_comm_unique_id() and hands the 128 bytes to every rank by
 * whatever means the application has (MPI, a file, torch.distributed); every rank then calls mpcc_cuda_comm_init().
 * world == 1 needs no NCCL.  mpcc_cuda_gather_results() ENQUEUES, behind the current cycle and without blocking the host or
 * the next cycle, an all-gather of [u0 (8 doubles) | status, iters (2 x int32)] of every instance of every rank on a side
 * stream (double-buffered); mpcc_cuda_read_gathered() waits for the most recent one and unpacks it into host arrays
 * u_all [world*B][8], status_all / iters_all [world*B] (rank-major; any may be NULL).  mpcc_cuda_gathered_pointer(): the
 * packed device buffer [world*B][9] of the most recent gather and the stream it is ordered on. */
int mpcc_cuda_comm_unique_id(uint8_t* id128);
int mpcc_cuda_comm_init(mpcc_cuda_handle* h, const uint8_t* id128, int32_t rank, int32_t world);
int mpcc_cuda_gather_results(mpcc_cuda_handle* h);
int mpcc_cuda_read_gathered(mpcc_cuda_handle* h, double* u_all, int32_t* status_all, int32_t* iters_all);
int mpcc_cuda_gathered_pointer(mpcc_cuda_handle* h, double** d_packed, void** stream);

/* counters of the last cycle: [0] kernels launched, [1] total SQP iterations, [2] total QP (IPM) iterations,
 * [3] QP failures, [4] instances SOLVED, [5] instances with ok == 1 */
int mpcc_cuda_get_stats(mpcc_cuda_handle* h, int64_t* stats6);
/* kernels (and collectives) enqueued by the last cycle; host-side counter, does not synchronise */
int64_t mpcc_cuda_launch_count(mpcc_cuda_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MPCC_CUDA_H */
