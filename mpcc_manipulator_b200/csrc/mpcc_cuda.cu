// libmpcc_b200.so: CUDA kernels of the batched MPCC control cycle and the C ABI (include/mpcc_cuda.h).
// sm_100a only.  There is no CPU fallback: every entry point fails with MPCC_ERR_CUDA if no device works.
#include "../../include/mpcc_cuda.h"
#include "mpcc_types.h"
#include "dev_panda.cuh"
#include "dev_track.cuh"
#include "dev_track_fit.cuh"
#include "dev_stage.cuh"
#include "dev_qp.cuh"
#include "dev_sqp.cuh"
#include "cycle_args.h"
#include "mlp_kernel.cuh"
#include "mlp_oz_kernel.cuh"
#include "host/params_io.h"
#include "host/track_fit.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <thread>
#include <dlfcn.h>

// ---- NCCL, bound at run time (dlopen) so that the library has no link-time dependency on it: only the multi-GPU entry
// points need it.  Minimal declarations (stable across NCCL 2.x): opaque communicator, 128-byte unique id, ncclDouble = 8.
typedef struct ncclComm* mpcc_ncclComm_t;
typedef struct { char internal[128]; } mpcc_ncclUniqueId;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(mpcc_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(mpcc_ncclComm_t*, int, mpcc_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(mpcc_ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, mpcc_ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) { err = "NCCL symbols missing"; lib = nullptr; return false; }
        return true;
    }
};
static NcclApi g_nccl;

using namespace mpcc;

static_assert(PARAMS_DOUBLES == MPCC_PARAMS_DOUBLES, "params layout");
static_assert(TRACK_DOUBLES == MPCC_TRACK_DOUBLES, "track layout");
static_assert(sizeof(StageLin) == MPCC_STAGE_LIN_DOUBLES * sizeof(double), "stage lin layout");
static_assert(RB_DOUBLES == MPCC_RB_DOUBLES, "robot data layout");


// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// prologue of runMPC_ (mpc.cpp:104-124): one thread per instance
__global__ void k_prologue(CycleArgs a) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const Params& P = a.params[a.params_per_instance ? b : 0];
    const TrackTable& T = a.tracks[a.track_id[b]];
    double x0[NX], u0[NU];
    for (int i = 0; i < NX; i++) x0[i] = a.x0[b * NX + i];
    for (int i = 0; i < NU; i++) u0[i] = a.u0[b * NU + i];
    WarmFlags fl = a.flags[b];
    WsRef warm{a.warm + b, (size_t)a.B};
    cycle_prologue(P, T, a.Ts, a.N, x0, u0, warm, fl);
    a.flags[b] = fl;
    a.x0[b * NX + 7] = x0[7];
    a.x0[b * NX + 8] = x0[8];
    const size_t NS = (size_t)a.B * a.S;
    for (int k = 0; k < a.S; k++)
        for (int j = 0; j < DOF; j++) a.qs[(size_t)j * NS + (size_t)b * a.S + k] = warm[k * HZ + j];
}

// kinematic part of RobotData::update (robot_data.h:55-64): one thread per (instance, stage)
__global__ void k_kin(const double* __restrict__ qs, double* __restrict__ rb, int NS) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= NS) return;
    double q[DOF];
#pragma unroll
    for (int j = 0; j < DOF; j++) q[j] = qs[(size_t)j * NS + n];
    PandaKin kin;
    panda_kinematics(q, kin);
#pragma unroll
    for (int j = 0; j < DOF; j++) rb[(size_t)(RB_Q + j) * NS + n] = q[j];
#pragma unroll
    for (int i = 0; i < 3; i++) rb[(size_t)(RB_P + i) * NS + n] = kin.p[i];
#pragma unroll
    for (int i = 0; i < 9; i++) rb[(size_t)(RB_R + i) * NS + n] = kin.R[i];
#pragma unroll
    for (int i = 0; i < 21; i++) { rb[(size_t)(RB_JV + i) * NS + n] = kin.Jv[i]; rb[(size_t)(RB_JW + i) * NS + n] = kin.Jw[i]; }
    rb[(size_t)RB_MANIP * NS + n] = panda_manipulability_from(kin.Jv, kin.Jw);
    double dm[DOF];
    panda_dmanipulability(q, dm);
#pragma unroll
    for (int j = 0; j < DOF; j++) rb[(size_t)(RB_DMANIP + j) * NS + n] = dm[j];
}

// ---- probe kernels ------------------------------------------------------------------------------
__global__ void k_eval_stage(const Params* params, const TrackTable* tracks, double Ts, int N, const double* x, const double* u, const double* up,
                             const double* un, const double* xn, const double* rb, const int32_t* k, int n, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    StageLin sl;
    memset(&sl, 0, sizeof(sl));
    RbView rv{rb + (size_t)i * RB_DOUBLES, 1};
    stage_eval<true>(params[0], tracks[0], Ts, N, k[i], x + i * NX, u + i * NU, up + i * DOF, un + i * DOF, xn + i * NX, rv, sl);
    const double* src = (const double*)&sl;
    for (int e = 0; e < LIN_SIZE; e++) out[(size_t)i * LIN_SIZE + e] = src[e];
}
__global__ void k_eval_track(const TrackTable* tracks, const double* s, int n, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TrackPoint tp;
    track_eval_pos(tracks[0], s[i], tp);
    double* o = out + (size_t)i * 21;
    for (int c = 0; c < 3; c++) { o[c] = tp.pos[c]; o[3 + c] = tp.dpos[c]; o[6 + c] = tp.ddpos[c]; }
    track_eval_rot(tracks[0], s[i], o + 9, o + 18);
}
__global__ void k_transpose_rb_out(const double* rb_soa, int NS, int n, double* rb_aos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * RB_DOUBLES) return;
    int s = i / RB_DOUBLES, e = i % RB_DOUBLES;
    rb_aos[i] = rb_soa[(size_t)e * NS + s];
}
__global__ void k_scatter_q(const double* q_aos, int n, int NS, double* qs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NS * DOF) return;
    int j = i / NS, s = i % NS;
    qs[i] = (s < n) ? q_aos[s * DOF + j] : 0.0;
}
__global__ void k_warm_io(double* warm_soa, double* hor_aos, int B, int HN, int to_aos) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * HN) return;
    int b = (int)(i / HN), e = (int)(i % HN);
    if (to_aos) hor_aos[i] = warm_soa[(size_t)e * B + b];
    else warm_soa[(size_t)e * B + b] = hor_aos[i];
}
// [u0 (8) | status, iters packed into one double slot] per instance: the record the ranks exchange
__global__ void k_pack_results(const double* __restrict__ u_out, const int32_t* __restrict__ status, const int32_t* __restrict__ iters, int B, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 9) return;
    const int b = i / 9, e = i - b * 9;
    if (e < 8) out[i] = u_out[b * 8 + e];
    else { int2 v = make_int2(status[b], iters[b]); out[i] = *reinterpret_cast<double*>(&v); }
}
__global__ void k_invalidate_warm(WarmFlags* fl, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) fl[b].valid = 0;
}
// Integrator::simTimeStep (integrator.cpp:55-68): (int)(ts/1e-3) RK4 steps of the (linear) model
__global__ void k_sim_step(const double* x, const double* u, double ts, int B, double* xn) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double fine = 0.001;
    const int steps = (int)(ts / fine);
    double s[NX];
    for (int i = 0; i < NX; i++) s[i] = x[b * NX + i];
    const double* uu = u + b * NU;
    for (int it = 0; it < steps; it++) {
        // RK4 on f = [dq; vs; dVs] (model.cpp:31-45), written out stage by stage like integrator.cpp:29-43
        double k1[NX], k2[NX], k3[NX], k4[NX];
        for (int i = 0; i < 7; i++) k1[i] = k2[i] = k3[i] = k4[i] = uu[i];
        k1[8] = k2[8] = k3[8] = k4[8] = uu[7];
        k1[7] = s[8];
        k2[7] = s[8] + fine / 2. * k1[8];
        k3[7] = s[8] + fine / 2. * k2[8];
        k4[7] = s[8] + fine * k3[8];
        for (int i = 0; i < NX; i++) s[i] = s[i] + fine * (k1[i] / 6. + k2[i] / 3. + k3[i] / 3. + k4[i] / 6.);
    }
    for (int i = 0; i < NX; i++) xn[b * NX + i] = s[i];
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                                     \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess) return fail(MPCC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct mpcc_cuda_handle {
    mpcc_cuda_config cfg;
    int B, N, S;
    size_t NS;
    cudaStream_t stream = nullptr;
    bool have_nn = false, have_params = false, have_track = false;
    int n_param_sets = 0, n_tracks = 0;
    int num_sms = 148;
    // device memory
    Params* d_params = nullptr;
    TrackTable* d_tracks = nullptr;
    int32_t* d_track_id = nullptr;
    double *d_x0 = nullptr, *d_u0 = nullptr, *d_obs = nullptr, *d_obs_dummy = nullptr;
    double *d_warm = nullptr, *d_step = nullptr, *d_trial = nullptr, *d_filt = nullptr, *d_ws = nullptr, *d_qs = nullptr, *d_rb = nullptr;
    WarmFlags* d_flags = nullptr;
    double *d_u_out = nullptr, *d_horizon = nullptr;
    double* horizon_host = nullptr;  // this cycle's mapped host destination of the horizon (mpcc_cuda_run_cycle with a pinned buffer), else nullptr
    int32_t *d_status = nullptr, *d_iters = nullptr, *d_ok = nullptr, *d_qp_iters = nullptr, *d_qp_fail = nullptr, *d_accept = nullptr;
    long long* d_sqp_ns = nullptr;
    int32_t *d_hist = nullptr, *d_order = nullptr;
    double* d_wws = nullptr; size_t wws_per = 0, wsm_per = 0;  // warp-kernel workspace (doubles per instance / per warp)
    double *d_wpack = nullptr, *d_bias = nullptr, *d_w_out_env = nullptr, *d_w_out_self = nullptr;
    double *d_oz_dpack = nullptr, *d_oz_rowscale = nullptr;  // int8-split MLP kernel (mlp_oz_kernel.cuh)
    uint8_t* d_oz_wq = nullptr;
    long long* d_oz_dbg = nullptr;
    bool mlp_oz = false;
    int64_t launches = 0;
    bool profiling = false;
    bool use_cta = false;   // SQP kernel family of this handle: k_sqp_cta (one CTA per instance) or k_sqp_warp
    bool soc = false;       // some parameter set has sqp.do_SOC: the kernels of k_sqp_soc.cu (second-order correction compiled in) run instead
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // prologue | kin | mlp | sqp boundaries
    cudaStream_t aux = nullptr;                 // high-priority stream of the exclusive-SM straggler launch
    cudaEvent_t ev_pre = nullptr, ev_order = nullptr, ev_aux = nullptr;
    int32_t* hint = nullptr;                    // pinned host memory: [n(>= 2 iterations last cycle), n(>= 15 lately)], written by k_order
    std::vector<double> h_params;  // host copy of set 0 (validation)
    std::vector<void*> allocs;
    // multi-GPU result gather: packed [u0 (8) | status, iters (2 x int32)] per instance, double-buffered,
    // all-gathered on a side stream so that no rank ever waits for another inside its control cycle
    mpcc_ncclComm_t comm = nullptr;
    int rank = 0, world = 0;                    // world == 0: no communicator
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_pack[2] = {nullptr, nullptr}, ev_gath[2] = {nullptr, nullptr};
    double* d_stage[2] = {nullptr, nullptr};    // [B][9]
    double* d_gath[2] = {nullptr, nullptr};     // [world][B][9]
    double* h_gath = nullptr;                   // pinned, [world][B][9]
    int gslot = 0, glast = -1;
    bool gused[2] = {false, false};
    // grow-only device arena of the per-call probes / bindings (solveOCP at 100 Hz must not cudaMalloc per call)
    char* scratch = nullptr; size_t scratch_bytes = 0;
    cudaError_t need_scratch(size_t bytes) {
        if (bytes <= scratch_bytes) return cudaSuccess;
        if (scratch) { cudaStreamSynchronize(stream); cudaFree(scratch); scratch = nullptr; scratch_bytes = 0; }
        bytes = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMalloc((void**)&scratch, bytes);
        if (e == cudaSuccess) scratch_bytes = bytes;
        return e;
    }

    template <class T>
    cudaError_t alloc(T** p, size_t count) {
        cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
        if (e == cudaSuccess) { allocs.push_back(*p); e = cudaMemsetAsync(*p, 0, count * sizeof(T), stream); }
        return e;
    }
};

static CycleArgs make_args(mpcc_cuda_handle* h, double* d_x0, const double* d_u0, const double* d_obs) {
    CycleArgs a;
    a.B = h->B; a.N = h->N; a.S = h->S; a.Ts = h->cfg.Ts;
    a.params = h->d_params; a.params_per_instance = (h->n_param_sets > 1) ? 1 : 0;
    a.tracks = h->d_tracks; a.track_id = h->d_track_id;
    a.x0 = d_x0; a.u0 = d_u0; a.obs = d_obs;
    a.warm = h->d_warm; a.step = h->d_step; a.trial = h->d_trial; a.filt = h->d_filt; a.ws = h->d_ws; a.flags = h->d_flags;
    a.qs = h->d_qs; a.rb = h->d_rb; a.u_out = h->d_u_out; a.horizon = h->d_horizon; a.horizon_host = h->horizon_host;
    a.status = h->d_status; a.iters = h->d_iters; a.ok = h->d_ok; a.qp_iters = h->d_qp_iters; a.qp_fail = h->d_qp_fail; a.accept_mask = h->d_accept; a.sqp_ns = h->d_sqp_ns; a.hist = h->d_hist; a.order = h->d_order;
    a.qp = QpOptions{h->cfg.qp_max_iter, h->cfg.qp_eps};
    return a;
}

static int launch_robot_data(mpcc_cuda_handle* h, const double* d_obs, int S_for_obs, bool mark = false) {
    const int NS = (int)h->NS;
    k_kin<<<(NS + 127) / 128, 128, 0, h->stream>>>(h->d_qs, h->d_rb, NS);
    if (mark) cudaEventRecord(h->ev[2], h->stream);
    MlpArgs m;
    m.wpack = h->d_wpack; m.bias = h->d_bias; m.w_out_env = h->d_w_out_env; m.w_out_self = h->d_w_out_self;
    m.qs = h->d_qs; m.obs = d_obs; m.rb = h->d_rb; m.NS = NS; m.S = S_for_obs;
    m.n_tiles = (NS + MLP_TILE_S - 1) / MLP_TILE_S;
    int grid = m.n_tiles < h->num_sms ? m.n_tiles : h->num_sms;
    if (h->mlp_oz) {
        MlpOzArgs oa;
        oa.m = m; oa.m.wpack = h->d_oz_dpack; oa.wq = h->d_oz_wq; oa.rowscale = h->d_oz_rowscale; oa.dbg = h->d_oz_dbg; oa.dbg_flags = ((h->cfg.reserved >> 5) & 3) | (((h->cfg.reserved >> 9) & 1) << 2);
        if (h->cfg.reserved & 128) grid = 16;  // experiment: few CTAs (is the weight stream limited per SM or by the whole chip's L2 traffic?)
        k_mlp_oz<<<grid, MLP_THREADS, OZ_SMEM_BYTES, h->stream>>>(oa);
    } else {
        k_mlp<<<grid, MLP_THREADS, MLP_SMEM_BYTES, h->stream>>>(m);
    }
    h->launches += 2;
    CK(cudaGetLastError());
    return MPCC_OK;
}

static int check_ready(mpcc_cuda_handle* h) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    if (!h->have_nn) return fail(MPCC_ERR_STATE, "networks not uploaded (mpcc_cuda_upload_nn / mpcc_cuda_load_nn)");
    if (!h->have_params) return fail(MPCC_ERR_STATE, "parameters not set (mpcc_cuda_set_params)");
    if (!h->have_track) return fail(MPCC_ERR_STATE, "track not set (mpcc_cuda_set_tracks)");
    return MPCC_OK;
}

extern "C" {

const char* mpcc_cuda_last_error(void) { return g_err.c_str(); }

static int create_impl(mpcc_cuda_handle* h, const mpcc_cuda_config* cfg);
int mpcc_cuda_create(const mpcc_cuda_config* cfg, mpcc_cuda_handle** out) {
    if (!cfg || !out) return fail(MPCC_ERR_INVALID, "null argument");
    if (cfg->batch < 1) return fail(MPCC_ERR_INVALID, "batch must be >= 1");
    if (cfg->horizon < 2 || cfg->horizon > MAX_N) return fail(MPCC_ERR_INVALID, "horizon must be in [2, 64]");
    if (!(cfg->Ts > 0)) return fail(MPCC_ERR_INVALID, "Ts must be positive");
    if (cfg->sqp_kernel != 0) return fail(MPCC_ERR_INVALID, "sqp_kernel must be 0: the library ships one SQP kernel (warp per instance)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(MPCC_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MPCC_ERR_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(cfg->device));
    mpcc_cuda_handle* h = new mpcc_cuda_handle();
    h->cfg = *cfg;
    const int rc = create_impl(h, cfg);
    if (rc != MPCC_OK) { const std::string keep = g_err; mpcc_cuda_destroy(h); g_err = keep; return rc; }  // nothing of a half-built handle leaks
    *out = h;
    return MPCC_OK;
}
static int create_impl(mpcc_cuda_handle* h, const mpcc_cuda_config* cfg) {
    if (h->cfg.qp_max_iter <= 0) h->cfg.qp_max_iter = 60;
    if (!(h->cfg.qp_eps > 0)) h->cfg.qp_eps = 1e-9;
    h->B = cfg->batch; h->N = cfg->horizon; h->S = cfg->horizon + 1; h->NS = (size_t)h->B * h->S;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    h->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_aux, cudaEventDisableTiming));
        CK(cudaHostAlloc((void**)&h->hint, 2 * sizeof(int32_t), cudaHostAllocPortable));
        h->hint[0] = h->B; h->hint[1] = 0;  // the first cycle is a cold start
    }
    const size_t B = h->B, S = h->S, HN = S * HZ;
    cudaError_t ae = cudaSuccess;
    auto A = [&](cudaError_t r) { if (ae == cudaSuccess) ae = r; };
    A(h->alloc(&h->d_track_id, B));
    A(h->alloc(&h->d_x0, B * NX)); A(h->alloc(&h->d_u0, B * NU)); A(h->alloc(&h->d_obs, B * 4)); A(h->alloc(&h->d_obs_dummy, B * 4));
    A(h->alloc(&h->d_warm, B * HN)); A(h->alloc(&h->d_step, B * HN)); A(h->alloc(&h->d_trial, B * HN));
    A(h->alloc(&h->d_filt, B * FILT_DOUBLES)); A(h->alloc(&h->d_ws, B * S * STAGE_WS));
    h->wws_per = sqp_warp_ws_doubles(h->N);
    A(h->alloc(&h->d_wws, B * h->wws_per));
    A(h->alloc(&h->d_qs, h->NS * DOF)); A(h->alloc(&h->d_rb, h->NS * RB_DOUBLES));
    A(h->alloc(&h->d_flags, B));
    A(h->alloc(&h->d_u_out, B * NU)); A(h->alloc(&h->d_horizon, B * HN));
    A(h->alloc(&h->d_status, B)); A(h->alloc(&h->d_iters, B)); A(h->alloc(&h->d_ok, B)); A(h->alloc(&h->d_qp_iters, B)); A(h->alloc(&h->d_qp_fail, B)); A(h->alloc(&h->d_accept, B)); A(h->alloc(&h->d_sqp_ns, 4 * B)); A(h->alloc(&h->d_hist, B)); A(h->alloc(&h->d_order, B + 1));
    A(h->alloc(&h->d_wpack, (size_t)MLP_NCHUNK * MLP_CHUNK_D)); A(h->alloc(&h->d_bias, MLP_BIAS_TOTAL));
    A(h->alloc(&h->d_w_out_env, 9 * 256)); A(h->alloc(&h->d_w_out_self, 64));
    A(h->alloc(&h->d_oz_dpack, OZ_DPACK_D)); A(h->alloc(&h->d_oz_rowscale, 3 * 256)); A(h->alloc(&h->d_oz_wq, (size_t)OZ_CHUNKS_PER_TILE * OZ_CHUNK));
    h->mlp_oz = (h->cfg.reserved & 8) == 0;  // default: the int8-split tcgen05 kernel; bit 3 selects the fp64 DMMA kernel (k_mlp)
    if (h->cfg.reserved & 16) A(h->alloc(&h->d_oz_dbg, 64));
    if (ae != cudaSuccess) return fail(MPCC_ERR_CUDA, std::string("device allocation failed: ") + cudaGetErrorString(ae));
    std::vector<double> dummy(B * 4);
    for (size_t b = 0; b < B; b++) { dummy[4 * b] = 3; dummy[4 * b + 1] = 3; dummy[4 * b + 2] = 3; dummy[4 * b + 3] = 0; }  // mpc.cpp:97-100
    CK(cudaMemcpyAsync(h->d_obs_dummy, dummy.data(), dummy.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaFuncSetAttribute(k_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_mlp_oz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM_BYTES));
    CK(configure_sqp_warp(h->N));
    CK(configure_sqp_cta());
    CK(configure_sqp_soc());
    // kernel family: one CTA per instance (latency path) when the batch cannot fill the machine with warps anyway
    h->use_cta = (h->cfg.reserved & 4) ? true : (h->cfg.reserved & 2) ? false : (h->B <= 2 * h->num_sms);
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_destroy(mpcc_cuda_handle* h) {
    if (!h) return MPCC_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (cudaEvent_t e : h->ev) if (e) cudaEventDestroy(e);
    for (void* p : h->allocs) cudaFree(p);
    if (h->scratch) cudaFree(h->scratch);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (int i = 0; i < 2; i++) {
        if (h->d_stage[i]) cudaFree(h->d_stage[i]);
        if (h->d_gath[i]) cudaFree(h->d_gath[i]);
        if (h->ev_pack[i]) cudaEventDestroy(h->ev_pack[i]);
        if (h->ev_gath[i]) cudaEventDestroy(h->ev_gath[i]);
    }
    if (h->h_gath) cudaFreeHost(h->h_gath);
    if (h->gstream) cudaStreamDestroy(h->gstream);
    if (h->d_params) cudaFree(h->d_params);
    if (h->d_tracks) cudaFree(h->d_tracks);
    if (h->aux) cudaStreamDestroy(h->aux);
    if (h->hint) cudaFreeHost(h->hint);
    if (h->ev_pre) cudaEventDestroy(h->ev_pre);
    if (h->ev_order) cudaEventDestroy(h->ev_order);
    if (h->ev_aux) cudaEventDestroy(h->ev_aux);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MPCC_OK;
}

int mpcc_cuda_upload_nn(mpcc_cuda_handle* h, const double* self_w, const double* self_b, const double* env_w, const double* env_b) {
    if (!h || !self_w || !self_b || !env_w || !env_b) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const double* eW[5];
    const double* sW[3];
    const int e_out[5] = {256, 256, 256, 256, 9}, e_in[5] = {30, 256, 256, 256, 256};
    const int s_out[3] = {256, 64, 1}, s_in[3] = {21, 256, 64};
    size_t o = 0;
    for (int l = 0; l < 5; l++) { eW[l] = env_w + o; o += (size_t)e_out[l] * e_in[l]; }
    o = 0;
    for (int l = 0; l < 3; l++) { sW[l] = self_w + o; o += (size_t)s_out[l] * s_in[l]; }
    std::vector<double> pack((size_t)MLP_NCHUNK * MLP_CHUNK_D);
    pack_mlp_weights(eW, sW, pack.data());
    std::vector<double> bias(MLP_BIAS_TOTAL);
    std::memcpy(&bias[MLP_BIAS_ENV], env_b, (4 * 256 + 9) * 8);
    std::memcpy(&bias[MLP_BIAS_SELF0], self_b, (256 + 64 + 1) * 8);
    CK(cudaMemcpyAsync(h->d_wpack, pack.data(), pack.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_bias, bias.data(), bias.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_w_out_env, eW[4], 9 * 256 * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_w_out_self, sW[2], 64 * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<double> oz_d(OZ_DPACK_D), oz_rs(3 * 256);
    std::vector<uint8_t> oz_q((size_t)OZ_CHUNKS_PER_TILE * OZ_CHUNK);
    pack_mlp_oz_weights(eW, sW, oz_d.data(), oz_q.data(), oz_rs.data());
    CK(cudaMemcpyAsync(h->d_oz_dpack, oz_d.data(), oz_d.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_oz_rowscale, oz_rs.data(), oz_rs.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_oz_wq, oz_q.data(), oz_q.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_nn = true;
    return MPCC_OK;
}

int mpcc_cuda_load_nn(mpcc_cuda_handle* h, const char* self_path, const char* env_path) {
    if (!h || !self_path || !env_path) return fail(MPCC_ERR_INVALID, "null argument");
    try {
        auto load = [](const std::string& p, const std::vector<std::pair<int, int>>& dims) {
            bool packed = p.size() > 4 && p.substr(p.size() - 4) == ".f64";
            MlpWeights m = packed ? load_mlp_packed(p) : load_mlp_text(p, dims);
            if (m.W.size() != dims.size()) throw std::runtime_error("nn: unexpected layer count in '" + p + "'");
            for (size_t l = 0; l < dims.size(); l++)
                if (m.out_dim[l] != dims[l].first || m.in_dim[l] != dims[l].second) throw std::runtime_error("nn: unexpected layer shape in '" + p + "'");
            return m;
        };
        MlpWeights s = load(self_path, {{256, 21}, {64, 256}, {1, 64}});
        MlpWeights e = load(env_path, {{256, 30}, {256, 256}, {256, 256}, {256, 256}, {9, 256}});
        std::vector<double> sw, sb, ew, eb;
        for (auto& w : s.W) sw.insert(sw.end(), w.begin(), w.end());
        for (auto& b : s.b) sb.insert(sb.end(), b.begin(), b.end());
        for (auto& w : e.W) ew.insert(ew.end(), w.begin(), w.end());
        for (auto& b : e.b) eb.insert(eb.end(), b.begin(), b.end());
        return mpcc_cuda_upload_nn(h, sw.data(), sb.data(), ew.data(), eb.data());
    } catch (const std::exception& ex) {
        return fail(MPCC_ERR_IO, ex.what());
    }
}

int mpcc_cuda_set_params(mpcc_cuda_handle* h, const double* params, int32_t n_sets) {
    if (!h || !params) return fail(MPCC_ERR_INVALID, "null argument");
    if (n_sets != 1 && n_sets != h->B) return fail(MPCC_ERR_INVALID, "n_sets must be 1 or batch");
    bool any_soc = false;
    for (int s = 0; s < n_sets; s++) {
        const Params& p = *(const Params*)(params + (size_t)s * PARAMS_DOUBLES);
        if (p.max_iter < 1 || p.max_iter > MAX_SQP_ITER) return fail(MPCC_ERR_INVALID, "sqp.max_iter must be in [1, 128]");
        if (p.line_search_max_iter < 1) return fail(MPCC_ERR_INVALID, "sqp.line_search_max_iter must be >= 1");
        if (p.use_BFGS != 0) return fail(MPCC_ERR_INVALID, "sqp.use_BFGS is not implemented on this path (reference default: false; DESIGN.md 9)");
        any_soc = any_soc || p.do_SOC != 0;
        for (int i = 0; i < NX; i++) if (!(p.Tx[i] > 0)) return fail(MPCC_ERR_INVALID, "normalization entries must be positive");
        for (int i = 0; i < NU; i++) if (!(p.Tu[i] > 0)) return fail(MPCC_ERR_INVALID, "normalization entries must be positive");
    }
    CK(cudaSetDevice(h->cfg.device));
    if (h->d_params && h->n_param_sets != n_sets) { CK(cudaFree(h->d_params)); h->d_params = nullptr; }
    if (!h->d_params) CK(cudaMalloc((void**)&h->d_params, (size_t)n_sets * sizeof(Params)));
    CK(cudaMemcpyAsync(h->d_params, params, (size_t)n_sets * sizeof(Params), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_param_sets = n_sets;
    h->soc = any_soc;
    h->h_params.assign(params, params + PARAMS_DOUBLES);
    h->have_params = true;
    return MPCC_OK;
}

int mpcc_load_params_json(const char* model_path, const char* cost_path, const char* bounds_path, const char* normalization_path,
                          const char* sqp_path, const char* const* over_keys, const double* over_vals, int32_t n_over, double* params_out) {
    if (!model_path || !cost_path || !bounds_path || !normalization_path || !sqp_path || !params_out) return fail(MPCC_ERR_INVALID, "null argument");
    try {
        PathToJson p;
        p.param_path = model_path; p.cost_path = cost_path; p.bounds_path = bounds_path; p.normalization_path = normalization_path; p.sqp_path = sqp_path;
        ParamValue ov;
        for (int i = 0; i < n_over; i++) {
            std::string k = over_keys[i];
            size_t dot = k.find('.');
            if (dot == std::string::npos) return fail(MPCC_ERR_INVALID, "override key must be written file.key: " + k);
            std::string f = k.substr(0, dot), key = k.substr(dot + 1);
            if (f == "model" || f == "param") ov.param[key] = over_vals[i];
            else if (f == "cost") ov.cost[key] = over_vals[i];
            else if (f == "bounds") ov.bounds[key] = over_vals[i];
            else if (f == "normalization") ov.normalization[key] = over_vals[i];
            else if (f == "sqp") ov.sqp[key] = over_vals[i];
            else return fail(MPCC_ERR_INVALID, "unknown override file: " + f);
        }
        Params pr = load_params(p, ov);
        std::memcpy(params_out, &pr, sizeof(pr));
        return MPCC_OK;
    } catch (const std::exception& ex) {
        return fail(MPCC_ERR_IO, ex.what());
    }
}

int mpcc_fit_track(int32_t n, const double* X, const double* Y, const double* Z, const double* R, double* table_out) {
    if (!X || !Y || !Z || !R || !table_out) return fail(MPCC_ERR_INVALID, "null argument");
    if (n < 3) return fail(MPCC_ERR_INVALID, "a track needs at least 3 waypoints");
    try {
        Waypoints w;
        w.X.assign(X, X + n); w.Y.assign(Y, Y + n); w.Z.assign(Z, Z + n); w.R.assign(R, R + (size_t)9 * n);
        fit_track(w, *(TrackTable*)table_out);
        return MPCC_OK;
    } catch (const std::exception& ex) {
        return fail(MPCC_ERR_INVALID, ex.what());
    }
}

// Bulk track ingestion for heterogeneous batches (configuration C4): n_tracks independent fits on the host cores.
int mpcc_fit_tracks(int32_t n_tracks, int32_t n, const double* X, const double* Y, const double* Z, const double* R, double* tables_out, int32_t n_threads) {
    if (!X || !Y || !Z || !R || !tables_out || n_tracks < 1 || n < 3) return fail(MPCC_ERR_INVALID, "bad argument (n_tracks >= 1, at least 3 waypoints per track)");
    unsigned hw = std::thread::hardware_concurrency();
    int nt = n_threads > 0 ? n_threads : (int)(hw ? hw : 1);
    if (nt > n_tracks) nt = n_tracks;
    std::vector<std::string> errs(nt);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&, t]() {
            try {
                for (int i = t; i < n_tracks; i += nt) {
                    Waypoints w;
                    const size_t o = (size_t)i * n;
                    w.X.assign(X + o, X + o + n); w.Y.assign(Y + o, Y + o + n); w.Z.assign(Z + o, Z + o + n); w.R.assign(R + 9 * o, R + 9 * (o + n));
                    fit_track(w, *(TrackTable*)(tables_out + (size_t)i * TRACK_DOUBLES));
                }
            } catch (const std::exception& ex) { errs[t] = ex.what(); }
        });
    for (auto& th : pool) th.join();
    for (auto& e : errs) if (!e.empty()) return fail(MPCC_ERR_INVALID, e);
    return MPCC_OK;
}

int mpcc_load_track_json(const char* track_path, const double* init_position3, double* table_out) {
    if (!track_path || !table_out) return fail(MPCC_ERR_INVALID, "null argument");
    try {
        Waypoints w = load_track_json(track_path);
        if (init_position3) shift_track(w, init_position3);
        fit_track(w, *(TrackTable*)table_out);
        return MPCC_OK;
    } catch (const std::exception& ex) {
        return fail(MPCC_ERR_IO, ex.what());
    }
}

// Track ingestion ON THE DEVICE for heterogeneous batches: waypoints of n_tracks tracks in, fitted tables installed
// as the handle's tracks (same effect as mpcc_fit_tracks + mpcc_cuda_set_tracks, without the host fit).
int mpcc_cuda_fit_tracks(mpcc_cuda_handle* h, int32_t n_tracks, int32_t n, const double* X, const double* Y, const double* Z, const double* R,
                         const int32_t* track_of_instance) {
    if (!h || !X || !Y || !Z || !R || n_tracks < 1) return fail(MPCC_ERR_INVALID, "bad argument");
    if (n < 3 || n > 4096) return fail(MPCC_ERR_INVALID, "a track needs 3 .. 4096 waypoints");
    std::vector<int32_t> ids(h->B, 0);
    if (track_of_instance)
        for (int b = 0; b < h->B; b++) {
            if (track_of_instance[b] < 0 || track_of_instance[b] >= n_tracks) return fail(MPCC_ERR_INVALID, "track index out of range");
            ids[b] = track_of_instance[b];
        }
    CK(cudaSetDevice(h->cfg.device));
    if (h->d_tracks && h->n_tracks != n_tracks) { CK(cudaFree(h->d_tracks)); h->d_tracks = nullptr; }
    if (!h->d_tracks) CK(cudaMalloc((void**)&h->d_tracks, (size_t)n_tracks * sizeof(TrackTable)));
    const int chunk = n_tracks < 4096 ? n_tracks : 4096;
    const size_t per = track_fit_scratch_doubles(n), wp = (size_t)chunk * n;
    CK(h->need_scratch((12 * wp + per * chunk) * 8));
    double* dX = (double*)h->scratch; double* dY = dX + wp; double* dZ = dY + wp; double* dR = dZ + wp; double* dS = dR + 9 * wp;
    for (int t0 = 0; t0 < n_tracks; t0 += chunk) {
        const int nt = (n_tracks - t0 < chunk) ? n_tracks - t0 : chunk;
        const size_t o = (size_t)t0 * n, cnt = (size_t)nt * n;
        CK(cudaMemcpyAsync(dX, X + o, cnt * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dY, Y + o, cnt * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dZ, Z + o, cnt * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dR, R + 9 * o, cnt * 72, cudaMemcpyHostToDevice, h->stream));
        launch_fit_tracks(nt, n, dX, dY, dZ, dR, dS, h->d_tracks + t0, h->stream);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(h->d_track_id, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, h->stream));
    k_invalidate_warm<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->d_flags, h->B);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->n_tracks = n_tracks;
    h->have_track = true;
    if (h->hint) h->hint[0] = h->B;
    return MPCC_OK;
}

int mpcc_cuda_get_tracks(mpcc_cuda_handle* h, double* tables_out, int32_t n_tracks) {
    if (!h || !tables_out) return fail(MPCC_ERR_INVALID, "null argument");
    if (!h->have_track) return fail(MPCC_ERR_STATE, "track not set");
    if (n_tracks < 1 || n_tracks > h->n_tracks) return fail(MPCC_ERR_INVALID, "n_tracks out of range");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(tables_out, h->d_tracks, (size_t)n_tracks * sizeof(TrackTable), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

// Table of an ALREADY FITTED ArcLengthSpline: its N_SPLINE = 100 knots (arc lengths, positions, rotations) in, no fit / resample
// pass (the binding of SolverInterface::setTrack(ArcLengthSpline)): only the final regular spline
// (arc_length_spline.cpp:244-252) is rebuilt from the knots.
int mpcc_track_from_knots(const double* s, const double* X, const double* Y, const double* Z, const double* R, double* table_out) {
    if (!s || !X || !Y || !Z || !R || !table_out) return fail(MPCC_ERR_INVALID, "null argument");
    for (int i = 0; i + 1 < N_SPLINE; i++) if (!(s[i + 1] > s[i])) return fail(MPCC_ERR_INVALID, "knot arc lengths must increase strictly");
    std::vector<double> w((size_t)14 * N_SPLINE);
    tf_table_from_knots(TArr{(double*)s, 1}, TArr{(double*)X, 1}, TArr{(double*)Y, 1}, TArr{(double*)Z, 1}, TArr{(double*)R, 1}, TArr{w.data(), 1}, N_SPLINE, *(TrackTable*)table_out);
    return MPCC_OK;
}

int mpcc_cuda_reset(mpcc_cuda_handle* h) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemsetAsync(h->d_flags, 0, (size_t)h->B * sizeof(WarmFlags), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->hint) h->hint[0] = h->B;  // a cold start is a transient: every instance needs several SQP iterations
    return MPCC_OK;
}

int mpcc_cuda_set_tracks(mpcc_cuda_handle* h, const double* tables, int32_t n_tracks, const int32_t* track_of_instance) {
    if (!h || !tables || n_tracks < 1) return fail(MPCC_ERR_INVALID, "bad argument");
    std::vector<int32_t> ids(h->B, 0);
    if (track_of_instance)
        for (int b = 0; b < h->B; b++) {
            if (track_of_instance[b] < 0 || track_of_instance[b] >= n_tracks) return fail(MPCC_ERR_INVALID, "track index out of range");
            ids[b] = track_of_instance[b];
        }
    CK(cudaSetDevice(h->cfg.device));
    if (h->d_tracks && h->n_tracks != n_tracks) { CK(cudaFree(h->d_tracks)); h->d_tracks = nullptr; }
    if (!h->d_tracks) CK(cudaMalloc((void**)&h->d_tracks, (size_t)n_tracks * sizeof(TrackTable)));
    CK(cudaMemcpyAsync(h->d_tracks, tables, (size_t)n_tracks * sizeof(TrackTable), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_track_id, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_tracks = n_tracks;
    h->have_track = true;
    // MPC::setTrack (mpc.cpp:192-197) clears valid_initial_guess_ only; num_valid_guess_failed_ keeps counting
    k_invalidate_warm<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->d_flags, h->B);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    if (h->hint) h->hint[0] = h->B;
    return MPCC_OK;
}

int mpcc_cuda_run_cycle_device(mpcc_cuda_handle* h, double* d_x0, const double* d_u0, const double* d_obs) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!d_x0 || !d_u0) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const double* obs = d_obs ? d_obs : h->d_obs_dummy;
    CycleArgs a = make_args(h, d_x0, d_u0, obs);
    h->launches = 0;
    const bool prof = h->profiling;
    if (prof) cudaEventRecord(h->ev[0], h->stream);
    k_prologue<<<(h->B + 63) / 64, 64, 0, h->stream>>>(a);
    h->launches++;
    if (prof) cudaEventRecord(h->ev[1], h->stream);
    rc = launch_robot_data(h, obs, h->S, prof);
    if (rc) return rc;
    if (prof) cudaEventRecord(h->ev[3], h->stream);
    if (h->soc) launch_sqp_soc(a, h->d_wws, h->use_cta, h->stream);
    else if (h->use_cta) launch_sqp_cta(a, h->d_wws, h->stream);
    else {
        launch_sqp_warp(a, h->d_wws, h->stream, (h->cfg.reserved & 1) ? nullptr : h->aux, h->ev_pre, h->ev_order, h->ev_aux, h->hint);
        h->launches += (h->cfg.reserved & 1) ? 1 : 2;  // + the launch-order kernel (+ the exclusive launch)
    }
    h->launches++;
    if (prof) cudaEventRecord(h->ev[4], h->stream);
    CK(cudaGetLastError());
    return MPCC_OK;
}

int mpcc_cuda_read_results(mpcc_cuda_handle* h, double* u_out, double* horizon, int32_t* status, int32_t* sqp_iters, int32_t* ok) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    const size_t B = h->B;
    if (u_out) CK(cudaMemcpyAsync(u_out, h->d_u_out, B * NU * 8, cudaMemcpyDeviceToHost, h->stream));
    if (horizon) CK(cudaMemcpyAsync(horizon, h->d_horizon, B * h->S * HZ * 8, cudaMemcpyDeviceToHost, h->stream));
    if (status) CK(cudaMemcpyAsync(status, h->d_status, B * 4, cudaMemcpyDeviceToHost, h->stream));
    if (sqp_iters) CK(cudaMemcpyAsync(sqp_iters, h->d_iters, B * 4, cudaMemcpyDeviceToHost, h->stream));
    if (ok) CK(cudaMemcpyAsync(ok, h->d_ok, B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_run_cycle(mpcc_cuda_handle* h, double* x0, const double* u0, const double* obs, double* u_out, double* horizon,
                        int32_t* status, int32_t* sqp_iters, int32_t* ok) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x0 || !u0) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const size_t B = h->B;
    CK(cudaMemcpyAsync(h->d_x0, x0, B * NX * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_u0, u0, B * NU * 8, cudaMemcpyHostToDevice, h->stream));
    if (obs) CK(cudaMemcpyAsync(h->d_obs, obs, B * 4 * 8, cudaMemcpyHostToDevice, h->stream));
    // MPCReturn::mpc_horizon is by far the largest result (B x (N+1) x 17 doubles).  If the caller's buffer is pinned (mapped) host memory the SQP
    // kernel writes it there directly as each instance finishes, so it crosses PCIe under the kernel; otherwise it is copied behind the kernel.
    double* hor_mapped = nullptr;
    if (horizon && !(h->cfg.reserved & 256)) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, horizon) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) hor_mapped = (double*)at.devicePointer;
        else cudaGetLastError();
    }
    h->horizon_host = hor_mapped;
    rc = mpcc_cuda_run_cycle_device(h, h->d_x0, h->d_u0, obs ? h->d_obs : nullptr);
    h->horizon_host = nullptr;
    if (rc) return rc;
    CK(cudaMemcpyAsync(x0, h->d_x0, B * NX * 8, cudaMemcpyDeviceToHost, h->stream));
    return mpcc_cuda_read_results(h, u_out, hor_mapped ? nullptr : horizon, status, sqp_iters, ok);
}

int mpcc_cuda_result_pointers(mpcc_cuda_handle* h, double** d_u_out, double** d_horizon, int32_t** d_status, int32_t** d_iters, int32_t** d_ok) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    if (d_u_out) *d_u_out = h->d_u_out;
    if (d_horizon) *d_horizon = h->d_horizon;
    if (d_status) *d_status = h->d_status;
    if (d_iters) *d_iters = h->d_iters;
    if (d_ok) *d_ok = h->d_ok;
    return MPCC_OK;
}
void* mpcc_cuda_stream(mpcc_cuda_handle* h) { return h ? (void*)h->stream : nullptr; }
int mpcc_cuda_synchronize(mpcc_cuda_handle* h) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_get_warm_state(mpcc_cuda_handle* h, double* horizon, int32_t* valid, int32_t* failed) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    const int HN = h->S * HZ;
    const size_t tot = (size_t)h->B * HN;
    if (horizon) {
        k_warm_io<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->d_warm, h->d_trial, h->B, HN, 1);  // d_trial is free between cycles
        CK(cudaGetLastError());
        // d_trial is SoA-sized scratch; here it temporarily holds the AoS copy
        CK(cudaMemcpyAsync(horizon, h->d_trial, tot * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    std::vector<WarmFlags> fl(h->B);
    CK(cudaMemcpyAsync(fl.data(), h->d_flags, fl.size() * sizeof(WarmFlags), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < h->B; b++) { if (valid) valid[b] = fl[b].valid; if (failed) failed[b] = fl[b].failed; }
    return MPCC_OK;
}
int mpcc_cuda_set_warm_state(mpcc_cuda_handle* h, const double* horizon, const int32_t* valid, const int32_t* failed) {
    if (!h || !horizon || !valid || !failed) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const int HN = h->S * HZ;
    const size_t tot = (size_t)h->B * HN;
    CK(cudaMemcpyAsync(h->d_trial, horizon, tot * 8, cudaMemcpyHostToDevice, h->stream));
    k_warm_io<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->d_warm, h->d_trial, h->B, HN, 0);
    CK(cudaGetLastError());
    std::vector<WarmFlags> fl(h->B);
    for (int b = 0; b < h->B; b++) { fl[b].valid = valid[b]; fl[b].failed = failed[b]; }
    CK(cudaMemcpyAsync(h->d_flags, fl.data(), fl.size() * sizeof(WarmFlags), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_sim_time_step(mpcc_cuda_handle* h, const double* x, const double* u, double ts, double* x_next) {
    if (!h || !x || !u || !x_next) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const size_t B = h->B;
    double* d_x = h->d_trial;            // scratch between cycles
    double* d_xn = h->d_trial + B * NX;
    CK(cudaMemcpyAsync(d_x, x, B * NX * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_u0, u, B * NU * 8, cudaMemcpyHostToDevice, h->stream));
    k_sim_step<<<(h->B + 127) / 128, 128, 0, h->stream>>>(d_x, h->d_u0, ts, h->B, d_xn);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x_next, d_xn, B * NX * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_eval_robot_data(mpcc_cuda_handle* h, const double* q, const double* obs, int32_t n, double* rb_out) {
    if (!h || !q || !rb_out) return fail(MPCC_ERR_INVALID, "null argument");
    if (!h->have_nn) return fail(MPCC_ERR_STATE, "networks not uploaded");
    if (n < 1 || (size_t)n > h->NS) return fail(MPCC_ERR_INVALID, "n must be in [1, batch*(N+1)]");
    CK(cudaSetDevice(h->cfg.device));
    const int NS = (int)h->NS;
    // stage the AoS inputs in the (idle) workspace, obstacle per SAMPLE (S = 1)
    double* d_q = h->d_ws;
    double* d_o = h->d_ws + (size_t)NS * DOF;
    double* d_out = d_o + (size_t)NS * 4;
    if ((size_t)NS * (DOF + 4 + RB_DOUBLES) > (size_t)h->B * h->S * STAGE_WS) return fail(MPCC_ERR_INVALID, "workspace too small");
    CK(cudaMemcpyAsync(d_q, q, (size_t)n * DOF * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<double> o((size_t)NS * 4);
    for (int i = 0; i < NS; i++) {
        if (obs && i < n) for (int c = 0; c < 4; c++) o[4 * (size_t)i + c] = obs[4 * (size_t)i + c];
        else { o[4 * (size_t)i] = 3; o[4 * (size_t)i + 1] = 3; o[4 * (size_t)i + 2] = 3; o[4 * (size_t)i + 3] = 0; }
    }
    CK(cudaMemcpyAsync(d_o, o.data(), o.size() * 8, cudaMemcpyHostToDevice, h->stream));
    k_scatter_q<<<(NS * DOF + 255) / 256, 256, 0, h->stream>>>(d_q, n, NS, h->d_qs);
    int rc = launch_robot_data(h, d_o, 1);
    if (rc) return rc;
    k_transpose_rb_out<<<(n * RB_DOUBLES + 255) / 256, 256, 0, h->stream>>>(h->d_rb, NS, n, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rb_out, d_out, (size_t)n * RB_DOUBLES * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_oz_dbg) {  // diagnostics (reserved bit 4): cycles of CTA 0 of k_mlp_oz per phase
        long long t[64];
        CK(cudaMemcpy(t, h->d_oz_dbg, sizeof(t), cudaMemcpyDeviceToHost));
        fprintf(stderr, "k_mlp_oz CTA 0, %lld tiles: cycles per tile: first layers %lld | split %lld | MMA passes %lld | epilogues %lld | env output %lld | self net %lld || issuer waiting for weight chunks %lld\n", t[6],
                t[0] / t[6], t[1] / t[6], t[2] / t[6], t[3] / t[6], t[4] / t[6], t[5] / t[6], t[7] / t[6]);
        fprintf(stderr, "  trace (tile 1, layer 0): pass 0: issue %lld, until complete %lld | epilogue + barrier %lld | pass 1: issue %lld, until complete %lld ; chunk waits of the issuer per tile %lld\n",
                t[9] - t[8], t[10] - t[8], t[11] - t[10], t[12] - t[11], t[13] - t[11], t[7] / t[6]);
        if (h->cfg.reserved & 512) fprintf(stderr, "  passes issued twice: first (cold) %lld, second (warm) %lld cycles per pass\n", t[40] / (6 * t[6]), t[41] / (6 * t[6]));
        fprintf(stderr, "  first layers (tile 1): env: staging %lld | 4 chunks of DFMA %lld | epilogue + column maxima %lld | barrier %lld ;  self: staging %lld | DFMA %lld | epilogue %lld | barrier %lld\n",
                t[25] - t[24], t[26] - t[25], t[27] - t[26], t[28] - t[27], t[30] - t[29], t[31] - t[30], t[32] - t[31], t[33] - t[32]);
    }
    return MPCC_OK;
}

int mpcc_cuda_eval_stage(mpcc_cuda_handle* h, const double* x, const double* u, const double* u_prev, const double* u_next,
                         const double* x_next, const double* rb, const int32_t* k, int32_t n, double* lin_out) {
    if (!h || !x || !u || !u_prev || !u_next || !x_next || !rb || !k || !lin_out) return fail(MPCC_ERR_INVALID, "null argument");
    if (!h->have_params || !h->have_track) return fail(MPCC_ERR_STATE, "parameters / track not set");
    if (n < 1) return fail(MPCC_ERR_INVALID, "n must be >= 1");
    CK(cudaSetDevice(h->cfg.device));
    const size_t per = NX + NU + DOF + DOF + NX + RB_DOUBLES + LIN_SIZE + 1;
    if ((size_t)n * per > (size_t)h->B * h->S * STAGE_WS) return fail(MPCC_ERR_INVALID, "n too large for the workspace");
    double* p = h->d_ws;
    double *dx = p; p += (size_t)n * NX;
    double *du = p; p += (size_t)n * NU;
    double *dup = p; p += (size_t)n * DOF;
    double *dun = p; p += (size_t)n * DOF;
    double *dxn = p; p += (size_t)n * NX;
    double *drb = p; p += (size_t)n * RB_DOUBLES;
    double *dout = p; p += (size_t)n * LIN_SIZE;
    int32_t* dk = (int32_t*)p;
    CK(cudaMemcpyAsync(dx, x, (size_t)n * NX * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(du, u, (size_t)n * NU * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dup, u_prev, (size_t)n * DOF * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dun, u_next, (size_t)n * DOF * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dxn, x_next, (size_t)n * NX * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(drb, rb, (size_t)n * RB_DOUBLES * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dk, k, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    k_eval_stage<<<(n + 63) / 64, 64, 0, h->stream>>>(h->d_params, h->d_tracks, h->cfg.Ts, h->N, dx, du, dup, dun, dxn, drb, dk, n, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(lin_out, dout, (size_t)n * LIN_SIZE * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_eval_track(mpcc_cuda_handle* h, const double* s, int32_t n, double* out) {
    if (!h || !s || !out) return fail(MPCC_ERR_INVALID, "null argument");
    if (!h->have_track) return fail(MPCC_ERR_STATE, "track not set");
    if (n < 1 || (size_t)n * 22 > (size_t)h->B * h->S * STAGE_WS) return fail(MPCC_ERR_INVALID, "bad n");
    CK(cudaSetDevice(h->cfg.device));
    double* ds = h->d_ws;
    double* dout = h->d_ws + n;
    CK(cudaMemcpyAsync(ds, s, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    k_eval_track<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_tracks, ds, n, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, (size_t)n * 21 * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_solve_ocp(mpcc_cuda_handle* h, double* guess, const double* rb, const double* cur_u, int32_t n, int32_t* status, int32_t* iters,
                        double* steps, double* alphas, int32_t max_log, int32_t* n_logged) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!guess || !rb || !cur_u || !status || !iters) return fail(MPCC_ERR_INVALID, "null argument");
    if (n < 1 || n > h->B) return fail(MPCC_ERR_INVALID, "n must be in [1, batch]");
    if (max_log < 0) max_log = 0;
    if (max_log > 0 && (!alphas || !n_logged)) return fail(MPCC_ERR_INVALID, "log arrays missing");
    CK(cudaSetDevice(h->cfg.device));
    const size_t HN = (size_t)h->S * HZ;
    // buffers carved from the handle's grow-only arena: no allocation (and no implicit device synchronisation) per call
    const int ml = max_log > 0 ? max_log : 1;
    const bool want_steps = steps && max_log > 0;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_g = up(n * HN * 8), b_rb = up((size_t)n * h->S * RB_DOUBLES * 8), b_cu = up((size_t)n * NU * 8), b_steps = want_steps ? up((size_t)n * ml * HN * 8) : 0,
                 b_al = up((size_t)n * ml * 8), b_ok = up((size_t)n * ml * 4), b_nl = up((size_t)n * 4);
    CK(h->need_scratch(b_g + b_rb + b_cu + b_steps + b_al + b_ok + b_nl));
    char* sp = h->scratch;
    double* d_g = (double*)sp; sp += b_g;
    double* d_rb = (double*)sp; sp += b_rb;
    double* d_cu = (double*)sp; sp += b_cu;
    double* d_steps = want_steps ? (double*)sp : nullptr; sp += b_steps;
    double* d_alphas = (double*)sp; sp += b_al;
    int32_t* d_qpok = (int32_t*)sp; sp += b_ok;
    int32_t* d_nl = (int32_t*)sp;
    CK(cudaMemcpyAsync(d_g, guess, n * HN * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_rb, rb, (size_t)n * h->S * RB_DOUBLES * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_cu, cur_u, (size_t)n * NU * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(d_nl, 0, (size_t)n * 4, h->stream));
    CycleArgs a = make_args(h, h->d_x0, h->d_u0, h->d_obs_dummy);
    if (h->soc) launch_solve_ocp_soc(a, h->d_wws, h->use_cta, d_g, d_rb, d_cu, n, d_steps, d_alphas, d_qpok, max_log, d_nl, h->stream);
    else if (h->use_cta) launch_solve_ocp_cta(a, h->d_wws, d_g, d_rb, d_cu, n, d_steps, d_alphas, d_qpok, max_log, d_nl, h->stream);
    else launch_solve_ocp_warp(a, h->d_wws, d_g, d_rb, d_cu, n, d_steps, d_alphas, d_qpok, max_log, d_nl, h->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(guess, d_g, n * HN * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(status, h->d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(iters, h->d_iters, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (max_log > 0) {
        if (steps) CK(cudaMemcpyAsync(steps, d_steps, (size_t)n * ml * HN * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(alphas, d_alphas, (size_t)n * ml * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(n_logged, d_nl, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_sim_time_step_device(mpcc_cuda_handle* h, const double* d_x, const double* d_u, double ts, double* d_x_next) {
    if (!h || !d_x || !d_u || !d_x_next) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    k_sim_step<<<(h->B + 127) / 128, 128, 0, h->stream>>>(d_x, d_u, ts, h->B, d_x_next);
    CK(cudaGetLastError());
    return MPCC_OK;
}

int mpcc_cuda_set_profiling(mpcc_cuda_handle* h, int32_t on) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    h->profiling = on != 0;
    return MPCC_OK;
}
int mpcc_cuda_get_kernel_times(mpcc_cuda_handle* h, double* ms4) {
    if (!h || !ms4) return fail(MPCC_ERR_INVALID, "null argument");
    if (!h->profiling) return fail(MPCC_ERR_STATE, "profiling is off (mpcc_cuda_set_profiling)");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaEventSynchronize(h->ev[4]));
    for (int i = 0; i < 4; i++) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
        ms4[i] = t;
    }
    return MPCC_OK;
}

// FP64 peak of the device (the larger of the DFMA and the DMMA figure; one shared pipe): 8 independent FMA chains per thread, enough CTAs to fill every SM
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}
// the same pipe driven by the tensor instruction the MLP kernel uses (mma.sync.m8n8k4.f64): 8 independent accumulator pairs
__global__ void __launch_bounds__(256) k_fp64_peak_mma(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i * 1e-3;
    const double x = 0.5 + threadIdx.x * 1e-6, y = 0.25;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) dmma884(c[2 * j], c[2 * j + 1], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}
int mpcc_cuda_fp64_peak(int32_t device, double* tflops) {
    if (!tflops) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    k_fp64_peak<<<grid, 256>>>(d, 64);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k_fp64_peak<<<grid, 256>>>(d, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 64.0 * iters * 256.0 * grid;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        // tensor instruction: 64 mma per thread and iteration, 256 MAC per warp-wide mma = 8 per lane
        CK(cudaEventRecord(e0));
        k_fp64_peak_mma<<<grid, 256>>>(d, iters / 8);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        fl = 2.0 * 64.0 * 8.0 * (iters / 8) * 256.0 * grid;
        tf = fl / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return MPCC_OK;
}

int mpcc_cuda_read_compute_time(mpcc_cuda_handle* h, double* seconds) {
    if (!h || !seconds) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    std::vector<long long> ns((size_t)4 * h->B);
    CK(cudaMemcpyAsync(ns.data(), h->d_sqp_ns, ns.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < ns.size(); i++) seconds[i] = 1e-9 * (double)ns[i];
    return MPCC_OK;
}

int mpcc_cuda_read_qp_counters(mpcc_cuda_handle* h, int32_t* qp_iters, int32_t* qp_fail) {
    if (!h || !qp_iters || !qp_fail) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(qp_iters, h->d_qp_iters, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(qp_fail, h->d_qp_fail, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

int mpcc_cuda_read_decisions(mpcc_cuda_handle* h, int32_t* accept_mask) {
    if (!h || !accept_mask) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(accept_mask, h->d_accept, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCC_OK;
}

// ---- multi-GPU: one process per GPU, one handle per process; the batch is sharded by the caller (instances are independent),
// the only exchange is this gather of the per-instance results ----
int mpcc_cuda_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(MPCC_ERR_INVALID, "null argument");
    std::string err;
    if (!g_nccl.load(err)) return fail(MPCC_ERR_STATE, err);
    mpcc_ncclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(MPCC_ERR_CUDA, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    std::memcpy(id128, id.internal, 128);
    return MPCC_OK;
}

int mpcc_cuda_comm_init(mpcc_cuda_handle* h, const uint8_t* id128, int32_t rank, int32_t world) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    if (world < 1 || rank < 0 || rank >= world) return fail(MPCC_ERR_INVALID, "rank / world out of range");
    if (h->world != 0) return fail(MPCC_ERR_STATE, "communicator already initialised");
    if (world > 1 && !id128) return fail(MPCC_ERR_INVALID, "unique id missing");
    CK(cudaSetDevice(h->cfg.device));
    if (world > 1) {
        std::string err;
        if (!g_nccl.load(err)) return fail(MPCC_ERR_STATE, err);
        mpcc_ncclUniqueId id;
        std::memcpy(id.internal, id128, 128);
        const int rc = g_nccl.CommInitRank(&h->comm, world, id, rank);
        if (rc != 0) return fail(MPCC_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
    }
    const size_t rec = (size_t)h->B * 9;
    CK(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CK(cudaEventCreateWithFlags(&h->ev_pack[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_gath[i], cudaEventDisableTiming));
        CK(cudaMalloc((void**)&h->d_stage[i], rec * 8));
        CK(cudaMalloc((void**)&h->d_gath[i], rec * 8 * world));
    }
    CK(cudaHostAlloc((void**)&h->h_gath, rec * 8 * world, cudaHostAllocPortable));
    h->rank = rank; h->world = world;
    return MPCC_OK;
}

int mpcc_cuda_gather_results(mpcc_cuda_handle* h) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    if (h->world == 0) return fail(MPCC_ERR_STATE, "no communicator (mpcc_cuda_comm_init)");
    CK(cudaSetDevice(h->cfg.device));
    const int slot = h->gslot;
    const size_t rec = (size_t)h->B * 9;
    // the gather that read this staging slot two cycles ago must be through before the slot is rewritten (normally long done)
    if (h->gused[slot]) CK(cudaStreamWaitEvent(h->stream, h->ev_gath[slot], 0));
    k_pack_results<<<(unsigned)((rec + 255) / 256), 256, 0, h->stream>>>(h->d_u_out, h->d_status, h->d_iters, h->B, h->d_stage[slot]);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_pack[slot], h->stream));
    CK(cudaStreamWaitEvent(h->gstream, h->ev_pack[slot], 0));
    if (h->world > 1) {
        const int rc = g_nccl.AllGather(h->d_stage[slot], h->d_gath[slot], rec, 8 /* ncclDouble */, h->comm, h->gstream);
        if (rc != 0) return fail(MPCC_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc));
    } else {
        CK(cudaMemcpyAsync(h->d_gath[slot], h->d_stage[slot], rec * 8, cudaMemcpyDeviceToDevice, h->gstream));
    }
    CK(cudaEventRecord(h->ev_gath[slot], h->gstream));
    h->gused[slot] = true; h->glast = slot; h->gslot ^= 1;
    h->launches++;
    return MPCC_OK;
}

int mpcc_cuda_read_gathered(mpcc_cuda_handle* h, double* u_all, int32_t* status_all, int32_t* iters_all) {
    if (!h) return fail(MPCC_ERR_INVALID, "null handle");
    if (h->glast < 0) return fail(MPCC_ERR_STATE, "nothing gathered yet (mpcc_cuda_gather_results)");
    CK(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->B * h->world;
    CK(cudaMemcpyAsync(h->h_gath, h->d_gath[h->glast], n * 9 * 8, cudaMemcpyDeviceToHost, h->gstream));  // ordered behind the gather
    CK(cudaStreamSynchronize(h->gstream));
    for (size_t i = 0; i < n; i++) {
        const double* r = h->h_gath + i * 9;
        if (u_all) std::memcpy(u_all + i * 8, r, 64);
        int32_t v[2];
        std::memcpy(v, r + 8, 8);
        if (status_all) status_all[i] = v[0];
        if (iters_all) iters_all[i] = v[1];
    }
    return MPCC_OK;
}

int mpcc_cuda_gathered_pointer(mpcc_cuda_handle* h, double** d_packed, void** stream) {
    if (!h || !d_packed) return fail(MPCC_ERR_INVALID, "null argument");
    if (h->glast < 0) return fail(MPCC_ERR_STATE, "nothing gathered yet (mpcc_cuda_gather_results)");
    *d_packed = h->d_gath[h->glast];
    if (stream) *stream = (void*)h->gstream;
    return MPCC_OK;
}

int64_t mpcc_cuda_launch_count(mpcc_cuda_handle* h) { return h ? h->launches : 0; }

int mpcc_cuda_get_stats(mpcc_cuda_handle* h, int64_t* st) {
    if (!h || !st) return fail(MPCC_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const size_t B = h->B;
    std::vector<int32_t> status(B), iters(B), ok(B), qi(B), qf(B);
    CK(cudaMemcpyAsync(status.data(), h->d_status, B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(iters.data(), h->d_iters, B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(ok.data(), h->d_ok, B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(qi.data(), h->d_qp_iters, B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(qf.data(), h->d_qp_fail, B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    st[0] = h->launches; st[1] = st[2] = st[3] = st[4] = st[5] = 0;
    for (size_t b = 0; b < B; b++) { st[1] += iters[b]; st[2] += qi[b]; st[3] += qf[b]; st[4] += (status[b] == SOLVED); st[5] += ok[b]; }
    return MPCC_OK;
}

}  // extern "C"
