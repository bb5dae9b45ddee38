// Warp-cooperative SQP loop: ONE WARP PER MPCC INSTANCE.
//
// Same algorithm as the one-thread-per-instance formulation in dev_sqp.cuh / dev_qp.cuh:
//   reference SQP loop           cpp/src/Interfaces/osqp_interface.cpp:398-590
//   filter line search           :759-808
//   QP (replaces OSQP, :592-656) Mehrotra predictor-corrector interior point; every Newton system solved by a
//                                Riccati recursion on the state augmented with the previous joint-velocity step.
//
// The code is written as a sequence of PHASES: inside a phase every lane works on its own items and only reads
// data written in earlier phases; phases are separated by a warp barrier.  On the device a phase is one call
// of its body with lane = threadIdx.x & 31 (reductions by warp shuffles); compiled for the host (tests/emul) a
// phase is a loop over the 32 lanes, forwards or backwards, which is how the CPU-only test tier exercises this
// exact code (and checks that no phase depends on the lane order).
//
// Work distribution:
//   * stage linearisation / trial evaluation: lane = stage (stage_eval is independent per stage);
//   * per-constraint passes of the interior-point iteration: tiles of the flat per-constraint vectors streamed through a
//     shared-memory ring by asynchronous copies, one generic tile loop for all passes (stream_constraints);
//   * Riccati factorisation and the two vector sweeps: sequential in the stage; inside a stage the 32 lanes work
//     on the 8x8 / 8x16 / 16x16 / 14x14 blocks in this warp's shared-memory scratch (the two dense products on
//     mma.m8n8k4.f64, the 8x8 Cholesky replicated in registers); the factor is stored as L^-1 and Lam = L^-1 Mnx so
//     that both sweeps are pure mat-vecs.
// Code size is a first-order concern: with ten warps per SM in different phases the interior-point iteration was bound
// by instruction fetch until its code was made compact (one copy of every pass, rolled lane loops).
// Memory: the iterate, the QP point and the Newton step live in shared memory; the per-stage QP blocks, the
// per-constraint interior-point vectors and the Riccati factors live in a per-instance CONTIGUOUS global
// workspace (every warp access is a run of consecutive doubles).
//
// The dynamics multipliers are not iterated: the dual residual is evaluated with the costates of the
// xi-stationarity recursion  p_N = g_N,  p_k = g_k + A' p_{k+1}  (g = gradient of the Lagrangian in xi with the
// current inequality multipliers; p = -y of dev_qp.cuh), which zeroes the xi-residual by construction; the
// termination test is on the nu-stationarity residual g_nu,k + B' p_{k+1}, the primal residual and the
// complementarity gap, with the thresholds of dev_qp.cuh.
#pragma once
#include "dev_sqp.cuh"

namespace mpcc {

// ---- the 32-lane execution abstraction -----------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
// butterfly reductions, one copy each in the kernel image (the QP path calls them ~10 times per interior-point iteration)
static __device__ __noinline__ double warp_all_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
static __device__ __noinline__ double warp_all_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
static __device__ __noinline__ double warp_all_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif
// NL lanes working in lock step on one instance: a warp (NL = 32, the throughput kernel: one warp per instance, many
// instances per SM) or a whole CTA (NL = 128, the latency kernel: one CTA per instance, for batches that do not fill the
// machine).  Phases end with the group's barrier (__syncwarp / __syncthreads).
template <int NL>
struct Lanes {
    static_assert(NL % 32 == 0 && NL >= 32, "whole warps");
#if defined(__CUDA_ARCH__)
    int lane;
    double* red;  // NL > 32: 2 * (NL / 32) doubles of shared memory for the cross-warp step of the reductions
    __device__ __forceinline__ void sync() const { if (NL == 32) __syncwarp(); else __syncthreads(); }
    template <class Op> __device__ __forceinline__ double across(double v, Op op) const {
        if (NL == 32) { __syncwarp(); return v; }
        if ((lane & 31) == 0) red[lane >> 5] = v;
        __syncthreads();
        double r = red[0];
#pragma unroll
        for (int w = 1; w < NL / 32; w++) r = op(r, red[w]);
        __syncthreads();
        return r;
    }
    template <class F> __device__ __forceinline__ void each(F f) const { f(lane); sync(); }
    template <class F> __device__ __forceinline__ double rmax(F f) const { return across(warp_all_max(f(lane)), [](double a, double b) { return fmax(a, b); }); }
    template <class F> __device__ __forceinline__ double rmin(F f) const { return across(warp_all_min(f(lane)), [](double a, double b) { return fmin(a, b); }); }
    template <class F> __device__ __forceinline__ double rsum(F f) const { return across(warp_all_sum(f(lane)), [](double a, double b) { return a + b; }); }
    template <class F> __device__ __forceinline__ bool any(F f) const {
        if (NL == 32) { const bool r = __any_sync(0xffffffffu, f(lane)); __syncwarp(); return r; }
        return __syncthreads_or(f(lane) ? 1 : 0) != 0;
    }
#else
    bool reverse = false;
    template <class F> void each(F f) const {
        if (!reverse) for (int l = 0; l < NL; l++) f(l);
        else for (int l = NL - 1; l >= 0; l--) f(l);
    }
    template <class F> double rmax(F f) const { double v = -INFINITY; each([&](int l) { v = fmax(v, f(l)); }); return v; }
    template <class F> double rmin(F f) const { double v = INFINITY; each([&](int l) { v = fmin(v, f(l)); }); return v; }
    template <class F> double rsum(F f) const { double v = 0; each([&](int l) { v += f(l); }); return v; }
    template <class F> bool any(F f) const { bool v = false; each([&](int l) { v = f(l) || v; }); return v; }
#endif
};
using Warp = Lanes<32>;

// The interior-point iteration is bound by instruction fetch, not issue (ncu: stall_no_instruction dominates with ten
// warps per SM in different phases of ~100 KB of code): lane-strided loops of the QP path stay rolled.
#if defined(__CUDACC__) && !defined(MPCC_NO_ROLL)
#define MPCC_ROLLED _Pragma("unroll 1")
#else
#define MPCC_ROLLED
#endif

#if defined(__CUDA_ARCH__)
#define MPCC_RSQRT(x) rsqrt(x)
#else
#define MPCC_RSQRT(x) (1.0 / sqrt(x))
#endif

// 8-byte asynchronous global -> shared copy (LDGSTS): the data does not pass through registers, so a lane can keep
// many copies in flight; groups complete in commit order.  Host build: plain copies.
MPCC_HD void async_copy8(double* smem_dst, const double* gsrc) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
#else
    *smem_dst = *gsrc;
#endif
}
#if defined(__CUDA_ARCH__)
// D (8x8) += A (8x4, row) * B (4x8, col) on the FP64 tensor path.  Lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2], D[l >> 2][2 (l & 3) + {0, 1}].
__device__ __forceinline__ void warp_dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif
// 16-byte variant: both addresses 16-byte aligned; copies smem_dst[0..1]
MPCC_HD void async_copy16(double* smem_dst, const double* gsrc) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
#else
    smem_dst[0] = gsrc[0]; smem_dst[1] = gsrc[1];
#endif
}
MPCC_HD void async_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
template <int PENDING>
MPCC_HD void async_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory");
#endif
}

// ---- layouts -------------------------------------------------------------------------------------------------
// per-stage QP record written by the linearisation (polytopic rows live in the cycle-constant record)
constexpr int WL_Q = 0, WL_q = 81, WL_RD = 90, WL_r = 98, WL_b = 106, WL_XLO = 115, WL_XHI = 124, WL_DLO = 133, WL_DHI = 140, WL_PRHS = 147, WL_SIZE = 158;
constexpr int WC_SIZE = NPOLY * 14;   // cycle constants per stage: 11 normalised polytopic rows [ax(7) | au(7)]
constexpr int WF_X = 0, WF_LAM = 64, WF_SIZE = 192;  // factor record: L^-1 (8 x 8, lower), Lam (8 x 16)
// shared-memory scratch of one warp (doubles)
constexpr int SC_P = 0, SC_PM = 256, SC_MNN = 512, SC_MNX = 576, SC_X = 704, SC_LAM = 768, SC_U = 896, SC_STG = 1092, SC_FF = 1668,
              SC_VEC = 1740, SC_TXU = 1836, SC_DYN = 1854, SC_RED = 1872, SC_SIZE = 2032;  // RED: 5 x 32 per-lane accumulators
// staged inputs of one stage of the factorisation (two slots, filled by asynchronous copies one stage ahead):
// polytopic rows, their barrier weights, Q, Rd, box / rate barrier weights
// (every block but WP starts on an even double in both the source and the slot, so it travels in 16-byte pairs)
constexpr int SG_GS = 0, SG_Q = 154, SG_RD = 236, SG_WB = 244, SG_WR = 262, SG_WP = 276, SG_SIZE = 288;
constexpr int V_P = 0, V_MN = 16, V_MX = 24, V_KAP = 40, V_D0 = 48, V_D1 = 64, V_RHS = 80, V_DN = 88;  // inside SC_VEC (96)
constexpr int MAX_SQP_FILTER = 128;

// Everything the 16-byte asynchronous copies touch starts on an even double: the per-constraint vectors have an even stride
// and the per-instance workspace / per-warp shared block an even size (WL_SIZE, WC_SIZE, the tile bases are even already).
MPCC_HD size_t cvec_stride(int N) { return ((size_t)(N + 1) * NINEQ + 1) & ~(size_t)1; }
MPCC_HD size_t warp_ws_doubles(int N) {
    const size_t S = N + 1;
    const size_t n = S * (WL_SIZE + WC_SIZE + HZ /*G*/ + 8 /*KAP*/ + WF_SIZE + 2 * HZ /*persistent step, iterate*/ + 2 * NPOLY /*barrier values*/) + 8 * cvec_stride(N) + 2 * (MAX_SQP_FILTER + 2);
    return (n + 1) & ~(size_t)1;
}
// the sweeps keep the gradient (17 S) and kappa (8 S) in the scratch below SC_VEC behind a 4-slot factor ring; longer
// horizons get a separate block appended after the scratch
constexpr int SW_RING = 4, SW_GK = SW_RING * WF_SIZE;
constexpr int XG_ROOM = SC_U;  // the iterate's working copy sits in the factorisation's block area between QP solves
MPCC_HD size_t warp_smem_extra(int N) {
    return ((25 * (N + 1) <= SC_VEC - SW_GK) ? 0 : (size_t)25 * (N + 1)) + ((HZ * (N + 1) <= XG_ROOM) ? 0 : (size_t)HZ * (N + 1));
}
MPCC_HD size_t warp_smem_doubles(int N) { return ((size_t)2 * (N + 1) * HZ + SC_SIZE + warp_smem_extra(N) + 1) & ~(size_t)1; }
// a group of NL > 32 lanes has its own tile ring (three slots of 19 NL doubles: five vectors of NL polytopic rows and their
// 14 NL coefficients), 5 NL per-lane accumulators and the cross-warp reduction cells behind the warp layout
template <int NL>
MPCC_HD size_t group_smem_doubles(int N) { return warp_smem_doubles(N) + (NL > 32 ? (size_t)(3 * 19 * NL + 5 * NL + 2 * (NL / 32) + 2) : (size_t)0); }

// isPosdef / isNan of one packed-lower 9 x 9 Hessian block, fully unrolled (static indices: registers).
// pd is cleared at the first non-positive pivot unless that pivot is NaN (NaN is reported through `nan`).
MPCC_HD void block9_pd_nan(const double* Qp, bool& pd, bool& nan) {
    double A[45];
#pragma unroll
    for (int e = 0; e < 45; e++) { A[e] = Qp[e]; if (A[e] != A[e]) nan = true; }
    bool stop = false;
#pragma unroll
    for (int j = 0; j < 9; j++) {
        double d = A[sym9(j, j)];
#pragma unroll
        for (int t = 0; t < 9; t++) if (t < j) d -= A[sym9(j, t)] * A[sym9(j, t)];
        if (!stop && !(d > 0.0)) { stop = true; if (d == d) pd = false; }
        const double isd = 1.0 / sqrt(d);  // one division per column (the verdict only needs the pivots' signs)
#pragma unroll
        for (int i = 0; i < 9; i++)
            if (i > j) {
                double v = A[sym9(i, j)];
#pragma unroll
                for (int t = 0; t < 9; t++) if (t < j) v -= A[sym9(i, t)] * A[sym9(j, t)];
                A[sym9(i, j)] = v * isd;
            }
    }
}

// SOC: compile the reference's optional second-order correction (sqp.json "do_SOC", osqp_interface.cpp:506-533,658-681) into the loop.  It is a
// template flag, not a run-time branch, so that the default kernels stay the exact code they were tuned as (the interior-point path is bound by
// its instruction footprint); the SOC kernels live in their own translation unit (k_sqp_soc.cu) and are launched when a parameter set asks for it.
template <int NL, bool SOC = false>
struct GroupSqp {
    static constexpr int TS1 = 3 * NL, TS2 = NL;   // tile sizes of the streamed per-constraint passes (box + rate rows | polytopic rows)
    const Params& P;
    const TrackTable& T;
    DynConst dyn;
    double Ts;
    int N, S;
    QpOptions opt;
    Lanes<NL> W;
    // per-instance global workspace
    double *LIN, *CST, *IT, *ILAM, *IRP, *IW, *IV, *IDT, *IDLAM, *IH, *G, *KAP, *FACT, *SSTEP, *GUESS, *RBFV, *FILT;
    // per-instance shared memory
    double *VAR, *STEP, *SC;
    double* GRP;   // NL > 32 only: the group's own tile ring | per-lane accumulators | reduction cells (behind the warp layout)
    // the warp instantiation keeps both in the (idle) factorisation scratch: no extra pointers live in registers
    MPCC_HD double* ring_base() const { return NL == 32 ? SC : GRP; }
    MPCC_HD double* red_base() const { return NL == 32 ? SC + SC_RED : GRP + 3 * 19 * NL; }
    // working copies between QP solves: the iterate (in the scratch) and the persistent step (in STEP); their homes
    // GUESS / SSTEP in the global workspace are written before and re-read after every executed QP solve
    double *XG, *XS;
    int OR_, OP_;  // offsets of the rate / polytopic sub-arrays inside the per-constraint vectors
    // the reference's ComputeTime phases (osqp_interface.h:71-79) for this instance, in ns: set_qp, solve_qp, get_alpha
    double tm_set_qp = 0, tm_solve_qp = 0, tm_get_alpha = 0;
    MPCC_HD static double now_ns() {
#if defined(__CUDA_ARCH__)
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return (double)t;
#else
        return 0.0;
#endif
    }

    MPCC_HD void carve(double* gws, double* sm) {
        const size_t S_ = S;
        LIN = gws; gws += S_ * WL_SIZE;
        CST = gws; gws += S_ * WC_SIZE;
        const size_t VST = cvec_stride(N);
        IT = gws; gws += VST; ILAM = gws; gws += VST; IRP = gws; gws += VST; IW = gws; gws += VST;
        IV = gws; gws += VST; IDT = gws; gws += VST; IDLAM = gws; gws += VST; IH = gws; gws += VST;
        // FACT and G on even offsets (16-byte copies)
        FACT = gws; gws += S_ * WF_SIZE; G = gws; gws += S_ * HZ; KAP = gws; gws += S_ * 8; SSTEP = gws; gws += S_ * HZ; GUESS = gws; gws += S_ * HZ; RBFV = gws; gws += S_ * 2 * NPOLY; FILT = gws;
        VAR = sm; STEP = sm + S_ * HZ; SC = sm + 2 * S_ * HZ;
        GRP = sm + warp_smem_doubles(N);
#if defined(__CUDA_ARCH__)
        W.red = (NL == 32) ? nullptr : red_base() + 5 * NL;
#endif
        OR_ = 18 * S; OP_ = 32 * S;
        XG = (HZ * S <= XG_ROOM) ? SC : SC + SC_SIZE;
        XS = STEP;
    }
    MPCC_HD void init_scratch() const {
        W.each([&](int lane) {
            if (lane < NX) SC[SC_TXU + lane] = P.Tx[lane];
            else if (lane < HZ) SC[SC_TXU + lane] = P.Tu[lane - NX];
            if (lane == 0) { SC[SC_DYN] = dyn.asv; SC[SC_DYN + 1] = dyn.bs; SC[SC_DYN + 2] = dyn.bv; }
            if (lane < DOF) { SC[SC_DYN + 3 + lane] = dyn.bq[lane]; SC[SC_DYN + 10 + lane] = dyn.cpl[lane]; }
        });
    }
    MPCC_HD double Tx(int m) const { return SC[SC_TXU + m]; }
    MPCC_HD double Tu(int j) const { return SC[SC_TXU + NX + j]; }
    // normalised dynamics constants, read from shared memory (lane-dependent indices would otherwise put the struct in local memory)
    MPCC_HD double d_asv() const { return SC[SC_DYN]; }
    MPCC_HD double d_bs() const { return SC[SC_DYN + 1]; }
    MPCC_HD double d_bv() const { return SC[SC_DYN + 2]; }
    MPCC_HD double d_bq(int j) const { return SC[SC_DYN + 3 + j]; }
    MPCC_HD double d_cpl(int j) const { return SC[SC_DYN + 10 + j]; }

    // G z of one constraint on a [S][17] (xi | nu) vector in shared memory
    MPCC_HD double gz_box(const double* Z, int k, int c) const { return (c < 9) ? -Z[k * HZ + c] : Z[k * HZ + c - 9]; }
    MPCC_HD double gz_rate(const double* Z, int k, int c) const {
        const int j = (c < 7) ? c : c - 7;
        double d = Z[k * HZ + NX + j];
        if (k >= 1) d -= Z[(k - 1) * HZ + NX + j];
        return (c < 7) ? -d : d;
    }
    MPCC_HD double gz_poly(const double* Z, int k, int j) const {
        const double* row = CST + ((size_t)k * NPOLY + j) * 14;
        const double* z = Z + k * HZ;
        double s = 0;
#pragma unroll
        for (int m = 0; m < DOF; m++) s += row[m] * z[m] + row[7 + m] * z[NX + m];
        return s;
    }
    MPCC_HD double h_box(int k, int c) const { const double* L = LIN + (size_t)k * WL_SIZE; return (c < 9) ? -L[WL_XLO + c] : L[WL_XHI + c - 9]; }
    MPCC_HD double h_rate(int k, int c) const { const double* L = LIN + (size_t)k * WL_SIZE; return (c < 7) ? -L[WL_DLO + c] : L[WL_DHI + c - 7]; }
    MPCC_HD double h_poly(int k, int j) const { return LIN[(size_t)k * WL_SIZE + WL_PRHS + j]; }

    // f(i, gz, h) for every PRESENT constraint (i = index into the per-constraint vectors), on vector Z
    template <class F>
    MPCC_HD void for_present(int lane, const double* Z, bool need_h, F f) const {
        for (int i = 18 + lane; i < 18 * S; i += NL) { const int k = i / 18, c = i - k * 18; f(i, gz_box(Z, k, c), need_h ? h_box(k, c) : 0.0); }
        MPCC_ROLLED
        for (int i = lane; i < 14 * N; i += NL) { const int k = i / 14, c = i - k * 14; f(OR_ + i, gz_rate(Z, k, c), need_h ? h_rate(k, c) : 0.0); }
        MPCC_ROLLED
        for (int i = lane; i < NPOLY * N; i += NL) { const int k = i / NPOLY, j = i - k * NPOLY; f(OP_ + i, gz_poly(Z, k, j), need_h ? h_poly(k, j) : 0.0); }
    }

    // Same traversal, four rounds at a time: f4(idx[4], gz[4], h[4]) gets four items of one lane (idx < 0: none) so that it
    // can issue all of its loads before the first dependent use (memory-level parallelism inside a lane).
    template <class F4>
    MPCC_HD void for_present4(int lane, const double* Z, bool need_h, F4 f4) const {
        int idx[4]; double g[4], h[4];
        for (int base = 18 + lane; base < 18 * S; base += 4 * NL) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + NL * u;
                idx[u] = (i < 18 * S) ? i : -1; g[u] = 0; h[u] = 0;
                if (i < 18 * S) { const int k = i / 18, c = i - k * 18; g[u] = gz_box(Z, k, c); if (need_h) h[u] = h_box(k, c); }
            }
            f4(idx, g, h);
        }
        MPCC_ROLLED
        for (int base = lane; base < 14 * N; base += 4 * NL) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + NL * u;
                idx[u] = (i < 14 * N) ? OR_ + i : -1; g[u] = 0; h[u] = 0;
                if (i < 14 * N) { const int k = i / 14, c = i - k * 14; g[u] = gz_rate(Z, k, c); if (need_h) h[u] = h_rate(k, c); }
            }
            f4(idx, g, h);
        }
        MPCC_ROLLED
        for (int base = lane; base < NPOLY * N; base += 4 * NL) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + NL * u;
                idx[u] = (i < NPOLY * N) ? OP_ + i : -1; g[u] = 0; h[u] = 0;
                if (i < NPOLY * N) { const int k = i / NPOLY, j = i - k * NPOLY; g[u] = gz_poly(Z, k, j); if (need_h) h[u] = h_poly(k, j); }
            }
            f4(idx, g, h);
        }
    }

    // ---- streamed per-constraint passes ------------------------------------------------------------------------------
    // The per-constraint vectors of an instance (t, lam, rp, ... : 43 S doubles each, global) are consumed in flat order.
    // Instead of every lane waiting for its own loads round after round, tiles of consecutive items travel through a 3-slot
    // ring in the (idle) factorisation scratch by asynchronous copies issued two tiles ahead, so that only the first tile of a
    // pass exposes memory latency.  The PRESENT constraints are two flat ranges: box and rate rows, adjacent in the vectors
    // ([18, 18 S + 14 N), tiles of 96), and the polytopic rows ([OP_, OP_ + 11 N), tiles of 32 with their coefficient rows).
    // body(lane, i, kind, k, c, v, j, row): constraint i (kind 0 box / 1 rate / 2 polytopic, stage k, row c), its staged
    // inputs v[a * ts + j] (ts = 96 or 32 by kind) for the a-th vector of `ids`, and for polytopic rows the staged
    // coefficient row (14 doubles).  `ids` packs the vectors' indices (IT = 0, ILAM, IRP, IW, IV, IDT, IDLAM, IH = 7), 3 bits each.
    // One generic tile loop and ONE out-of-line copy routine serve all passes: the interior-point iteration is bound by
    // instruction fetch (ncu: stall_no_instruction), so its code footprint matters more than a few address computations.
    enum : unsigned { CV_T = 0, CV_LAM, CV_RP, CV_W, CV_V, CV_DT, CV_DLAM, CV_H };  // order of carve()
    MPCC_HD static unsigned vec_ids(unsigned a, unsigned b = 0, unsigned c = 0, unsigned d = 0, unsigned e = 0) { return a | (b << 3) | (c << 6) | (d << 9) | (e << 12); }
    struct TileGeom { int base, cnt, ts, rows; };
    MPCC_HD TileGeom tile_geom(int t, int nt1) const {
        TileGeom g;
        if (t < nt1) { g.base = 18 + TS1 * t; g.ts = TS1; g.rows = 0; const int end = OR_ + 14 * N; g.cnt = (end - g.base < TS1) ? end - g.base : TS1; }
        else { g.base = OP_ + TS2 * (t - nt1); g.ts = TS2; g.rows = 1; const int end = OP_ + NPOLY * N; g.cnt = (end - g.base < TS2) ? end - g.base : TS2; }
        return g;
    }
    static MPCC_HDNI void issue_tile(int lane, const double* vec0, size_t vec_stride, const double* cst, int op, double* dst, unsigned ids, int na,
                                     int base, int cnt, int ts, int rows) {
        // pairs of doubles (an odd polytopic tail copies one in-bounds double more than needed); a tile holds at most 96
        // constraints = 48 pairs per vector (two predicated copies per lane, no loop: ~6 instructions per copy instead of
        // ~18 per loop round) and 32 x 14 coefficient doubles = 224 pairs (seven)
        const int half = (cnt + 1) >> 1;
        const bool c0 = lane < half, c1 = lane + NL < half;
        MPCC_ROLLED
        for (int a = 0; a < na; a++) {
            const double* src = vec0 + (size_t)((ids >> (3 * a)) & 7u) * vec_stride + base + 2 * lane;
            double* d = dst + a * ts + 2 * lane;
            if (c0) async_copy16(d, src);
            if (c1) async_copy16(d + 2 * NL, src + 2 * NL);
        }
        if (rows) {
            const double* src = cst + (size_t)(base - op) * 14 + 2 * lane;
            double* d = dst + na * ts + 2 * lane;
            const int n2 = cnt * 7 - lane;
#pragma unroll
            for (int r = 0; r < 7; r++) if (NL * r < n2) async_copy16(d + 2 * NL * r, src + 2 * NL * r);
        }
    }
    template <class F>
    MPCC_HD void stream_constraints(unsigned ids, int na, F body) const {
        const int nt1 = (OR_ + 14 * N - 18 + TS1 - 1) / TS1, nt = nt1 + (NPOLY * N + TS2 - 1) / TS2;
        const int s1 = na * TS1, s2 = na * TS2 + TS2 * 14, slot = (s1 > s2) ? s1 : s2;  // NL = 32: <= 608 for na <= 5, three slots stay below SC_TXU
        double* ring = ring_base();
        const size_t vstride = cvec_stride(N);
        auto issue = [&](int lane, int t) {
            if (t < nt) {
                const TileGeom g = tile_geom(t, nt1);
                issue_tile(lane, IT, vstride, CST, OP_, ring + (t % 3) * slot, ids, na, g.base, g.cnt, g.ts, g.rows);
            }
            async_commit();
        };
        W.each([&](int lane) { issue(lane, 0); issue(lane, 1); async_wait<1>(); });  // tile 0 has landed
        for (int t = 0; t < nt; t++) {
            W.each([&](int lane) {
                issue(lane, t + 2);  // into the slot consumed in the previous phase
                const double* v = ring + (t % 3) * slot;
                const TileGeom g = tile_geom(t, nt1);
                MPCC_ROLLED
                for (int j = lane; j < g.cnt; j += NL) {
                    const int i = g.base + j;
                    int kind, k, c;
                    if (g.rows) { kind = 2; k = (i - OP_) / NPOLY; c = (i - OP_) - k * NPOLY; }
                    else if (i < OR_) { kind = 0; k = i / 18; c = i - k * 18; }
                    else { kind = 1; k = (i - OR_) / 14; c = (i - OR_) - k * 14; }
                    body(lane, i, kind, k, c, v, j, g.rows ? v + na * g.ts + j * 14 : nullptr);
                }
                async_wait<1>();  // pending: tiles t+1, t+2 -> tile t+1 has landed
            });
        }
        W.each([&](int) { async_wait<0>(); });
    }
    // G z of constraint (kind, k, c) on a [S][17] vector in shared memory; polytopic rows use the staged coefficient row
    MPCC_HD double gz_of(const double* Z, int kind, int k, int c, const double* row) const {
        if (kind == 0) return gz_box(Z, k, c);
        if (kind == 1) return gz_rate(Z, k, c);
        const double* z = Z + k * HZ;
        double s = 0;
#pragma unroll
        for (int m = 0; m < DOF; m++) s += row[m] * z[m] + row[7 + m] * z[NX + m];
        return s;
    }

    // ---- gradient of the step QP: G <- H z + f + G'(IV); dst0 (shared) <- same with multipliers ILAM ----
    // cost == false: the cost terms H z + f are left out (dst0 <- G' ILAM alone: the infeasibility certificate)
    MPCC_HD void gradient(double* dst0, bool cost = true) const {
        W.each([&](int lane) {
            MPCC_ROLLED
            for (int o = lane; o < NX * S; o += NL) {
                const int k = o / NX, r = o - k * NX;
                const double* L = LIN + (size_t)k * WL_SIZE;
                const double* z = VAR + k * HZ;
                double base = 0.0;
                if (cost) {
                    base = L[WL_q + r];
#pragma unroll
                    for (int c = 0; c < NX; c++) base += L[WL_Q + r * 9 + c] * z[c];
                }
                double sv = IV[k * 18 + 9 + r] - IV[k * 18 + r];
                double sl = dst0 ? ILAM[k * 18 + 9 + r] - ILAM[k * 18 + r] : 0.0;
                if (k < N && r < DOF) {
                    const double* col = CST + (size_t)k * WC_SIZE + r;
#pragma unroll
                    for (int j = 0; j < NPOLY; j++) {
                        const double a = col[j * 14];
                        sv += IV[OP_ + k * NPOLY + j] * a;
                        if (dst0) sl += ILAM[OP_ + k * NPOLY + j] * a;
                    }
                }
                G[k * HZ + r] = base + sv;
                if (dst0) dst0[k * HZ + r] = base + sl;
            }
            MPCC_ROLLED
            for (int o = lane; o < NU * N; o += NL) {
                const int k = o / NU, j = o - k * NU;
                const double* L = LIN + (size_t)k * WL_SIZE;
                const double* z = VAR + k * HZ + NX;
                double base = cost ? L[WL_RD + j] * z[j] + L[WL_r + j] : 0.0;
                double sv = 0, sl = 0;
                if (j < DOF) {
                    if (cost && k >= 1) base += d_cpl(j) * z[j - HZ];
                    if (cost && k <= N - 2) base += d_cpl(j) * z[j + HZ];
                    const int ir = OR_ + k * 14;
                    sv = IV[ir + 7 + j] - IV[ir + j];
                    if (dst0) sl = ILAM[ir + 7 + j] - ILAM[ir + j];
                    if (k + 1 < N) {
                        sv -= IV[ir + 14 + 7 + j] - IV[ir + 14 + j];
                        if (dst0) sl -= ILAM[ir + 14 + 7 + j] - ILAM[ir + 14 + j];
                    }
                    const double* col = CST + (size_t)k * WC_SIZE + 7 + j;
#pragma unroll
                    for (int jj = 0; jj < NPOLY; jj++) {
                        const double a = col[jj * 14];
                        sv += IV[OP_ + k * NPOLY + jj] * a;
                        if (dst0) sl += ILAM[OP_ + k * NPOLY + jj] * a;
                    }
                }
                G[k * HZ + NX + j] = base + sv;
                if (dst0) dst0[k * HZ + NX + j] = base + sl;
            }
        });
    }

    // costates of the xi-stationarity recursion p_N = g_N, p_k = g_k + A' p_{k+1}, in place on the xi part of STEP (which holds g);
    // returns the inf-norm of the nu-stationarity residual g_nu,k + B' p_{k+1}
    MPCC_HD double costate_residual() const {
        W.each([&](int lane) {
            if (lane < 8) for (int k = N - 1; k >= 1; k--) STEP[k * HZ + lane] += STEP[(k + 1) * HZ + lane];
        });
        W.each([&](int lane) {
            if (lane == 8) for (int k = N - 1; k >= 1; k--) STEP[k * HZ + 8] += STEP[(k + 1) * HZ + 8] + d_asv() * STEP[(k + 1) * HZ + 7];
        });
        return W.rmax([&](int lane) {
            double nr = 0;
            MPCC_ROLLED
            for (int o = lane; o < NU * N; o += NL) {
                const int k = o / NU, j = o - k * NU;
                const double* pn = STEP + (k + 1) * HZ;
                const double btp = (j < 7) ? d_bq(j) * pn[j] : d_bs() * pn[7] + d_bv() * pn[8];
                nr = fmax(nr, fabs(STEP[k * HZ + NX + j] + btp));
            }
            return nr;
        });
    }

    // Primal infeasibility certificate (Farkas; the counterpart of OSQP's primal-infeasibility test, which makes the reference's
    // solveQP return PrimalInfeasible quickly, osqp_interface.cpp:495-497): for multipliers lam >= 0 and the costates p of
    // G' lam (recursion above), every z that satisfies the dynamics has lam' G z = r' nu + sum_k p_{k+1}' b_k with r the
    // nu-stationarity residual.  If r vanishes (relative to |lam|) and sum_k p_{k+1}' b_k - h' lam > 0, no such z has G z <= h.
    // The diverging multipliers of an interior-point run on an infeasible QP converge to such a ray.  Tested only on
    // iterations that follow a short step (the hot path never gets here); tolerance eps_inf relative to |lam|_inf.
    // Out of line and called on a COPY of this object: a noinline member call would make `this` escape, and the compiler then
    // keeps the object in local memory for the whole kernel (+500 local loads in the hot loops, steady-state kernel 7.2 instead
    // of 5.1 ms, although the call never executes in a healthy batch).
    static MPCC_HDNI bool primal_infeasible(GroupSqp g, double eps_inf) { return g.primal_infeasible_impl(eps_inf); }
    MPCC_HD bool primal_infeasible_impl(double eps_inf) const {
        double* RED = red_base();
        W.each([&](int lane) { RED[lane] = 0.0; RED[NL + lane] = 0.0; });
        stream_constraints(vec_ids(CV_LAM, CV_H), 2, [&](int lane, int, int kind, int, int, const double* v, int j, const double*) {
            const int ts = (kind == 2) ? TS2 : TS1;
            const double lam = v[j], h = v[ts + j];
            RED[lane] = fmax(RED[lane], lam);
            RED[NL + lane] += lam * h;
        });
        const double lmax = W.rmax([&](int lane) { return RED[lane]; });
        const double hl = W.rsum([&](int lane) { return RED[NL + lane]; });
        if (!(lmax > 0.0) || !(lmax < 1e300)) return false;
        gradient(STEP, false);  // STEP <- G' lam  (G, the gradient with the predictor multipliers, is recomputed by the next pass)
        const double nr = costate_residual();
        const double c = W.rsum([&](int lane) {
            double a = 0;
            MPCC_ROLLED
            for (int o = lane; o < NX * N; o += NL) { const int k = o / NX, r = o - k * NX; a += STEP[(k + 1) * HZ + r] * LIN[(size_t)k * WL_SIZE + WL_b + r]; }
            return a;
        });
#if defined(MPCC_QP_TRACE) && !defined(__CUDA_ARCH__)
        fprintf(stderr, "          certificate: |lam| %.3e  r/|lam| %.3e  (c - h'lam)/|lam| %.3e\n", lmax, nr / lmax, (c - hl) / lmax);
#endif
        // rigorous for every z with |nu|_1 <= 1e4 (normalised input steps are O(1)): lam'(G z - h) = r'nu + c - h'lam > 0
        return (c - hl) >= eps_inf * lmax && nr * 1e4 <= (c - hl);
    }

    // asynchronous fetch of the factorisation inputs of stage k into slot k & 1: 138 pairs (polytopic rows 77, Q 41 -- the last
    // pair carries one unused double --, Rd 4, box weights 9, rate weights 7) + the 11 polytopic weights (odd source offsets)
    MPCC_HD void issue_stage_copy(int lane, int k, double* SG) const {
        if (k >= 0) {
            double* dst = SG + (k & 1) * SG_SIZE;
            const double* L = LIN + (size_t)k * WL_SIZE;
#pragma unroll
            for (int r = 0; r < (138 + NL - 1) / NL; r++) {
                const int e = lane + NL * r;
                const double* src = nullptr;
                int d = 0;
                if (e < 77) { src = CST + (size_t)k * WC_SIZE + 2 * e; d = SG_GS + 2 * e; }
                else if (e < 118) { src = L + WL_Q + 2 * (e - 77); d = SG_Q + 2 * (e - 77); }
                else if (e < 122) { src = L + WL_RD + 2 * (e - 118); d = SG_RD + 2 * (e - 118); }
                else if (e < 131) { src = IW + k * 18 + 2 * (e - 122); d = SG_WB + 2 * (e - 122); }
                else if (e < 138) { src = IW + OR_ + k * 14 + 2 * (e - 131); d = SG_WR + 2 * (e - 131); }
                if (src) async_copy16(dst + d, src);
            }
            if (lane < NPOLY) async_copy8(dst + SG_WP + lane, IW + OP_ + k * NPOLY + lane);
        }
        async_commit();
    }

    // ---- Riccati factorisation; false if some M_nunu is not positive definite ----
    MPCC_HD bool factor() const {
        double* Pc = SC + SC_P;      // 16 x 16 cost-to-go of stage k+1 (rows/cols 0..8 xi, 9..15 previous dq step)
        double* PM = SC + SC_PM;     // [Mxx 0; 0 Mww] of stage k
        double* Mnn = SC + SC_MNN;   // 8 x 8
        double* Mnx = SC + SC_MNX;   // 8 x 16
        double* X = SC + SC_X;       // L^-1
        double* Lam = SC + SC_LAM;   // L^-1 Mnx
        double* U = SC + SC_U;       // 14 x 14 polytopic barrier Hessian  sum_p w_p g_p g_p'
        double* SG = SC + SC_STG;
        double* FF = SC + SC_FF;
        // terminal stage: P_N = Q_N + box W
        W.each([&](int lane) {
            const double* L = LIN + (size_t)N * WL_SIZE;
            MPCC_ROLLED
            for (int e = lane; e < 256; e += NL) {
                const int r = e >> 4, c = e & 15;
                double v = 0;
                if (r < 9 && c < 9) {
                    v = L[WL_Q + r * 9 + c];
                    if (r == c) v += IW[N * 18 + r] + IW[N * 18 + 9 + r];
                }
                Pc[e] = v;
                PM[e] = 0.0;                 // only the structural entries of PM / Mnx are rewritten per stage
                if (e < 128) Mnx[e] = 0.0;
            }
            issue_stage_copy(lane, N - 1, SG);
            async_wait<0>();
        });
        bool ok = true;
        for (int k = N - 1; k >= 0; k--) {
            const double* SGk = SG + (k & 1) * SG_SIZE;
            const double* GS = SGk + SG_GS;
            const double* WP = SGk + SG_WP;
            const double* Qs = SGk + SG_Q;
            const double* RDs = SGk + SG_RD;
            const double* wB = SGk + SG_WB;
            const double* wR = SGk + SG_WR;
            // F1: start fetching the next stage's inputs; FF = B'Pxx + E'Pwx (8 x 9)
            W.each([&](int lane) {
                issue_stage_copy(lane, k - 1, SG);
                MPCC_ROLLED
                for (int e = lane; e < 72; e += NL) {
                    const int i = e / 9, c = e - i * 9;
                    FF[e] = (i < 7) ? d_bq(i) * Pc[i * 16 + c] + Pc[(9 + i) * 16 + c] : d_bs() * Pc[7 * 16 + c] + d_bv() * Pc[8 * 16 + c];
                }
                // F2 (same phase: independent of FF): U = sum_p (w_p g_p) g_p'
#if defined(__CUDA_ARCH__)
                if (lane < 32) {   // (warp 0 of a wider group: mma.sync is a warp-level instruction)
                    // (14 x 11)(11 x 14) padded to 16 x 12 x 16: 2 x 2 fragments x 3 k-steps of mma.m8n8k4.f64; as in F6 the A
                    // fragment of block row m is w times the B fragment of block column m
                    const int fr = lane >> 2, fq = lane & 3;
                    double gf[2][3], wf[3];
#pragma unroll
                    for (int ks = 0; ks < 3; ks++) {
                        const int p = ks * 4 + fq;
                        wf[ks] = (p < NPOLY) ? WP[p] : 0.0;
#pragma unroll
                        for (int m = 0; m < 2; m++) { const int a = m * 8 + fr; gf[m][ks] = (p < NPOLY && a < 14) ? GS[p * 14 + a] : 0.0; }
                    }
#pragma unroll
                    for (int mb = 0; mb < 2; mb++)
#pragma unroll
                        for (int nb = 0; nb < 2; nb++) {
                            double c0 = 0.0, c1 = 0.0;
#pragma unroll
                            for (int ks = 0; ks < 3; ks++) warp_dmma884(c0, c1, wf[ks] * gf[mb][ks], gf[nb][ks]);
                            const int a = mb * 8 + fr, b0 = nb * 8 + 2 * fq;
                            if (a < 14 && b0 < 14) { U[a * 14 + b0] = c0; U[a * 14 + b0 + 1] = c1; }
                        }
                }
#else
                if (lane < 28) {  // 2 x 2 register blocks, upper triangle + mirror
                    int bi = 0, t = lane;
                    while (t >= 7 - bi) { t -= 7 - bi; bi++; }
                    const int bj = bi + t;
                    double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
#pragma unroll
                    for (int p = 0; p < NPOLY; p++) {
                        const double wp = WP[p];
                        const double w0 = wp * GS[p * 14 + 2 * bi], w1 = wp * GS[p * 14 + 2 * bi + 1];
                        const double g0 = GS[p * 14 + 2 * bj], g1 = GS[p * 14 + 2 * bj + 1];
                        a00 += w0 * g0; a01 += w0 * g1; a10 += w1 * g0; a11 += w1 * g1;
                    }
                    U[(2 * bi) * 14 + 2 * bj] = a00; U[(2 * bi) * 14 + 2 * bj + 1] = a01;
                    U[(2 * bi + 1) * 14 + 2 * bj] = a10; U[(2 * bi + 1) * 14 + 2 * bj + 1] = a11;
                    if (bi != bj) {
                        U[(2 * bj) * 14 + 2 * bi] = a00; U[(2 * bj + 1) * 14 + 2 * bi] = a01;
                        U[(2 * bj) * 14 + 2 * bi + 1] = a10; U[(2 * bj + 1) * 14 + 2 * bi + 1] = a11;
                    }
                }
#endif
            });
            // F3: Mnn (8 x 8), Mnx (8 x 16) and [Mxx 0; 0 Mww] (16 x 16)
            W.each([&](int lane) {
                MPCC_ROLLED
                for (int e = lane; e < 64; e += NL) {
                    const int i = e >> 3, j = e & 7;
                    double v;
                    if (j < 7) {
                        v = d_bq(j) * FF[i * 9 + j];
                        if (i < 7) v += d_bq(i) * Pc[(9 + j) * 16 + i] + Pc[(9 + i) * 16 + 9 + j] + U[(7 + i) * 14 + 7 + j];
                        else v += d_bs() * Pc[(9 + j) * 16 + 7] + d_bv() * Pc[(9 + j) * 16 + 8];
                    } else {
                        v = d_bs() * FF[i * 9 + 7] + d_bv() * FF[i * 9 + 8];
                    }
                    if (i == j) { v += RDs[j]; if (j < 7) v += wR[j] + wR[7 + j]; }
                    Mnn[e] = v;
                }
                MPCC_ROLLED
                for (int e = lane; e < 79; e += NL) {  // structural entries of Mnx: the 8 x 9 block and the rate coupling diagonal
                    int i, c;
                    double v;
                    if (e < 72) {
                        i = e / 9; c = e - i * 9;
                        v = FF[e];
                        if (c == 8) v += d_asv() * FF[i * 9 + 7];
                        if (i < 7 && c < 7) v += U[(7 + i) * 14 + c];
                    } else {
                        i = e - 72; c = 9 + i;
                        v = (k >= 1) ? d_cpl(i) - (wR[i] + wR[7 + i]) : 0.0;
                    }
                    Mnx[i * 16 + c] = v;
                }
                MPCC_ROLLED
                for (int e = lane; e < 88; e += NL) {  // structural entries of [Mxx 0; 0 Mww]: the 9 x 9 block and 7 diagonal entries
                    if (e < 81) {
                        const int r = e / 9, c = e - r * 9;
                        double v = Pc[r * 16 + c];
                        if (c == 8) v += d_asv() * Pc[r * 16 + 7];
                        if (r == 8) v += d_asv() * (Pc[7 * 16 + c] + ((c == 8) ? d_asv() * Pc[7 * 16 + 7] : 0.0));
                        v += Qs[e];
                        if (r == c && k >= 1) v += wB[r] + wB[9 + r];
                        if (r < 7 && c < 7) v += U[r * 14 + c];
                        PM[r * 16 + c] = v;
                    } else {
                        const int j = e - 81;
                        PM[(9 + j) * 17] = (k >= 1) ? wR[j] + wR[7 + j] : 0.0;
                    }
                }
            });
            // F4 + F5: every lane c < 24 factors Mnn = Lc Lc' in its own registers (no barriers or shared-memory round trips on
            //     the pivot chain), then solves its column of [Lam | X] = Lc^-1 [Mnx | I] by forward substitution
            const bool bad = W.any([&](int lane) {
                if (lane >= 24) return false;
                double Lr[36], inv[8];
                bool pd = true;
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) if (j <= i) Lr[i * (i + 1) / 2 + j] = Mnn[i * 8 + j];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    double d = Lr[j * (j + 1) / 2 + j];
#pragma unroll
                    for (int t = 0; t < 8; t++) if (t < j) d -= Lr[j * (j + 1) / 2 + t] * Lr[j * (j + 1) / 2 + t];
                    if (!(d > 0.0 && d < 1e300)) pd = false;
                    const double iv = MPCC_RSQRT(d);
                    inv[j] = iv;
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if (i > j) {
                            double v = Lr[i * (i + 1) / 2 + j];
#pragma unroll
                            for (int t = 0; t < 8; t++) if (t < j) v -= Lr[i * (i + 1) / 2 + t] * Lr[j * (j + 1) / 2 + t];
                            Lr[i * (i + 1) / 2 + j] = v * iv;
                        }
                }
                const int c = lane;
                double col[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    double v = (c < 16) ? Mnx[i * 16 + (c & 15)] : ((i == c - 16) ? 1.0 : 0.0);
#pragma unroll
                    for (int t = 0; t < 8; t++) if (t < i) v -= Lr[i * (i + 1) / 2 + t] * col[t];
                    col[i] = v * inv[i];
                }
#pragma unroll
                for (int i = 0; i < 8; i++) { if (c < 16) Lam[i * 16 + c] = col[i]; else X[i * 8 + c - 16] = col[i]; }
                return !pd;
            });
            if (bad) { ok = false; break; }
            // F6: P_k = [Mxx 0; 0 Mww] - Lam' Lam; F7: store the factor record
            W.each([&](int lane) {
#if defined(__CUDA_ARCH__)
                // 16 x 16 += (16 x 8)(8 x 16) as 2 x 2 fragments x 2 k-steps of mma.m8n8k4.f64: the A fragment of block row m
                // and the B fragment of block column m hold the same Lam entries, so a lane loads 4 doubles and 4 pairs
                if (lane < 32) {
                const int fr = lane >> 2, fq = lane & 3;
                double lf[2][2];
#pragma unroll
                for (int m = 0; m < 2; m++)
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) lf[m][ks] = Lam[(ks * 4 + fq) * 16 + m * 8 + fr];
#pragma unroll
                for (int mb = 0; mb < 2; mb++)
#pragma unroll
                    for (int nb = 0; nb < 2; nb++) {
                        const int o = (mb * 8 + fr) * 16 + nb * 8 + 2 * fq;
                        double c0 = PM[o], c1 = PM[o + 1];
                        warp_dmma884(c0, c1, -lf[mb][0], lf[nb][0]);
                        warp_dmma884(c0, c1, -lf[mb][1], lf[nb][1]);
                        Pc[o] = c0; Pc[o + 1] = c1;
                    }
                }
#else
                if (lane < 32) {
                const int r0 = 4 * (lane >> 3), c0 = 2 * (lane & 7);
                double acc[4][2];
#pragma unroll
                for (int a = 0; a < 4; a++) { acc[a][0] = PM[(r0 + a) * 16 + c0]; acc[a][1] = PM[(r0 + a) * 16 + c0 + 1]; }
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const double l0 = Lam[t * 16 + c0], l1 = Lam[t * 16 + c0 + 1];
#pragma unroll
                    for (int a = 0; a < 4; a++) { const double lr = Lam[t * 16 + r0 + a]; acc[a][0] -= lr * l0; acc[a][1] -= lr * l1; }
                }
#pragma unroll
                for (int a = 0; a < 4; a++) { Pc[(r0 + a) * 16 + c0] = acc[a][0]; Pc[(r0 + a) * 16 + c0 + 1] = acc[a][1]; }
                }
#endif
                double* F = FACT + (size_t)k * WF_SIZE;
                MPCC_ROLLED
                for (int e = lane; e < WF_SIZE; e += NL) F[e] = X[e];  // X and Lam are contiguous in the scratch
                async_wait<0>();  // the next stage's inputs have landed
            });
        }
        return ok;
    }

    // ---- Riccati vector sweeps: Newton step for the gradient in G -> STEP (shared) ----
    // The factor records stream through a 4-slot shared-memory ring filled by asynchronous copies three stages ahead;
    // the gradient is pulled into shared memory in one burst and kappa never leaves it.
    MPCC_HD void issue_factor_copy(int lane, int k, double* ring) const {
        if (k >= 0 && k < N) {
            const double* F = FACT + (size_t)k * WF_SIZE + 2 * lane;
            double* dst = ring + (k & (SW_RING - 1)) * WF_SIZE + 2 * lane;
#pragma unroll
            for (int t = 0; t < (WF_SIZE / 2 + NL - 1) / NL; t++) if (lane + NL * t < WF_SIZE / 2) async_copy16(dst + 2 * NL * t, F + 2 * NL * t);
        }
        async_commit();
    }
    MPCC_HD void solve_step() const {
        double* ring = SC;                    // the factorisation's blocks are dead during the sweeps
        const int gk_room = SC_VEC - SW_GK;   // doubles available behind the ring
        double* GS_ = (25 * S <= gk_room) ? SC + SW_GK : SC + SC_SIZE + ((HZ * S <= XG_ROOM) ? 0 : HZ * S);  // gradient copy [S][17]
        double* KS = GS_ + S * HZ;                                      // kappa [S][8]
        double* V = SC + SC_VEC;
        W.each([&](int lane) {
            issue_factor_copy(lane, N - 1, ring);
            issue_factor_copy(lane, N - 2, ring);
            issue_factor_copy(lane, N - 3, ring);
            if (((GS_ - SC) & 1) == 0) {  // pairs; an odd S * HZ copies one unused double (kappa's first slot, written later)
                MPCC_ROLLED
                for (int e = lane; 2 * e < S * HZ; e += NL) async_copy16(GS_ + 2 * e, G + 2 * e);
            } else {
                MPCC_ROLLED
                for (int e = lane; e < S * HZ; e += NL) async_copy8(GS_ + e, G + e);
            }
            async_commit();
            if (lane < 16) V[V_D0 + lane] = 0.0;
            async_wait<0>();
        });
        W.each([&](int lane) { if (lane < NX) V[V_D0 + lane] = GS_[N * HZ + lane]; });  // p_N
        // backward: kappa = L^-1 mn,  p_k = mx - Lam' kappa
        for (int k = N - 1; k >= 0; k--) {
            const double* Fs = ring + (k & (SW_RING - 1)) * WF_SIZE;
            const double* g = GS_ + k * HZ;
            const double* p = V + (((N - 1 - k) & 1) ? V_D1 : V_D0);
            double* pn = V + (((N - 1 - k) & 1) ? V_D0 : V_D1);
            W.each([&](int lane) {
                issue_factor_copy(lane, k - 3, ring);  // its slot held stage k+1, which is finished
                if (lane < 8) {
                    double kp = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const double mn = (t < 7) ? g[NX + t] + d_bq(t) * p[t] + p[9 + t] : g[NX + 7] + d_bs() * p[7] + d_bv() * p[8];
                        kp += Fs[WF_X + lane * 8 + t] * mn;  // X: lower triangular with explicit zeros
                    }
                    KS[k * 8 + lane] = kp;
                }
            });
            W.each([&](int lane) {
                if (lane < 16) {
                    double mx = 0;
                    if (lane < 9) { mx = g[lane] + p[lane]; if (lane == 8) mx += d_asv() * p[7]; }
#pragma unroll
                    for (int t = 0; t < 8; t++) mx -= Fs[WF_LAM + 16 * t + lane] * KS[k * 8 + t];
                    pn[lane] = mx;
                }
                async_wait<2>();  // pending: stages k-1, k-2, k-3 -> k-1 has landed
            });
        }
        // forward
        W.each([&](int lane) {
            issue_factor_copy(lane, 0, ring);
            issue_factor_copy(lane, 1, ring);
            issue_factor_copy(lane, 2, ring);
            if (lane < 16) { V[V_D0 + lane] = 0.0; V[V_D1 + lane] = 0.0; }
            if (lane < NX) STEP[lane] = 0.0;
            async_wait<2>();
        });
        for (int k = 0; k < N; k++) {
            const double* Fs = ring + (k & (SW_RING - 1)) * WF_SIZE;
            const double* d = V + ((k & 1) ? V_D1 : V_D0);
            double* dn_ = V + ((k & 1) ? V_D0 : V_D1);
            W.each([&](int lane) {
                if (k + 3 < N) issue_factor_copy(lane, k + 3, ring); else async_commit();
                if (lane < 8) {
                    double s = KS[k * 8 + lane];
#pragma unroll
                    for (int c = 0; c < 16; c++) s += Fs[WF_LAM + 16 * lane + c] * d[c];
                    V[V_RHS + lane] = -s;
                }
            });
            W.each([&](int lane) {
                if (lane < 16) {
                    const int i = (lane < 8) ? lane : ((lane == 8) ? 7 : lane - 9);  // the entry of dn = L^-T rhs this lane needs
                    double dni = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) dni += Fs[WF_X + t * 8 + i] * V[V_RHS + t];
                    double nx;
                    if (lane < 7) nx = d[lane] + d_bq(lane) * dni;
                    else if (lane == 7) nx = d[7] + d_asv() * d[8] + d_bs() * dni;
                    else if (lane == 8) nx = d[8] + d_bv() * dni;
                    else nx = dni;
                    dn_[lane] = nx;
                    if (lane < NX) STEP[(k + 1) * HZ + lane] = nx;
                    if (lane < 8) STEP[k * HZ + NX + lane] = dni;
                }
                async_wait<2>();
            });
        }
        W.each([&](int lane) { if (lane < NU) STEP[N * HZ + NX + lane] = 0.0; async_wait<0>(); });
    }

    // slack / multiplier steps from the primal step; largest step keeping t, lam > 0.  With sums != nullptr also returns
    // s1 = sum(t dl + lam dt) and s2 = sum(dt dl), from which mu(alpha) = (sum t lam + alpha s1 + alpha^2 s2) / m follows.
    MPCC_HD double ineq_steps(double* sums) const {
        double* RED = red_base();
        W.each([&](int lane) { RED[lane] = 1.0; RED[NL + lane] = 0.0; RED[2 * NL + lane] = 0.0; });
        double* dt_ = IDT; double* dl_ = IDLAM;
        stream_constraints(vec_ids(CV_RP, CV_LAM, CV_V, CV_W, CV_T), 5, [&](int lane, int i, int kind, int k, int c, const double* v, int j, const double* row) {
            constexpr int TSB = TS1, TSP = TS2;
            const int ts = (kind == 2) ? TSP : TSB;
            const double rp = v[j], lam = v[ts + j], vv = v[2 * ts + j], w = v[3 * ts + j], t = v[4 * ts + j];
            const double g = gz_of(STEP, kind, k, c, row);
            const double dt = -rp - g;
            const double dl = -lam + vv + w * g;
            dt_[i] = dt; dl_[i] = dl;
            double a = RED[lane];
            if (dt < 0) a = fmin(a, -t / dt);
            if (dl < 0) a = fmin(a, -lam / dl);
            RED[lane] = a;
            RED[NL + lane] += t * dl + lam * dt;
            RED[2 * NL + lane] += dt * dl;
        });
        if (sums) {
            sums[0] = W.rsum([&](int lane) { return RED[NL + lane]; });
            sums[1] = W.rsum([&](int lane) { return RED[2 * NL + lane]; });
        }
        return W.rmin([&](int lane) { return RED[lane]; });
    }

    // ---- interior-point loop; on success VAR holds the step (xi = exact rollout of nu) ----
    MPCC_HD QpStats solve() const {
        QpStats st;
        st.ok = 0; st.iters = 0; st.res_dual = 0; st.res_prim = 0; st.gap = 0; st.infeasible = 0;
        const int tot = S * NINEQ;
        double* RED = red_base();
        // feasibility of the boxes (stage 0: xi_0 = 0 must lie inside; others: lo <= hi)
        if (W.any([&](int lane) {
                bool bad = false;
                MPCC_ROLLED
                for (int o = lane; o < S * NX; o += NL) {
                    const int k = o / NX, m = o - k * NX;
                    const double* L = LIN + (size_t)k * WL_SIZE;
                    if (k == 0) { if (L[WL_XLO + m] > 1e-9 || L[WL_XHI + m] < -1e-9) bad = true; }
                    else if (L[WL_XLO + m] > L[WL_XHI + m]) bad = true;
                }
                return bad;
            })) return st;
        // initial point: nu = 0, xi = rollout of the defects (lane = state component), t = max(h - Gz, QP_INIT_SLACK), lam = QP_INIT_SLACK / t.
        // The loops below read the stage records from global memory; each lane issues FOUR independent loads before the first use (a rolled
        // load -> use -> store loop pays one memory round trip per item: the next load may not pass the store), and the defects b are staged in
        // shared memory (STEP is free here and again after the last iteration) before the sequential rollouts walk them.
        auto stage_defects = [&]() {
            W.each([&](int lane) {
                const int nb = N * NX;
                MPCC_ROLLED
                for (int o = lane; o < nb; o += 4 * NL) {
                    double t[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { const int i = o + u * NL; t[u] = (i < nb) ? LIN[(size_t)(i / NX) * WL_SIZE + WL_b + (i % NX)] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 4; u++) { const int i = o + u * NL; if (i < nb) STEP[i] = t[u]; }
                }
            });
        };
        stage_defects();
        const double qn = W.rmax([&](int lane) {
            double q = 0;
            MPCC_ROLLED
            for (int o = lane; o < S * HZ; o += 4 * NL) {
                double t[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = o + u * NL, k = i / HZ, r = i - k * HZ;
                    const double* L = LIN + (size_t)k * WL_SIZE;
                    t[u] = (i >= S * HZ) ? 0.0 : (r < NX) ? L[WL_q + r] : (k < N) ? L[WL_r + r - NX] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = o + u * NL;
                    if (i < S * HZ) { q = fmax(q, fabs(t[u])); if (i - (i / HZ) * HZ >= NX) { VAR[i] = 0.0; G[i] = 0.0; } }
                }
            }
            if (lane < NX && lane != 7) {
                double x = 0;
                for (int k = 0; k <= N; k++) { VAR[k * HZ + lane] = x; if (k < N) x += STEP[k * NX + lane]; }
            }
            return q;
        });
        W.each([&](int lane) {
            if (lane == 7) {
                double x = 0;
                for (int k = 0; k <= N; k++) { VAR[k * HZ + 7] = x; if (k < N) x += d_asv() * VAR[k * HZ + 8] + STEP[k * NX + 7]; }
            }
            MPCC_ROLLED
            for (int i = lane; i < tot; i += NL) { IT[i] = 1.0; ILAM[i] = 0.0; IW[i] = 0.0; IV[i] = 0.0; IRP[i] = 0.0; IDT[i] = 0.0; IDLAM[i] = 0.0; }
        });
        const double s0 = QP_INIT_SLACK;
        W.each([&](int lane) {
            for_present4(lane, VAR, true, [&](const int* idx, const double* g, const double* h) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (idx[u] >= 0) { const double t0 = fmax(h[u] - g[u], s0); IT[idx[u]] = t0; ILAM[idx[u]] = s0 / t0; IH[idx[u]] = h[u]; }
            });
        });
        const double m_tot = 43.0 * N;
        // The certificate test sits OUTSIDE the iteration loop (the loop leaves for it and is re-entered): with a call inside the
        // loop body the steady-state kernel was 40 % slower although the call never executed (measured: 7.19 vs 5.09 ms).
        int it = 0;
        bool suspicious = false;   // the last step was short: test for an infeasibility certificate before going on
        for (;;) {
        for (; it < opt.max_iter; it++) {
            // residuals, barrier weights, predictor v = lam rp / t
            W.each([&](int lane) { RED[lane] = 0.0; RED[NL + lane] = 0.0; });
            {
                double* rp_ = IRP; double* w_ = IW; double* v_ = IV;
                stream_constraints(vec_ids(CV_T, CV_LAM, CV_H), 3, [&](int lane, int i, int kind, int k, int c, const double* v, int j, const double* row) {
                    const int ts = (kind == 2) ? TS2 : TS1;
                    const double t = v[j], lam = v[ts + j], h = v[2 * ts + j];
                    const double rp = gz_of(VAR, kind, k, c, row) + t - h;
                    rp_[i] = rp;
                    RED[lane] = fmax(RED[lane], fabs(rp));
                    RED[NL + lane] += t * lam;
                    const double w = lam / t;
                    w_[i] = w;
                    v_[i] = w * rp;
                });
            }
            const double nrp = W.rmax([&](int lane) { return RED[lane]; });
            const double sum_tl = W.rsum([&](int lane) { return RED[NL + lane]; });
            const double mu = sum_tl / m_tot;
            // Two passes over ONE copy of gradient / sweeps / step lengths (the loop is kept rolled on purpose: code footprint).
            //   pass 0: Lagrangian gradient and residual test, factorisation, affine (predictor) step, centring parameter
            //   pass 1: corrector right-hand side, combined step, update
            double sums[2] = {0.0, 0.0};
            bool stop = false;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int pass = 0; pass < 2; pass++) {
                gradient(pass == 0 ? STEP : nullptr);  // G <- gradient of this pass; pass 0: STEP <- Lagrangian gradient (scratch until the step is computed)
                if (pass == 0) {
                    const double nrd = costate_residual();
                    st.iters = it; st.res_dual = nrd; st.res_prim = nrp; st.gap = mu;
#if defined(MPCC_QP_TRACE) && !defined(__CUDA_ARCH__)
                    fprintf(stderr, "  ipm %2d  rd %.3e  rp %.3e  mu %.3e\n", it, nrd, nrp, mu);
#endif
                    if (nrd <= opt.eps * (1.0 + qn) && nrp <= opt.eps && mu <= opt.eps) { st.ok = 1; stop = true; break; }
                    if (!(nrd == nrd) || !(mu == mu)) { stop = true; break; }
                    if (!factor()) { stop = true; break; }
                }
                solve_step();
                const double a_max = ineq_steps(pass == 0 ? sums : nullptr);
                if (pass == 0) {
                    const double mu_aff = (sum_tl + a_max * sums[0] + a_max * a_max * sums[1]) / m_tot;
                    const double sigma = (mu > 0) ? (mu_aff / mu) * (mu_aff / mu) * (mu_aff / mu) : 0.0;
                    // corrector: v = (lam rp + sigma mu - dt_a dlam_a) / t
                    const double sm = sigma * mu;
                    double* v_ = IV;
                    stream_constraints(vec_ids(CV_LAM, CV_RP, CV_DT, CV_DLAM, CV_T), 5, [&](int, int i, int kind, int, int, const double* v, int j, const double*) {
                        const int ts = (kind == 2) ? TS2 : TS1;
                        v_[i] = (v[j] * v[ts + j] + sm - v[2 * ts + j] * v[3 * ts + j]) / v[4 * ts + j];
                    });
                } else {
                    const double a = fmin(1.0, qp_step_tau(mu) * a_max);
                    if (a < 0.01 && it >= 1 && nrp > opt.eps) suspicious = true;   // a short step announces trouble
#if defined(MPCC_QP_TRACE) && !defined(__CUDA_ARCH__)
                    fprintf(stderr, "          alpha %.3e\n", a);
#endif
                    W.each([&](int lane) { MPCC_ROLLED for (int o = lane; o < S * HZ; o += NL) VAR[o] += a * STEP[o]; });
                    double* t_ = IT; double* l_ = ILAM;
                    stream_constraints(vec_ids(CV_T, CV_LAM, CV_DT, CV_DLAM), 4, [&](int, int i, int kind, int, int, const double* v, int j, const double*) {
                        const int ts = (kind == 2) ? TS2 : TS1;
                        t_[i] = v[j] + a * v[2 * ts + j];
                        l_[i] = v[ts + j] + a * v[3 * ts + j];
                    });
                }
            }
            if (stop) break;
            st.iters = it + 1;
            if (suspicious) { it++; break; }
        }
#if !defined(MPCC_NO_CERT)
        if (!suspicious) break;
        suspicious = false;
        // test the multipliers for a Farkas ray before spending the remaining iterations
        if (primal_infeasible(*this, 1e-8)) { st.infeasible = 1; break; }
#else
        break;
#endif
        }
        if (st.ok) {
            // make the equalities exact: xi = rollout(nu)
            stage_defects();
            W.each([&](int lane) {
                if (lane < NX && lane != 7) {
                    double x = 0;
                    for (int k = 0; k <= N; k++) {
                        VAR[k * HZ + lane] = x;
                        if (k < N) {
                            const double b = STEP[k * NX + lane];
                            x = (lane < 7) ? x + d_bq(lane) * VAR[k * HZ + NX + lane] + b : x + d_bv() * VAR[k * HZ + NX + 7] + b;
                        }
                    }
                }
            });
            W.each([&](int lane) {
                if (lane == 7) {
                    double x = 0;
                    for (int k = 0; k <= N; k++) {
                        VAR[k * HZ + 7] = x;
                        if (k < N) x = x + d_asv() * VAR[k * HZ + 8] + d_bs() * VAR[k * HZ + NX + 7] + STEP[k * NX + 7];
                    }
                }
            });
        }
        return st;
    }

    // the evaluation point guess + alpha T step, gathered once into shared memory (VAR is free outside the QP solve)
    MPCC_HD void gather_point(double alpha) const {
        double* XT = VAR;
        W.each([&](int lane) {
            for (int e = lane; e < S * HZ; e += NL) {
                const int kk = e / HZ, r = e - kk * HZ;
                double v = XG[e];
                if (alpha != 0.0) {
                    if (r < NX) v += alpha * (Tx(r) * XS[e]);
                    else v = (kk < N) ? v + alpha * (Tu(r - NX) * XS[e]) : 0.0;
                }
                XT[e] = v;
            }
        });
    }

    // ---- horizon evaluation at guess + alpha T step, lane = stage.
    //      FULL: fill LIN (and, if write_cst, the polytopic rows), test the Hessian blocks (notpd / nan);
    //      else only the objective and the l1 constraint violation ----
    template <bool FULL>
    MPCC_HD void eval_horizon(const double* cur_u, const double* rb, size_t rb_stride, size_t rb_stage, double alpha, bool write_cst, double& obj, double& gap,
                              bool* notpd, bool* nan, bool write_lin = true, bool gathered = false) const {
        double* RED = red_base();
        const double* XT = VAR;  // the evaluation point, see gather_point()
        if (!gathered) gather_point(alpha);
        obj = W.rsum([&](int lane) {
            double o_acc = 0, g_acc = 0;
            bool pd_l = true, nan_l = false;
            for (int k = lane; k <= N; k += NL) {
                double x[NX], u[NU], up[DOF], un[DOF], xn[NX];
                auto gx = [&](int kk, int e) -> double { return XT[kk * HZ + e]; };
                for (int e = 0; e < NX; e++) x[e] = gx(k, e);
                for (int e = 0; e < NU; e++) u[e] = gx(k, NX + e);
                for (int j = 0; j < DOF; j++) {
                    up[j] = (k == 0) ? cur_u[j] : gx(k - 1, NX + j);
                    un[j] = (k < N) ? gx(k + 1, NX + j) : 0.0;
                }
                for (int e = 0; e < NX; e++) xn[e] = (k < N) ? gx(k + 1, e) : 0.0;
                StageLin sl;
                // pull this stage's RobotData record in one burst (independent loads in flight together) instead of
                // one dependent global load at a time inside stage_eval
                double rbl[RB_DOUBLES];
                {
                    // (the joint angles are not read -- the evaluation point carries them -- and the Jacobians only by the
                    //  full linearisation: 101 / 143 of the 150 doubles)
                    const double* src = rb + (size_t)k * rb_stage;
#pragma unroll 12
                    for (int e = RB_P; e < RB_JV; e++) rbl[e] = src[(size_t)e * rb_stride];
                    if (FULL) {
#pragma unroll 21
                        for (int e = RB_JV; e < RB_MANIP; e++) rbl[e] = src[(size_t)e * rb_stride];
                    }
#pragma unroll 30
                    for (int e = RB_MANIP; e < RB_DOUBLES; e++) rbl[e] = src[(size_t)e * rb_stride];
                }
                RbView rv{rbl, 1};
                // the relaxed-barrier values of the polytopic rows are constant over the cycle: computed by the first
                // linearisation (write_cst), re-read afterwards (11 logarithms and divisions less per evaluation)
                double rbfv[2 * NPOLY];
                double* rbf_home = RBFV + (size_t)k * 2 * NPOLY;
                if (!write_cst) {
#pragma unroll
                    for (int j = 0; j < 2 * NPOLY; j++) rbfv[j] = rbf_home[j];
                }
                stage_eval<FULL>(P, T, Ts, N, k, x, u, up, un, xn, rv, sl, write_cst ? nullptr : rbfv, write_cst ? rbf_home : nullptr, FULL && !write_lin);
                o_acc += sl.obj; g_acc += sl.gap;
                if (FULL) {
                    block9_pd_nan(sl.Q, pd_l, nan_l);
                    if (k < N) for (int j = 0; j < NU; j++) if (sl.Rd[j] != sl.Rd[j]) nan_l = true;
                    if (!write_lin) continue;
                    double* L = LIN + (size_t)k * WL_SIZE;
                    for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) L[WL_Q + r * 9 + c] = sl.Q[(r >= c) ? sym9(r, c) : sym9(c, r)];
                    for (int m = 0; m < NX; m++) { L[WL_q + m] = sl.q[m]; L[WL_b + m] = sl.b[m]; L[WL_XLO + m] = sl.xlo[m]; L[WL_XHI + m] = sl.xhi[m]; }
                    for (int j = 0; j < NU; j++) { L[WL_RD + j] = sl.Rd[j]; L[WL_r + j] = sl.r[j]; }
                    for (int j = 0; j < DOF; j++) { L[WL_DLO + j] = sl.dlo[j]; L[WL_DHI + j] = sl.dhi[j]; }
                    for (int j = 0; j < NPOLY; j++) L[WL_PRHS + j] = sl.prhs[j];
                    if (write_cst && k < N) {
                        // the polytopic rows depend only on the (frozen) RobotData: constant over the cycle
                        double* C = CST + (size_t)k * WC_SIZE;
                        for (int j = 0; j < NPOLY; j++)
                            for (int m = 0; m < DOF; m++) {
                                const double g = sl.pg[j * DOF + m];
                                C[j * 14 + m] = sl.pd[j] * g * Tx(m);
                                C[j * 14 + 7 + m] = -g * Tu(m);
                            }
                    }
                }
            }
            if (FULL && lane < NU) {
                // input block of the Hessian: per joint a tridiagonal (Rd, cpl) chain over the stages, dVs diagonal.
                // Rd depends on the parameters only (dev_stage.cuh): same expression, no loads.
                const int j = lane;
                const double tu = P.Tu[j];
                double d = 0;
                for (int k = 0; k < N; k++) {
                    const double kap = (k == 0 || k == N - 1) ? 2.0 : 4.0;
                    const double rd = (j < 7) ? tu * ((2.0 * P.r_dq + 1e-6) + kap * P.r_ddq_solver) * tu : tu * (2.0 * P.r_dVs + 1e-6) * tu;
                    d = (k == 0 || j == 7) ? rd : rd - d_cpl(j) * d_cpl(j) / d;
                    if (d <= 0.0) { pd_l = false; break; }
                }
            }
            RED[NL + lane] = g_acc;
            RED[lane] = (pd_l ? 0.0 : 1.0) + (nan_l ? 2.0 : 0.0);
            return o_acc;
        });
        gap = W.rsum([&](int lane) { return RED[NL + lane]; });
        if (FULL) {
            *notpd = W.any([&](int lane) { const int f = (int)RED[lane]; return (f & 1) != 0; });
            *nan = W.any([&](int lane) { const int f = (int)RED[lane]; return (f & 2) != 0; });
        }
    }

    // Will the QP of the linearisation at the iterate fail solve()'s box test (xi_0 = 0 outside its box, or an empty box at a
    // later stage)?  The boxes depend on the iterate only (state bounds, s trust region, the mis-indexed input-bound rows):
    // same expressions as stage_eval + the quirk pass, evaluated before linearising so that a QP known to fail is not assembled.
    MPCC_HD bool qp_box_infeasible(const double* X) const {
        return W.any([&](int lane) {
            const double Lt = T.s[N_SPLINE - 1];
            bool bad = false;
            for (int o = lane; o < S * NX; o += NL) {
                const int k = o / NX, m = o - k * NX;
                const double x = X[k * HZ + m], sv = X[k * HZ + 7];
                double lo = P.lx[m], hi = P.ux[m];
                if (m == 7) { lo = fmax(sv - P.s_trust_region, 0.0); hi = fmin(sv + P.s_trust_region, Lt); }
                double xlo = (lo - x) / P.Tx[m], xhi = (hi - x) / P.Tx[m];
                if (o < NU * N) {  // flat column c = o of the input-bound rows: row (i, kk) with 8 i + kk = c
                    const int i = o / NU, kk = o - i * NU;
                    const double uv = X[i * HZ + NX + kk];
                    xlo = fmax(xlo, (P.lu[kk] - uv) / Tu(kk));
                    xhi = fmin(xhi, (P.uu[kk] - uv) / Tu(kk));
                }
                if (k == 0) { if (xlo > 1e-9 || xhi < -1e-9) bad = true; }
                else if (xlo > xhi) bad = true;
            }
            return bad;
        });
    }

    // ---- SecondOrderCorrection (osqp_interface.cpp:658-681): the QP is solved again with the same P, q, A and the bounds shifted by
    //      d = c(x (+) step) - A step, where (+) adds the NORMALISED step to the unnormalised iterate (:661).  Every row of c is affine in the
    //      iterate once RobotData is frozen (quirk 6), so the shifted bounds follow in closed form from the point x~ = x (+) step, the step and the
    //      rows already in LIN / CST; they overwrite the right-hand sides in LIN (dead after this QP: the next iteration linearises anew).
    //      In:  XG iterate, XS step of the first QP (or the stale one).  Uses VAR for x~. ----
    MPCC_HD void soc_right_hand_sides(const double* cur_u) const {
        double* XT = VAR;
        W.each([&](int lane) {
            for (int e = lane; e < S * HZ; e += NL) {
                const int kk = e / HZ, r = e - kk * HZ;
                XT[e] = (r < NX || kk < N) ? XG[e] + XS[e] : 0.0;   // vectorToOptvar zeroes uk[N] (:835-845)
            }
        });
        W.each([&](int lane) {
            const double Lt = T.s[N_SPLINE - 1];
            // state boxes incl. the mis-indexed input-bound rows (flat column o of the state block), dynamics defects
            for (int o = lane; o < S * NX; o += NL) {
                const int k = o / NX, m = o - k * NX;
                const double x = XT[k * HZ + m], sv = XT[k * HZ + 7];
                double lo = P.lx[m], hi = P.ux[m];
                if (m == 7) { lo = fmax(sv - P.s_trust_region, 0.0); hi = fmin(sv + P.s_trust_region, Lt); }
                double xlo = (lo - x) / P.Tx[m], xhi = (hi - x) / P.Tx[m];
                if (o < NU * N) {
                    const int i = o / NU, kk = o - i * NU;
                    const double uv = XT[i * HZ + NX + kk];
                    xlo = fmax(xlo, (P.lu[kk] - uv) / Tu(kk));
                    xhi = fmin(xhi, (P.uu[kk] - uv) / Tu(kk));
                }
                const double sx = XS[k * HZ + m];
                double* L = LIN + (size_t)k * WL_SIZE;
                L[WL_XLO + m] = xlo + sx; L[WL_XHI + m] = xhi + sx;
                if (k < N) {
                    const double* xk = XT + k * HZ;
                    double pred, as;   // prediction at x~ (osqp_interface.cpp:247) and the dynamics row applied to the step
                    const double* sk = XS + k * HZ;
                    if (m < 7) { pred = xk[m] + Ts * xk[NX + m]; as = XS[(k + 1) * HZ + m] - sk[m] - d_bq(m) * sk[NX + m]; }
                    else if (m == 7) { pred = xk[7] + Ts * xk[8] + 0.5 * Ts * Ts * xk[NX + 7]; as = XS[(k + 1) * HZ + 7] - sk[7] - d_asv() * sk[8] - d_bs() * sk[NX + 7]; }
                    else { pred = xk[8] + Ts * xk[NX + 7]; as = XS[(k + 1) * HZ + 8] - sk[8] - d_bv() * sk[NX + 7]; }
                    const double c = (1.0 / P.Tx[m]) * (XT[(k + 1) * HZ + m] - pred);
                    L[WL_b + m] = -c + as;
                }
            }
            // joint-acceleration rows (osqp_interface.cpp:279-297)
            for (int o = lane; o < N * DOF; o += NL) {
                const int k = o / DOF, j = o - k * DOF;
                const double uj = XT[k * HZ + NX + j];
                double lo, hi, c;
                if (k == 0) { c = 1. / Ts * uj; lo = P.ldd[j] + 1. / Ts * cur_u[j]; hi = P.udd[j] + 1. / Ts * cur_u[j]; }
                else { c = 1. / Ts * (uj - XT[(k - 1) * HZ + NX + j]); lo = P.ldd[j]; hi = P.udd[j]; }
                const double as = XS[k * HZ + NX + j] - (k > 0 ? XS[(k - 1) * HZ + NX + j] : 0.0);
                double* L = LIN + (size_t)k * WL_SIZE;
                L[WL_DLO + j] = (lo - c) * Ts / P.Tu[j] + as;
                L[WL_DHI + j] = (hi - c) * Ts / P.Tu[j] + as;
            }
            // polytopic rows: c(x~) = -g' (u + nu) + RBF(frozen), A step = pd g' Tx xi - g' Tu nu with the rows [pd g Tx | -g Tu] of CST
            for (int o = lane; o < N * NPOLY; o += NL) {
                const int k = o / NPOLY, j = o - k * NPOLY;
                const double* row = CST + ((size_t)k * NPOLY + j) * 14;
                const double* sk = XS + k * HZ;
                double add = 0;
#pragma unroll
                for (int m = 0; m < DOF; m++) add += row[m] * sk[m] + row[7 + m] * sk[NX + m] * (1.0 - 1.0 / Tu(m));
                LIN[(size_t)k * WL_SIZE + WL_PRHS + j] += add;
            }
        });
    }

    // ---- the SQP loop (solveOCP) ----
    MPCC_HD SqpResult run(const double* cur_u, const double* rb, size_t rb_stride, size_t rb_stage, SqpLogRef* log) {
        SqpResult res;
        res.status = SOLVED; res.iters = 0; res.qp_fail = 0; res.qp_iters = 0; res.accept_mask = 0;
        const int max_iter = (int)P.max_iter, ls_max = (int)P.line_search_max_iter;
        const int HN = S * HZ;
        const bool soc_on = SOC && P.do_SOC != 0.0;
        init_scratch();
        W.each([&](int lane) { for (int e = lane; e < HN; e += NL) { XS[e] = 0.0; XG[e] = GUESS[e]; } });
        int n_filt = 0, it = 0;
        bool done = false;
        bool have_lin = false, lin_notpd = false, lin_nan = false;  // LIN already holds the linearisation of the iterate
        bool lin_infeasible = false;                               // ... and its QP is known to fail the box test (LIN not written)
        double inf_step = 0.0;                                       // inf-norm of the persistent step
        bool last_rejected = false;
        for (it = 0; it < max_iter; it++) {
            const double t_a = now_ns();
            bool qp_known_infeasible = lin_infeasible;
            if (!have_lin) {
                double obj, gap;
                gather_point(0.0);
                qp_known_infeasible = (it > 0) && !soc_on && qp_box_infeasible(VAR);  // (it == 0 also writes the cycle constants; the correction needs the assembled QP)
                eval_horizon<true>(cur_u, rb, rb_stride, rb_stage, 0.0, it == 0, obj, gap, &lin_notpd, &lin_nan, !qp_known_infeasible, true);
            }
            have_lin = false; lin_infeasible = false;
            // mis-indexed input-bound rows (osqp_interface.cpp:273) intersected into the state boxes
            if (!qp_known_infeasible) W.each([&](int lane) {
                for (int c = lane; c < NU * N; c += NL) {
                    const int k = c / NX, m = c - k * NX, i = c / NU, kk = c - i * NU;
                    const double uv = XG[i * HZ + NX + kk];
                    const double lo = (P.lu[kk] - uv) / Tu(kk), hi = (P.uu[kk] - uv) / Tu(kk);
                    double* L = LIN + (size_t)k * WL_SIZE;
                    L[WL_XLO + m] = fmax(L[WL_XLO + m], lo);
                    L[WL_XHI + m] = fmin(L[WL_XHI + m], hi);
                }
            });
            // isPosdef / isNan on the block structure of the Hessian (osqp_interface.cpp:454-473), tested while linearising
            if (lin_notpd) { res.status = NON_PD_HESSIAN; done = true; break; }
            if (lin_nan) { res.status = NAN_HESSIAN; done = true; break; }
            const double t_b = now_ns();
            QpStats qs;
            if (qp_known_infeasible) { qs.ok = 0; qs.iters = 0; qs.res_dual = qs.res_prim = qs.gap = 0; }
            else {
                W.each([&](int lane) { for (int e = lane; e < HN; e += NL) { GUESS[e] = XG[e]; SSTEP[e] = XS[e]; } });  // the QP solve uses all of the scratch
                qs = solve();
                W.each([&](int lane) { for (int e = lane; e < HN; e += NL) XG[e] = GUESS[e]; });
            }
            res.qp_iters += qs.iters;
            if (qs.ok) {
                inf_step = W.rmax([&](int lane) {
                    double m = 0;
                    for (int e = lane; e < HN; e += NL) { const int k = e / HZ, r = e - k * HZ; const double v = (r < NX || k < N) ? VAR[e] : 0.0; XS[e] = v; m = fmax(m, fabs(v)); }
                    return m;
                });
            } else {
                res.qp_fail++;  // step keeps its previous value (osqp_interface.cpp:479-505)
                if (!qp_known_infeasible) W.each([&](int lane) { for (int e = lane; e < HN; e += NL) XS[e] = SSTEP[e]; });
            }
            if (SOC && soc_on) {
                // second-order correction: second QP on the shifted bounds; its solution replaces the step, a failure only counts (:506-533, :646-647)
                soc_right_hand_sides(cur_u);
                W.each([&](int lane) { for (int e = lane; e < HN; e += NL) { GUESS[e] = XG[e]; SSTEP[e] = XS[e]; } });
                const QpStats q2 = solve();
                W.each([&](int lane) { for (int e = lane; e < HN; e += NL) XG[e] = GUESS[e]; });
                res.qp_iters += q2.iters;
                if (q2.ok) {
                    inf_step = W.rmax([&](int lane) {
                        double m = 0;
                        for (int e = lane; e < HN; e += NL) { const int k = e / HZ, r = e - k * HZ; const double v = (r < NX || k < N) ? VAR[e] : 0.0; XS[e] = v; m = fmax(m, fabs(v)); }
                        return m;
                    });
                } else {
                    res.qp_fail++;
                    W.each([&](int lane) { for (int e = lane; e < HN; e += NL) XS[e] = SSTEP[e]; });
                }
            }
            const double t_c = now_ns();
            tm_set_qp += t_b - t_a; tm_solve_qp += t_c - t_b;
            // ---- filterLineSearch (osqp_interface.cpp:759-808) ----
            double alpha = 1.0;
            bool accepted = true;  // never reset inside the loop (:767)
            for (int i = 0; i < ls_max; i++) {
                if (accepted) {  // once a trial is rejected no later trial can be accepted: evaluating them is dead work
                    double o2, g2;
                    // If this trial is accepted and the loop goes on (alpha |step| >= eps_prim), the next iteration linearises
                    // exactly at this point: evaluate it in full right away instead of values now and everything later.
                    const bool spec = (i == 0) && (inf_step >= P.eps_prim) && (it + 1 < max_iter) && !last_rejected;
                    bool sp_notpd = false, sp_nan = false, sp_infeasible = false;
                    if (spec) {
                        gather_point(alpha);
                        sp_infeasible = !soc_on && qp_box_infeasible(VAR);
                        eval_horizon<true>(cur_u, rb, rb_stride, rb_stage, alpha, false, o2, g2, &sp_notpd, &sp_nan, !sp_infeasible, true);
                    } else eval_horizon<false>(cur_u, rb, rb_stride, rb_stage, alpha, false, o2, g2, nullptr, nullptr);
                    if (W.any([&](int lane) {
                            bool dom = false;
                            for (int j = lane; j < n_filt; j += NL) if (o2 >= FILT[2 * j] && g2 >= FILT[2 * j + 1]) dom = true;
                            return dom;
                        })) accepted = false;
                    if (accepted) {
                        // prune the entries the new point dominates, append it (sequential: order matters)
                        int w = 0;
                        for (int j = 0; j < n_filt; j++) {
                            const double fo = FILT[2 * j], fg = FILT[2 * j + 1];
                            if (o2 > fo || g2 > fg) {
                                if (w != j) W.each([&](int lane) { if (lane == 0) { FILT[2 * w] = fo; FILT[2 * w + 1] = fg; } });
                                w++;
                            }
                        }
                        W.each([&](int lane) { if (lane == 0) { FILT[2 * w] = o2; FILT[2 * w + 1] = g2; } });
                        n_filt = w + 1;
                        if (spec) { have_lin = true; lin_notpd = sp_notpd; lin_nan = sp_nan; lin_infeasible = sp_infeasible; }
                        break;
                    }
                }
                alpha *= P.line_search_tau;
            }
            last_rejected = !accepted;  // after a rejection do not speculate: rejections come in runs (stale step, tiny alpha)
            tm_get_alpha += now_ns() - t_c;
            if (accepted && it < 32) res.accept_mask |= (1u << it);
            // ---- take the step (osqp_interface.cpp:549-551) ----
            W.each([&](int lane) {
                for (int e = lane; e < HN; e += NL) {
                    const int k = e / HZ, r = e - k * HZ;
                    const double s = XS[e];
                    if (r < NX) XG[e] += alpha * (Tx(r) * s);
                    else if (k < N) XG[e] += alpha * (Tu(r - NX) * s);
                    else XG[e] = 0.0;
                }
            });
            const double inf = inf_step;
            if (log && log->n < log->max_log) {
                W.each([&](int lane) {
                    if (log->steps) for (int e = lane; e < HN; e += NL) log->steps[(size_t)log->n * HN + e] = XS[e];
                    if (lane == 0) { log->alphas[log->n] = alpha; log->qp_ok[log->n] = qs.ok; }
                });
                log->n++;
            }
            if (alpha * inf < P.eps_prim) { res.status = SOLVED; res.iters = it + 1; done = true; break; }
        }
        if (!done) { res.status = MAX_ITER_EXCEEDED; res.iters = max_iter; }
        else if (res.status != SOLVED) res.iters = it;
        W.each([&](int lane) { for (int e = lane; e < HN; e += NL) GUESS[e] = XG[e]; });
        return res;
    }
};

using WarpSqp = GroupSqp<32>;

}  // namespace mpcc
