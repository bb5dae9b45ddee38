// Warp-cooperative SQP loop: ONE WARP PER MPCC INSTANCE.
//
// Same algorithm as the one-thread-per-instance formulation in dev_sqp.cuh / dev_qp.cuh (which stays as the
// host-compilable statement of the method and as the device kernel selected by sqp_kernel = 1):
//   reference SQP loop          cpp/src/Interfaces/osqp_interface.cpp:398-590
//   filter line search          :759-808
//   QP (replaces OSQP, :592-656) Mehrotra predictor-corrector interior point; every Newton system solved by a
//                               Riccati recursion on the state augmented with the previous joint-velocity step.
// Work distribution inside the warp:
//   * stage linearisation / trial evaluation: lane = stage (stage_eval is independent per stage);
//   * every per-constraint and per-variable pass of the interior-point iteration: flat index over
//     (stage, item), 32 items per round, warp-shuffle reductions for norms / step lengths;
//   * the Riccati factorisation and the two triangular sweeps are sequential in the stage; inside a stage all
//     32 lanes work on the 8x8 / 8x16 / 16x16 blocks held in this warp's shared-memory scratch.
// Memory: the iterate, the current QP point and the Newton step live in shared memory; the per-stage QP blocks,
// the per-constraint interior-point vectors and the Riccati factors live in a per-instance CONTIGUOUS global
// workspace (every warp access is a run of consecutive doubles: full 128-byte lines).
//
// The dynamics multipliers are not iterated: the dual residual is evaluated with the costates obtained from the
// state-stationarity recursion  p_N = g_N,  p_k = g_k + A' p_{k+1}  (g = gradient of the Lagrangian in xi with the
// current inequality multipliers; p = -y of dev_qp.cuh), which zeroes the xi-residual by construction; the
// termination test is then on the nu-stationarity residual g_nu,k + B' p_{k+1}, the primal residual and the
// complementarity gap, with the same thresholds.
#pragma once
#include "dev_sqp.cuh"

#if defined(__CUDACC__)
namespace mpcc {

constexpr int MAX_SQP_FILTER = 128;  // largest sqp.max_iter a handle accepts (filter capacity)
constexpr int WF_L = 0, WF_INV = 36, WF_LAM = 44, WF_SIZE = 172;  // per-stage factor record: L (packed lower), 1/diag, Lam (8 x 16)
constexpr int WSC_PM = 0, WSC_MNX = 256, WSC_MNN = 384, WSC_FF = 448, WSC_GS = 520, WSC_WG = 674, WSC_VEC = 688, WSC_TXU = 752, WSC_SIZE = 772;

// doubles of global workspace per instance
__host__ __device__ inline size_t warp_ws_doubles(int N) {
    const size_t S = N + 1;
    return S * (LIN_SIZE + 7 * NINEQ + HZ /*G*/ + 8 /*KAP*/ + WF_SIZE + HZ /*persistent step*/) + 2 * (MAX_SQP_FILTER + 2);
}
// doubles of shared memory per warp
__host__ __device__ inline size_t warp_smem_doubles(int N) { return (size_t)3 * (N + 1) * HZ + WSC_SIZE; }

__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ bool wany(bool p) { return __any_sync(0xffffffffu, p); }

struct WarpSqp {
    const Params& P;
    const TrackTable& T;
    DynConst dyn;
    double Ts;
    int N, S, lane;
    QpOptions opt;
    // global, per instance
    double *LIN, *IT, *ILAM, *IRP, *IW, *IV, *IDT, *IDLAM, *G, *KAP, *FACT, *SSTEP, *FILT;
    // shared, per warp
    double *GUESS, *VAR, *STEP, *SC;

    __device__ void carve(double* gws, double* sm) {
        const size_t S_ = S;
        LIN = gws; gws += S_ * LIN_SIZE;
        IT = gws; gws += S_ * NINEQ; ILAM = gws; gws += S_ * NINEQ; IRP = gws; gws += S_ * NINEQ; IW = gws; gws += S_ * NINEQ;
        IV = gws; gws += S_ * NINEQ; IDT = gws; gws += S_ * NINEQ; IDLAM = gws; gws += S_ * NINEQ;
        G = gws; gws += S_ * HZ; KAP = gws; gws += S_ * 8; FACT = gws; gws += S_ * WF_SIZE; SSTEP = gws; gws += S_ * HZ; FILT = gws;
        GUESS = sm; VAR = sm + S_ * HZ; STEP = sm + 2 * S_ * HZ; SC = sm + 3 * S_ * HZ;
    }
    __device__ __forceinline__ double Tx(int m) const { return SC[WSC_TXU + m]; }
    __device__ __forceinline__ double Tu(int j) const { return SC[WSC_TXU + 9 + j]; }
    __device__ __forceinline__ static bool present(int N_, int k, int c) { return (c < NBOX) ? (k >= 1) : (k < N_); }

    // G z for constraint c of stage k on a [S][17] vector pair (xi | nu) held in shared memory
    __device__ double gz(const double* Z, int k, int c) const {
        const double* z = Z + k * HZ;
        if (c < 9) return -z[c];
        if (c < 18) return z[c - 9];
        if (c < 32) {
            const int j = (c < 25) ? c - 18 : c - 25;
            double d = z[NX + j];
            if (k >= 1) d -= z[NX + j - HZ];
            return (c < 25) ? -d : d;
        }
        const int j = c - 32;
        const double* L = LIN + (size_t)k * LIN_SIZE;
        const double pd = L[LIN_PD + j];
        const double* pg = L + LIN_PG + j * DOF;
        double s = 0;
#pragma unroll
        for (int m = 0; m < DOF; m++) {
            const double g = pg[m];
            s += (pd * g * Tx(m)) * z[m] + (-g * Tu(m)) * z[NX + m];
        }
        return s;
    }
    __device__ __forceinline__ double hval(int k, int c) const {
        const double* L = LIN + (size_t)k * LIN_SIZE;
        if (c < 9) return -L[LIN_XLO + c];
        if (c < 18) return L[LIN_XHI + c - 9];
        if (c < 25) return -L[LIN_DLO + c - 18];
        if (c < 32) return L[LIN_DHI + c - 25];
        return L[LIN_PRHS + c - 32];
    }

    // ---- gradient of the step QP: g = H z + f + G' mu.  Writes g(mu = IV) to G (global); optionally also
    //      g(mu = ILAM) to dst0 (shared) for the dual residual.
    __device__ void gradient(double* dst0) const {
        const int tot = S * HZ;
        for (int o = lane; o < tot; o += 32) {
            const int k = o / HZ, r = o - k * HZ;
            const double* L = LIN + (size_t)k * LIN_SIZE;
            const double* z = VAR + k * HZ;
            const double* mv = IV + (size_t)k * NINEQ;
            const double* ml = ILAM + (size_t)k * NINEQ;
            double base = 0, sv = 0, sl = 0;
            if (r < NX) {
                base = L[LIN_q + r];
#pragma unroll
                for (int c = 0; c < NX; c++) base += L[LIN_Q + ((r >= c) ? sym9(r, c) : sym9(c, r))] * z[c];
                sv = mv[9 + r] - mv[r];
                if (dst0) sl = ml[9 + r] - ml[r];
                if (k < N && r < DOF) {
                    const double tx = Tx(r);
#pragma unroll
                    for (int j = 0; j < NPOLY; j++) {
                        const double a = L[LIN_PD + j] * L[LIN_PG + j * DOF + r] * tx;
                        sv += mv[32 + j] * a;
                        if (dst0) sl += ml[32 + j] * a;
                    }
                }
            } else if (k < N) {
                const int j = r - NX;
                base = L[LIN_RD + j] * z[r] + L[LIN_r + j];
                if (j < DOF) {
                    if (k >= 1) base += dyn.cpl[j] * z[r - HZ];
                    if (k <= N - 2) base += dyn.cpl[j] * z[r + HZ];
                    sv = mv[25 + j] - mv[18 + j];
                    if (dst0) sl = ml[25 + j] - ml[18 + j];
                    if (k + 1 <= N - 1) {
                        sv -= mv[NINEQ + 25 + j] - mv[NINEQ + 18 + j];
                        if (dst0) sl -= ml[NINEQ + 25 + j] - ml[NINEQ + 18 + j];
                    }
                    const double tu = Tu(j);
#pragma unroll
                    for (int jj = 0; jj < NPOLY; jj++) {
                        const double a = -L[LIN_PG + jj * DOF + j] * tu;
                        sv += mv[32 + jj] * a;
                        if (dst0) sl += ml[32 + jj] * a;
                    }
                }
            }
            G[o] = base + sv;
            if (dst0) dst0[o] = base + sl;
        }
        __syncwarp();
    }

    // ---- Riccati factorisation; false if some M_nunu is not positive definite ----
    __device__ bool factor() const {
        double* Pm = SC + WSC_PM;    // 16 x 16 cost-to-go of stage k+1: rows/cols 0..8 = xi, 9..15 = previous dq step
        double* Mnx = SC + WSC_MNX;  // 8 x 16
        double* Mnn = SC + WSC_MNN;  // 8 x 8
        double* FF = SC + WSC_FF;    // 8 x 9
        double* GS = SC + WSC_GS;    // 11 x 14 polytopic rows (ax | au) of the current stage
        double* WG = SC + WSC_WG;    // 11 barrier weights of the polytopic rows
        // terminal stage: P_N = Q_N + box W
        {
            const double* L = LIN + (size_t)N * LIN_SIZE;
            const double* w = IW + (size_t)N * NINEQ;
            for (int e = lane; e < 256; e += 32) {
                const int r = e >> 4, c = e & 15;
                double v = 0;
                if (r < 9 && c < 9) {
                    v = L[LIN_Q + ((r >= c) ? sym9(r, c) : sym9(c, r))];
                    if (r == c) v += w[r] + w[9 + r];
                }
                Pm[e] = v;
            }
        }
        __syncwarp();
        bool ok = true;
        for (int k = N - 1; k >= 0; k--) {
            const double* L = LIN + (size_t)k * LIN_SIZE;
            const double* w = IW + (size_t)k * NINEQ;
            // stage the polytopic rows and their weights
            for (int e = lane; e < NPOLY * 14; e += 32) {
                const int j = e / 14, a = e - j * 14;
                const double g = L[LIN_PG + j * DOF + (a < 7 ? a : a - 7)];
                GS[e] = (a < 7) ? L[LIN_PD + j] * g * Tx(a) : -g * Tu(a - 7);
            }
            if (lane < NPOLY) WG[lane] = w[32 + lane];
            // FF = B'Pxx + E'Pwx  (8 x 9)
            for (int e = lane; e < 72; e += 32) {
                const int i = e / 9, c = e - i * 9;
                FF[e] = (i < 7) ? dyn.bq[i] * Pm[i * 16 + c] + Pm[(9 + i) * 16 + c] : dyn.bs * Pm[7 * 16 + c] + dyn.bv * Pm[8 * 16 + c];
            }
            __syncwarp();
            // Mnn (8 x 8)
            for (int e = lane; e < 64; e += 32) {
                const int i = e >> 3, j = e & 7;
                double v;
                if (j < 7) {
                    v = dyn.bq[j] * FF[i * 9 + j];
                    if (i < 7) v += dyn.bq[i] * Pm[(9 + j) * 16 + i] + Pm[(9 + i) * 16 + 9 + j];
                    else v += dyn.bs * Pm[(9 + j) * 16 + 7] + dyn.bv * Pm[(9 + j) * 16 + 8];
                } else {
                    v = dyn.bs * FF[i * 9 + 7] + dyn.bv * FF[i * 9 + 8];
                }
                if (i == j) { v += L[LIN_RD + j]; if (j < 7) v += w[18 + j] + w[25 + j]; }
#pragma unroll
                for (int p = 0; p < NPOLY; p++) if (i < 7 && j < 7) v += WG[p] * GS[p * 14 + 7 + i] * GS[p * 14 + 7 + j];
                Mnn[e] = v;
            }
            // Mnx (8 x 16) = [FF A + poly | (cpl - wr) diag]
            for (int e = lane; e < 128; e += 32) {
                const int i = e >> 4, c = e & 15;
                double v = 0;
                if (c < 9) {
                    v = FF[i * 9 + c];
                    if (c == 8) v += dyn.asv * FF[i * 9 + 7];
                    if (i < 7 && c < 7) {
#pragma unroll
                        for (int p = 0; p < NPOLY; p++) v += WG[p] * GS[p * 14 + 7 + i] * GS[p * 14 + c];
                    }
                } else if (c - 9 == i && k >= 1) {
                    v = dyn.cpl[i] - (w[18 + i] + w[25 + i]);
                }
                Mnx[e] = v;
            }
            __syncwarp();  // everybody has read P_{k+1}; overwrite it with [Mxx 0; 0 Mww]
            {
                double nv[8];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const int e = lane + 32 * t, r = e >> 4, c = e & 15;
                    double v = 0;
                    if (r < 9 && c < 9) {
                        v = Pm[e];
                        if (c == 8) v += dyn.asv * Pm[r * 16 + 7];
                        if (r == 8) v += dyn.asv * (Pm[7 * 16 + c] + ((c == 8) ? dyn.asv * Pm[7 * 16 + 7] : 0.0));
                        v += L[LIN_Q + ((r >= c) ? sym9(r, c) : sym9(c, r))];
                        if (r == c && k >= 1) v += w[r] + w[9 + r];
                        if (r < 7 && c < 7) {
#pragma unroll
                            for (int p = 0; p < NPOLY; p++) v += WG[p] * GS[p * 14 + r] * GS[p * 14 + c];
                        }
                    } else if (r == c && k >= 1) {
                        v = w[18 + (r - 9)] + w[25 + (r - 9)];
                    }
                    nv[t] = v;
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; t++) Pm[lane + 32 * t] = nv[t];
            }
            // Cholesky Mnn = L L' in place (lower triangle), right-looking; 1/L_jj kept in registers
            double inv_d[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                __syncwarp();
                const double d = Mnn[j * 8 + j];
                if (!(d > 0.0)) ok = false;
                const double inv = rsqrt(d);
                inv_d[j] = inv;
                // trailing update of the lower triangle below/right of j; reads column j (not written here)
#pragma unroll
                for (int rnd = 0; rnd < 2; rnd++) {
                    const int ee = lane + 32 * rnd, ii = ee >> 3, cc = ee & 7;
                    if (ii > j && cc > j && cc <= ii) Mnn[ee] -= (Mnn[ii * 8 + j] * inv) * (Mnn[cc * 8 + j] * inv);
                }
                __syncwarp();
                if (lane > j && lane < 8) Mnn[lane * 8 + j] *= inv;
                if (lane == j) Mnn[j * 8 + j] = d * inv;
            }
            __syncwarp();
            // Lam = L^-1 Mnx (8 x 16): one column per lane (lanes 0..15)
            if (lane < 16) {
                double col[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    double s = Mnx[i * 16 + lane];
#pragma unroll
                    for (int t = 0; t < i; t++) s -= Mnn[i * 8 + t] * col[t];
                    col[i] = s * inv_d[i];
                }
#pragma unroll
                for (int i = 0; i < 8; i++) Mnx[i * 16 + lane] = col[i];
            }
            __syncwarp();
            // P_k = [Mxx 0; 0 Mww] - Lam' Lam
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const int e = lane + 32 * t, r = e >> 4, c = e & 15;
                double s = Pm[e];
#pragma unroll
                for (int q = 0; q < 8; q++) s -= Mnx[q * 16 + r] * Mnx[q * 16 + c];
                Pm[e] = s;
            }
            // store the factor of this stage
            double* F = FACT + (size_t)k * WF_SIZE;
            for (int e = lane; e < 36; e += 32) {
                int i = 0;
                while ((i + 1) * (i + 2) / 2 <= e) i++;
                const int j = e - i * (i + 1) / 2;
                F[WF_L + e] = Mnn[i * 8 + j];
            }
            if (lane < 8) {
                double v = inv_d[0];
#pragma unroll
                for (int j = 1; j < 8; j++) if (lane == j) v = inv_d[j];
                F[WF_INV + lane] = v;
            }
            for (int e = lane; e < 128; e += 32) F[WF_LAM + e] = Mnx[e];
            __syncwarp();
        }
        return !wany(!ok);
    }

    // ---- Riccati vector sweeps: Newton step for the gradient in G -> STEP (shared), KAP (global) ----
    __device__ void solve_step() const {
        double* FS = SC + WSC_PM;    // staged factor record of the current stage (172 doubles)
        double* pv = SC + WSC_VEC;   // p = [px(9); pw(7)]
        double* mn = SC + WSC_VEC + 16;
        double* dv = SC + WSC_VEC + 32;  // forward: d = [dxi(9); dw(7)]
        if (lane < 16) pv[lane] = (lane < 9) ? G[(size_t)N * HZ + lane] : 0.0;
        __syncwarp();
        for (int k = N - 1; k >= 0; k--) {
            const double* F = FACT + (size_t)k * WF_SIZE;
            for (int e = lane; e < WF_SIZE; e += 32) FS[e] = F[e];
            const double* g = G + (size_t)k * HZ;
            if (lane < 8) {
                const int i = lane;
                mn[i] = (i < 7) ? g[NX + i] + dyn.bq[i] * pv[i] + pv[9 + i] : g[NX + 7] + dyn.bs * pv[7] + dyn.bv * pv[8];
            }
            double mx = 0;
            if (lane < 9) { mx = g[lane] + pv[lane]; if (lane == 8) mx += dyn.asv * pv[7]; }
            __syncwarp();
            // kappa = L^-1 mn (every lane, redundantly)
            double kap[8];
            {
                int q = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    double s = mn[i];
#pragma unroll
                    for (int t = 0; t < i; t++) s -= FS[WF_L + q++] * kap[t];
                    q++;
                    kap[i] = s * FS[WF_INV + i];
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int t = 0; t < 8; t++) mx -= FS[WF_LAM + 16 * t + lane] * kap[t];
            }
            __syncwarp();
            if (lane < 16) pv[lane] = mx;
            if (lane < 8) {
                double v = kap[0];
#pragma unroll
                for (int i = 1; i < 8; i++) if (lane == i) v = kap[i];
                KAP[(size_t)k * 8 + lane] = v;
            }
            __syncwarp();
        }
        // forward
        if (lane < 16) dv[lane] = 0.0;
        if (lane < NX) STEP[lane] = 0.0;
        __syncwarp();
        for (int k = 0; k < N; k++) {
            const double* F = FACT + (size_t)k * WF_SIZE;
            for (int e = lane; e < WF_SIZE; e += 32) FS[e] = F[e];
            __syncwarp();
            if (lane < 8) {
                double s = KAP[(size_t)k * 8 + lane];
#pragma unroll
                for (int c = 0; c < 16; c++) s += FS[WF_LAM + 16 * lane + c] * dv[c];
                mn[lane] = -s;
            }
            __syncwarp();
            // dn = L^-T rhs (every lane, redundantly)
            double dn[8];
#pragma unroll
            for (int i = 7; i >= 0; i--) {
                double s = mn[i];
#pragma unroll
                for (int t = i + 1; t < 8; t++) s -= FS[WF_L + t * (t + 1) / 2 + i] * dn[t];
                dn[i] = s * FS[WF_INV + i];
            }
            double nx = 0;
            double dsel = dn[0];  // dn[lane], selected with compile-time indices to keep dn in registers
#pragma unroll
            for (int i = 1; i < 8; i++) if (lane == i) dsel = dn[i];
            if (lane < 7) nx = dv[lane] + dyn.bq[lane] * dsel;
            else if (lane == 7) nx = dv[7] + dyn.asv * dv[8] + dyn.bs * dn[7];
            else if (lane == 8) nx = dv[8] + dyn.bv * dn[7];
            __syncwarp();
            if (lane < 8) STEP[k * HZ + NX + lane] = dsel;
            if (lane < 9) { dv[lane] = nx; STEP[(k + 1) * HZ + lane] = nx; }
            else if (lane < 16) {
                double v = dn[0];
#pragma unroll
                for (int i = 1; i < 7; i++) if (lane - 9 == i) v = dn[i];
                dv[lane] = v;
            }
            __syncwarp();
        }
        if (lane < NU) STEP[N * HZ + NX + lane] = 0.0;
        __syncwarp();
    }

    // slack / multiplier steps from the primal step; largest step keeping t, lam > 0
    __device__ double ineq_steps() const {
        double a = 1.0;
        const int tot = S * NINEQ;
        for (int i = lane; i < tot; i += 32) {
            const int k = i / NINEQ, c = i - k * NINEQ;
            if (!present(N, k, c)) continue;
            const double g = gz(STEP, k, c);
            const double dt = -IRP[i] - g;
            const double dl = -ILAM[i] + IV[i] + IW[i] * g;
            IDT[i] = dt; IDLAM[i] = dl;
            if (dt < 0) a = fmin(a, -IT[i] / dt);
            if (dl < 0) a = fmin(a, -ILAM[i] / dl);
        }
        __syncwarp();
        return wmin(a);
    }

    // ---- interior-point loop; on success VAR holds the step (xi exact rollout of nu) ----
    __device__ QpStats solve() const {
        QpStats st;
        st.ok = 0; st.iters = 0; st.res_dual = 0; st.res_prim = 0; st.gap = 0;
        const int tot = S * NINEQ;
        // feasibility of the boxes (stage 0: xi_0 = 0 must lie inside; others: lo <= hi)
        {
            bool bad = false;
            for (int o = lane; o < S * NX; o += 32) {
                const int k = o / NX, m = o - k * NX;
                const double* L = LIN + (size_t)k * LIN_SIZE;
                if (k == 0) { if (L[LIN_XLO + m] > 1e-9 || L[LIN_XHI + m] < -1e-9) bad = true; }
                else if (L[LIN_XLO + m] > L[LIN_XHI + m]) bad = true;
            }
            if (wany(bad)) return st;
        }
        // initial point: nu = 0, xi = rollout of the defects, t = max(h - Gz, 1), lam = 1
        double qn = 0;
        {
            double x = 0;
            for (int k = 0; k <= N; k++) {
                const double* L = LIN + (size_t)k * LIN_SIZE;
                if (lane < NX) { VAR[k * HZ + lane] = x; qn = fmax(qn, fabs(L[LIN_q + lane])); }
                else if (lane < HZ) { VAR[k * HZ + lane] = 0.0; if (k < N) qn = fmax(qn, fabs(L[LIN_r + lane - NX])); }
                const double x8 = __shfl_sync(0xffffffffu, x, 8);
                if (k < N && lane < NX) { x += L[LIN_b + lane]; if (lane == 7) x += dyn.asv * x8; }
            }
            qn = wmax(qn);
        }
        __syncwarp();
        for (int i = lane; i < tot; i += 32) {
            const int k = i / NINEQ, c = i - k * NINEQ;
            if (!present(N, k, c)) { IT[i] = 1; ILAM[i] = 0; IW[i] = 0; IV[i] = 0; IRP[i] = 0; continue; }
            IT[i] = fmax(hval(k, c) - gz(VAR, k, c), 1.0);
            ILAM[i] = 1.0;
        }
        __syncwarp();
        const double m_tot = 43.0 * N;
        for (int it = 0; it < opt.max_iter; it++) {
            // residuals, barrier weights, predictor v = lam rp / t
            double mu = 0, nrp = 0;
            for (int i = lane; i < tot; i += 32) {
                const int k = i / NINEQ, c = i - k * NINEQ;
                if (!present(N, k, c)) continue;
                const double t = IT[i], lam = ILAM[i];
                const double rp = gz(VAR, k, c) + t - hval(k, c);
                IRP[i] = rp;
                nrp = fmax(nrp, fabs(rp));
                mu += t * lam;
                const double w = lam / t;
                IW[i] = w;
                IV[i] = w * rp;
            }
            __syncwarp();
            mu = wsum(mu) / m_tot;
            nrp = wmax(nrp);
            gradient(STEP);  // G <- predictor gradient; STEP <- Lagrangian gradient (scratch until the step is computed)
            // costates by the xi-stationarity recursion and the nu-stationarity residual
            double nrd = 0;
            {
                double y = 0;  // lane m < 9 holds y_{k}[m]
                for (int k = N; k >= 1; k--) {
                    const double y7 = __shfl_sync(0xffffffffu, y, 7);
                    if (lane < NX) { y += STEP[k * HZ + lane]; if (lane == 8) y += dyn.asv * y7; }
                    const double y8 = __shfl_sync(0xffffffffu, y, 8);
                    if (lane < 7) nrd = fmax(nrd, fabs(STEP[(k - 1) * HZ + NX + lane] + dyn.bq[lane] * y));
                    else if (lane == 7) nrd = fmax(nrd, fabs(STEP[(k - 1) * HZ + NX + 7] + dyn.bs * y + dyn.bv * y8));
                }
                nrd = wmax(nrd);
            }
            __syncwarp();
            st.iters = it; st.res_dual = nrd; st.res_prim = nrp; st.gap = mu;
            if (nrd <= opt.eps * (1.0 + qn) && nrp <= opt.eps && mu <= opt.eps) { st.ok = 1; break; }
            if (!(nrd == nrd) || !(mu == mu)) break;
            if (!factor()) break;
            solve_step();
            const double a_aff = ineq_steps();
            double mu_aff = 0;
            for (int i = lane; i < tot; i += 32) {
                const int k = i / NINEQ, c = i - k * NINEQ;
                if (!present(N, k, c)) continue;
                mu_aff += (IT[i] + a_aff * IDT[i]) * (ILAM[i] + a_aff * IDLAM[i]);
            }
            mu_aff = wsum(mu_aff) / m_tot;
            const double sigma = (mu > 0) ? (mu_aff / mu) * (mu_aff / mu) * (mu_aff / mu) : 0.0;
            // corrector: v = (lam rp + sigma mu - dt_a dlam_a) / t
            for (int i = lane; i < tot; i += 32) {
                const int k = i / NINEQ, c = i - k * NINEQ;
                if (!present(N, k, c)) continue;
                IV[i] = (ILAM[i] * IRP[i] + sigma * mu - IDT[i] * IDLAM[i]) / IT[i];
            }
            __syncwarp();
            gradient(nullptr);
            solve_step();
            const double a = fmin(1.0, 0.995 * ineq_steps());
            for (int o = lane; o < S * HZ; o += 32) VAR[o] += a * STEP[o];
            for (int i = lane; i < tot; i += 32) {
                const int k = i / NINEQ, c = i - k * NINEQ;
                if (!present(N, k, c)) continue;
                IT[i] += a * IDT[i];
                ILAM[i] += a * IDLAM[i];
            }
            __syncwarp();
            st.iters = it + 1;
        }
        if (st.ok) {
            // make the equalities exact: xi = rollout(nu)
            double x = 0;
            for (int k = 0; k <= N; k++) {
                const double* L = LIN + (size_t)k * LIN_SIZE;
                if (lane < NX) VAR[k * HZ + lane] = x;
                const double x8 = __shfl_sync(0xffffffffu, x, 8);
                if (k < N && lane < NX) {
                    const double nu7 = VAR[k * HZ + NX + 7];
                    if (lane < 7) x = x + dyn.bq[lane] * VAR[k * HZ + NX + lane] + L[LIN_b + lane];
                    else if (lane == 7) x = x + dyn.asv * x8 + dyn.bs * nu7 + L[LIN_b + 7];
                    else x = x + dyn.bv * nu7 + L[LIN_b + 8];
                }
            }
            __syncwarp();
        }
        return st;
    }

    // ---- horizon evaluation, lane = stage.  FULL: fill LIN; else objective / violation of guess + alpha T step ----
    template <bool FULL>
    __device__ void eval_horizon(const double* cur_u, const double* rb, size_t rb_stride, size_t rb_stage, double alpha, double& obj, double& gap) const {
        double o_acc = 0, g_acc = 0;
        for (int k = lane; k <= N; k += 32) {
            double x[NX], u[NU], up[DOF], un[DOF], xn[NX];
            auto gx = [&](int kk, int e) -> double {
                double v = GUESS[kk * HZ + e];
                if (!FULL) {
                    if (e < NX) v += alpha * (Tx(e) * SSTEP[kk * HZ + e]);
                    else v = (kk < N) ? v + alpha * (Tu(e - NX) * SSTEP[kk * HZ + e]) : 0.0;
                }
                return v;
            };
#pragma unroll
            for (int e = 0; e < NX; e++) x[e] = gx(k, e);
#pragma unroll
            for (int e = 0; e < NU; e++) u[e] = gx(k, NX + e);
#pragma unroll
            for (int j = 0; j < DOF; j++) {
                up[j] = (k == 0) ? cur_u[j] : gx(k - 1, NX + j);
                un[j] = (k < N) ? gx(k + 1, NX + j) : 0.0;
            }
#pragma unroll
            for (int e = 0; e < NX; e++) xn[e] = (k < N) ? gx(k + 1, e) : 0.0;
            StageLin sl;
            RbView rv{rb + (size_t)k * rb_stage, rb_stride};
            stage_eval<FULL>(P, T, Ts, N, k, x, u, up, un, xn, rv, sl);
            o_acc += sl.obj; g_acc += sl.gap;
            if (FULL) {
                double* L = LIN + (size_t)k * LIN_SIZE;
                const double* src = (const double*)&sl;
                for (int e = 0; e < LIN_SIZE; e++) L[e] = src[e];
            }
        }
        __syncwarp();
        obj = wsum(o_acc); gap = wsum(g_acc);
    }

    // ---- the SQP loop (solveOCP) ----
    __device__ SqpResult run(const double* cur_u, const double* rb, size_t rb_stride, size_t rb_stage, SqpLogRef* log) {
        SqpResult res;
        res.status = SOLVED; res.iters = 0; res.qp_fail = 0; res.qp_iters = 0; res.accept_mask = 0;
        const int max_iter = (int)P.max_iter, ls_max = (int)P.line_search_max_iter;
        const int HN = S * HZ;
        for (int e = lane; e < HN; e += 32) SSTEP[e] = 0.0;
        __syncwarp();
        int n_filt = 0, it = 0;
        bool done = false;
        for (it = 0; it < max_iter; it++) {
            double obj, gap;
            eval_horizon<true>(cur_u, rb, rb_stride, rb_stage, 0.0, obj, gap);
            // mis-indexed input-bound rows (osqp_interface.cpp:273) intersected into the state boxes
            for (int c = lane; c < NU * N; c += 32) {
                const int k = c / NX, m = c - k * NX, i = c / NU, kk = c - i * NU;
                const double uv = GUESS[i * HZ + NX + kk];
                const double lo = (P.lu[kk] - uv) / Tu(kk), hi = (P.uu[kk] - uv) / Tu(kk);
                double* L = LIN + (size_t)k * LIN_SIZE;
                L[LIN_XLO + m] = fmax(L[LIN_XLO + m], lo);
                L[LIN_XHI + m] = fmin(L[LIN_XHI + m], hi);
            }
            __syncwarp();
            // isPosdef / isNan on the block structure of the Hessian (osqp_interface.cpp:454-473)
            {
                bool pd = true, nan = false;
                for (int k = lane; k <= N; k += 32) {
                    const double* L = LIN + (size_t)k * LIN_SIZE;
                    double A[81];
                    for (int r = 0; r < 9; r++) for (int c = 0; c <= r; c++) { A[9 * r + c] = L[LIN_Q + sym9(r, c)]; if (A[9 * r + c] != A[9 * r + c]) nan = true; }
                    for (int j = 0; j < 9 && pd; j++) {
                        double d = A[10 * j];
                        for (int t = 0; t < j; t++) d -= A[9 * j + t] * A[9 * j + t];
                        if (d <= 0.0) { pd = false; break; }
                        d = sqrt(d);
                        A[10 * j] = d;
                        for (int i = j + 1; i < 9; i++) {
                            double s = A[9 * i + j];
                            for (int t = 0; t < j; t++) s -= A[9 * i + t] * A[9 * j + t];
                            A[9 * i + j] = s / d;
                        }
                    }
                }
                if (lane < NU) {
                    const int j = lane;
                    double d = 0;
                    for (int k = 0; k < N; k++) {
                        const double rd = LIN[(size_t)k * LIN_SIZE + LIN_RD + j];
                        if (rd != rd) nan = true;
                        d = (k == 0 || j == 7) ? rd : rd - dyn.cpl[j] * dyn.cpl[j] / d;
                        if (d <= 0.0) { pd = false; break; }
                    }
                }
                pd = !wany(!pd); nan = wany(nan);
                if (!pd) { res.status = NON_PD_HESSIAN; done = true; break; }
                if (nan) { res.status = NAN_HESSIAN; done = true; break; }
            }
            QpStats qs = solve();
            res.qp_iters += qs.iters;
            if (qs.ok) {
                for (int e = lane; e < HN; e += 32) { const int k = e / HZ, r = e - k * HZ; SSTEP[e] = (r < NX || k < N) ? VAR[e] : 0.0; }
            } else {
                res.qp_fail++;  // step keeps its previous value (osqp_interface.cpp:479-505)
            }
            __syncwarp();
            // ---- filterLineSearch (osqp_interface.cpp:759-808) ----
            double alpha = 1.0;
            bool accepted = true;  // never reset inside the loop (:767)
            for (int i = 0; i < ls_max; i++) {
                if (accepted) {  // once a trial is rejected no later trial can be accepted: their evaluation is dead work
                    double o2, g2;
                    eval_horizon<false>(cur_u, rb, rb_stride, rb_stage, alpha, o2, g2);
                    bool dom = false;
                    for (int j = lane; j < n_filt; j += 32) if (o2 >= FILT[2 * j] && g2 >= FILT[2 * j + 1]) dom = true;
                    if (wany(dom)) accepted = false;
                    if (accepted) {
                        if (lane == 0) {
                            int w = 0;
                            for (int j = 0; j < n_filt; j++)
                                if (o2 > FILT[2 * j] || g2 > FILT[2 * j + 1]) { FILT[2 * w] = FILT[2 * j]; FILT[2 * w + 1] = FILT[2 * j + 1]; w++; }
                            FILT[2 * w] = o2; FILT[2 * w + 1] = g2;
                            n_filt = w + 1;
                        }
                        n_filt = __shfl_sync(0xffffffffu, n_filt, 0);
                        __syncwarp();
                        break;
                    }
                }
                alpha *= P.line_search_tau;
            }
            if (accepted && it < 32) res.accept_mask |= (1u << it);
            // ---- take the step (osqp_interface.cpp:549-551) ----
            double inf = 0;
            for (int e = lane; e < HN; e += 32) {
                const int k = e / HZ, r = e - k * HZ;
                const double s = SSTEP[e];
                if (r < NX) { GUESS[e] += alpha * (Tx(r) * s); inf = fmax(inf, fabs(s)); }
                else if (k < N) { GUESS[e] += alpha * (Tu(r - NX) * s); inf = fmax(inf, fabs(s)); }
                else GUESS[e] = 0.0;
            }
            __syncwarp();
            inf = wmax(inf);
            if (log && log->n < log->max_log) {
                if (log->steps) for (int e = lane; e < HN; e += 32) log->steps[(size_t)log->n * HN + e] = SSTEP[e];
                if (lane == 0) { log->alphas[log->n] = alpha; log->qp_ok[log->n] = qs.ok; }
                log->n++;
            }
            if (alpha * inf < P.eps_prim) { res.status = SOLVED; res.iters = it + 1; done = true; break; }
        }
        if (!done) { res.status = MAX_ITER_EXCEEDED; res.iters = max_iter; }
        else if (res.status != SOLVED) res.iters = it;
        return res;
    }
};

}  // namespace mpcc
#endif  // __CUDACC__
