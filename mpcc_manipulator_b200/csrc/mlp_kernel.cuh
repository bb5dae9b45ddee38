// Batched self-/environment-collision distance networks with forward-mode joint tangents, fp64.
// Replaces SelCollNNmodel / EnvCollNNmodel::calculateMlpOutput (reference
// cpp/src/Constraints/SelfCollision/SelfCollisionModel.cpp:140-250 and its Env clone) and
// RobotData::update / updateEnv (cpp/include/Model/robot_data.h:55-88) for every (instance, stage).
//
// Formulation.  For one sample the network input is z = [x, sin x, cos x]; the value h_l and the 7
// joint tangents T_l = d h_l / d q travel together as 8 columns  X_l = [h_l | T_l]  and every layer is
//     X_{l+1} = relu-mask( W_l X_l + [b_l | 0] ),   mask = 1[pre-activation of the value column > 0]
// (ReLU'(0) = 0 as SelfCollisionModel.h:66-69).  Only the 7 joint columns are propagated (the
// reference propagates 10 for the env net and discards 3, robot_data.h:85) and the encoding Jacobian
// is applied in its diagonal form -- same numbers, 1.746 MMAC instead of 2.447 MMAC per sample.
//
// Mapping.  One persistent CTA per SM walks over tiles of 8 samples = 64 columns.  The activation
// tile X (256 x 64 doubles, 128 KB) stays in shared memory across all layers of both networks; the
// weights stream from L2 through a double-buffered 2 x 32 KB shared-memory ring with cp.async, in
// chunks pre-packed on the host in exactly the order the threads read them.  Each thread owns an
// 8 (neurons) x 8 (columns of ONE sample) register tile, so the ReLU mask is thread-local.
// The work is a dense fp64 contraction: tcgen05 has no f64 kind, so the DFMA pipe is the roof.
#pragma once
#include "mpcc_types.h"
#include <cuda_runtime.h>

namespace mpcc {

constexpr int MLP_THREADS = 256;
constexpr int MLP_TILE_S = 8;                 // samples per tile
constexpr int MLP_KC = 16;                    // k-steps per weight chunk
constexpr int MLP_CHUNK_D = MLP_KC * 256;     // doubles per chunk (32 KB)
// chunk sequence of one tile: env L0 (2) L1 L2 L3 (16 each) | self L0 (2) L1 (4, K split over 4 thread groups)
constexpr int MLP_NCHUNK = 2 + 48 + 2 + 4;    // 56
constexpr size_t MLP_SMEM_BYTES = (size_t)(256 * 64 + 2 * MLP_CHUNK_D + 9 * 256) * sizeof(double);  // 215040: X tile | weight ring | env output layer

struct MlpArgs {
    const double* wpack;      // MLP_NCHUNK * MLP_CHUNK_D doubles, packed by pack_mlp_weights()
    const double* bias;       // env b0..b3 (4*256) | env b4 (9) | self b0 (256) | self b1 (64) | self b2 (1)
    const double* w_out_env;  // 9 x 256 row-major
    const double* w_out_self; // 1 x 64
    const double* qs;         // [7][NS] joint angles of every sample (SoA)
    const double* obs;        // [B][4] obstacle x,y,z,radius per instance
    double* rb;               // [RB_DOUBLES][NS] RobotData (SoA), this kernel fills sel/dsel/obs_r/env/denv
    int NS;                   // samples = instances * (N+1)
    int S;                    // stages per instance (N+1)
    int n_tiles;
};

constexpr int MLP_BIAS_ENV = 0, MLP_BIAS_ENV_OUT = 1024, MLP_BIAS_SELF0 = 1033, MLP_BIAS_SELF1 = 1289, MLP_BIAS_SELF_OUT = 1353, MLP_BIAS_TOTAL = 1354;

#if defined(__CUDACC__)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// one 16-k-step chunk of the 8x8 register-tile contraction.
//   Wb: chunk buffer as double2; thread reads Wb[(kk*WSTR_K + wofs) + i*WSTR_I]  for i = 0..3
//   Xs: activation tile as double2; thread reads Xs[((k0+kk)*4 + c4)*8 + tx] for c4 = 0..3
template <int WSTR_K, int WSTR_I>
__device__ __forceinline__ void mlp_chunk(const double2* __restrict__ Wb, int wofs, const double2* __restrict__ Xs, int k0, int tx,
                                          int ksteps, double (&acc)[8][8]) {
#pragma unroll 4
    for (int kk = 0; kk < ksteps; kk++) {
        double2 w2[4], x2[4];
#pragma unroll
        for (int i = 0; i < 4; i++) w2[i] = Wb[kk * WSTR_K + wofs + i * WSTR_I];
#pragma unroll
        for (int c = 0; c < 4; c++) x2[c] = Xs[((k0 + kk) * 4 + c) * 8 + tx];
        double w[8] = {w2[0].x, w2[0].y, w2[1].x, w2[1].y, w2[2].x, w2[2].y, w2[3].x, w2[3].y};
        double x[8] = {x2[0].x, x2[0].y, x2[1].x, x2[1].y, x2[2].x, x2[2].y, x2[3].x, x2[3].y};
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[r][c] = fma(w[r], x[c], acc[r][c]);
    }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) k_mlp(MlpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* Xs = reinterpret_cast<double2*>(smem_raw);                              // [256 k][4 c4][8 tx]
    double2* Wbuf = reinterpret_cast<double2*>(smem_raw + 256 * 64 * sizeof(double));  // 2 x [2048 double2]
    double* Xd = reinterpret_cast<double*>(Xs);
    double* Wout = reinterpret_cast<double*>(smem_raw + (256 * 64 + 2 * MLP_CHUNK_D) * sizeof(double));  // env output layer 9 x 256, resident
    for (int i = threadIdx.x; i < 9 * 256; i += MLP_THREADS) Wout[i] = a.w_out_env[i];  // visible after the first __syncthreads() below

    const int tid = threadIdx.x;
    const int tx = tid & 7;         // sample within the tile
    const int ty = tid >> 3;        // mode A: row group 0..31
    const int grp = tid >> 6;       // mode B: K group 0..3
    const int tyb = (tid & 63) >> 3;  // mode B: row group 0..7

    int p = 0;    // position in the chunk sequence
    int buf = 0;  // ring slot holding chunk p
    auto prefetch = [&](int chunk, int slot) {
        const double2* src = reinterpret_cast<const double2*>(a.wpack) + (size_t)chunk * (MLP_CHUNK_D / 2);
        double2* dst = Wbuf + slot * (MLP_CHUNK_D / 2);
#pragma unroll
        for (int i = 0; i < (MLP_CHUNK_D / 2) / MLP_THREADS; i++) cp_async16(dst + tid + i * MLP_THREADS, src + tid + i * MLP_THREADS);
        cp_async_commit();
    };
    if ((int)blockIdx.x < a.n_tiles) prefetch(0, 0);

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int n = tile * MLP_TILE_S + tx;  // this thread's sample
        const bool live = n < a.NS;
        double acc[8][8];

        for (int net = 0; net < 2; net++) {  // 0: env, 1: self
            // ---- stage the encoded input X0 (32 rows) : row k = ty of sample tx ----
            __syncthreads();  // previous users of Xs are done
            {
                const int nin = (net == 0) ? 10 : 7;  // raw inputs; encoded rows = 3 * nin
                const int k = ty;
                double v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = 0.0;
                if (live && k < 3 * nin) {
                    const int src = k % nin, kind = k / nin;  // kind 0: x, 1: sin x, 2: cos x
                    double xin;
                    if (src < 7) xin = a.qs[(size_t)src * a.NS + n];
                    else xin = a.obs[(size_t)(n / a.S) * 4 + (src - 7)];
                    double sn, cs;
                    sincos(xin, &sn, &cs);
                    v[0] = (kind == 0) ? xin : (kind == 1 ? sn : cs);
                    if (src < 7) v[1 + src] = (kind == 0) ? 1.0 : (kind == 1 ? cs : -sn);
                }
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) Xs[(k * 4 + c4) * 8 + tx] = make_double2(v[2 * c4], v[2 * c4 + 1]);
            }
            const int n_hidden = (net == 0) ? 4 : 1;  // mode-A layers producing 256 neurons
            for (int layer = 0; layer < n_hidden; layer++) {
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[r][c] = 0.0;
                const int nch = (layer == 0) ? 2 : 16;
                for (int ch = 0; ch < nch; ch++) {
                    cp_async_wait_all();
                    __syncthreads();
                    prefetch((p + 1) % MLP_NCHUNK, buf ^ 1);
                    mlp_chunk<4 * 32, 32>(Wbuf + buf * (MLP_CHUNK_D / 2), ty, Xs, ch * MLP_KC, tx, MLP_KC, acc);
                    buf ^= 1;
                    p = (p + 1) % MLP_NCHUNK;
                }
                // bias + ReLU mask, then the tile becomes the next layer's input
                const double* bias = a.bias + ((net == 0) ? (MLP_BIAS_ENV + layer * 256) : MLP_BIAS_SELF0);
                __syncthreads();  // everyone finished reading Xs
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int r = i * 2 + e, row = i * 64 + ty * 2 + e;
                        const double pre = acc[r][0] + bias[row];
                        const bool on = pre > 0.0;
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            double v0 = (c4 == 0) ? pre : acc[r][2 * c4];
                            double v1 = acc[r][2 * c4 + 1];
                            Xs[(row * 4 + c4) * 8 + tx] = on ? make_double2(v0, v1) : make_double2(0.0, 0.0);
                        }
                    }
            }
            if (net == 0) {
                // ---- env output layer: 9 x 256, value + 7 tangents ----
                __syncthreads();
                const int col = tid & 63, smp = col >> 3, cc = col & 7, rg = tid >> 6;
                double o[3] = {0.0, 0.0, 0.0};
                const double* xcol = Xd + ((cc >> 1) * 8 + smp) * 2 + (cc & 1);
#pragma unroll 4
                for (int k = 0; k < 256; k++) {
                    const double xv = xcol[k * 64];
                    o[0] = fma(Wout[rg * 256 + k], xv, o[0]);
                    o[1] = fma(Wout[(rg + 4) * 256 + k], xv, o[1]);
                    if (rg == 0) o[2] = fma(Wout[8 * 256 + k], xv, o[2]);
                }
                const int ns = tile * MLP_TILE_S + smp;
                if (ns < a.NS) {
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const int l = rg + 4 * j;
                        if (l < 9 && (j < 2 || rg == 0)) {
                            if (cc == 0) a.rb[(size_t)(RB_ENV + l) * a.NS + ns] = o[j] + a.bias[MLP_BIAS_ENV_OUT + l];
                            else a.rb[(size_t)(RB_DENV + l * 7 + (cc - 1)) * a.NS + ns] = o[j];
                        }
                    }
                    if (tid < 8 && (tile * MLP_TILE_S + tid) < a.NS) {
                        const int n2 = tile * MLP_TILE_S + tid;
                        a.rb[(size_t)RB_OBSR * a.NS + n2] = a.obs[(size_t)(n2 / a.S) * 4 + 3];
                    }
                }
            } else {
                // ---- self layer 1: 64 x 256, K split over the 4 thread groups (64 k each) ----
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[r][c] = 0.0;
                for (int ch = 0; ch < 4; ch++) {
                    cp_async_wait_all();
                    __syncthreads();
                    prefetch((p + 1) % MLP_NCHUNK, buf ^ 1);
                    mlp_chunk<4 * 4 * 8, 8>(Wbuf + buf * (MLP_CHUNK_D / 2), grp * 32 + tyb, Xs, grp * 64 + ch * MLP_KC, tx, MLP_KC, acc);
                    buf ^= 1;
                    p = (p + 1) % MLP_NCHUNK;
                }
                __syncthreads();  // done reading Xs: reuse it as reduction scratch [3][64 threads][64]
                if (grp > 0) {
                    double* dst = Xd + ((size_t)(grp - 1) * 64 + (tid & 63)) * 64;
#pragma unroll
                    for (int r = 0; r < 8; r++)
#pragma unroll
                        for (int c = 0; c < 8; c++) dst[r * 8 + c] = acc[r][c];
                }
                __syncthreads();
                if (grp == 0) {
#pragma unroll
                    for (int g2 = 0; g2 < 3; g2++) {
                        const double* src = Xd + ((size_t)g2 * 64 + tid) * 64;
#pragma unroll
                        for (int r = 0; r < 8; r++)
#pragma unroll
                            for (int c = 0; c < 8; c++) acc[r][c] += src[r * 8 + c];
                    }
                }
                __syncthreads();  // scratch consumed
                if (grp == 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = i * 2 + e, row = i * 16 + tyb * 2 + e;
                            const double pre = acc[r][0] + a.bias[MLP_BIAS_SELF1 + row];
                            const bool on = pre > 0.0;
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) {
                                double v0 = (c4 == 0) ? pre : acc[r][2 * c4];
                                double v1 = acc[r][2 * c4 + 1];
                                Xs[(row * 4 + c4) * 8 + tx] = on ? make_double2(v0, v1) : make_double2(0.0, 0.0);
                            }
                        }
                }
                __syncthreads();
                // ---- self output layer: 1 x 64 ----
                if (tid < 64) {
                    const int smp = tid >> 3, cc = tid & 7;
                    const double* xcol = Xd + ((cc >> 1) * 8 + smp) * 2 + (cc & 1);
                    double o = 0.0;
#pragma unroll 8
                    for (int k = 0; k < 64; k++) o = fma(__ldg(a.w_out_self + k), xcol[k * 64], o);
                    const int ns = tile * MLP_TILE_S + smp;
                    if (ns < a.NS) {
                        if (cc == 0) a.rb[(size_t)RB_SEL * a.NS + ns] = o + a.bias[MLP_BIAS_SELF_OUT];
                        else a.rb[(size_t)(RB_DSEL + cc - 1) * a.NS + ns] = o;
                    }
                }
            }
        }
    }
    cp_async_wait_all();
}

#endif  // __CUDACC__

// Host-side packing of both networks' hidden-layer weights into the chunk stream k_mlp consumes.
//   mode A chunk (256 output rows, 16 k):  [kk][i][ty][e]  = W[i*64 + ty*2 + e][k0 + kk]
//   mode B chunk (64 output rows, K split): [kk][g][i][ty'][e] = W[i*16 + ty'*2 + e][g*64 + ch*16 + kk]
// Layer 0 of each net is zero-padded from 30 / 21 encoded inputs to K = 32.
inline void pack_mlp_weights(const double* const env_W[5], const double* const self_W[3], double* out) {
    auto packA = [&](const double* W, int in_dim, int k_pad, double*& o) {
        for (int k0 = 0; k0 < k_pad; k0 += MLP_KC)
            for (int kk = 0; kk < MLP_KC; kk++)
                for (int i = 0; i < 4; i++)
                    for (int ty = 0; ty < 32; ty++)
                        for (int e = 0; e < 2; e++) {
                            int row = i * 64 + ty * 2 + e, k = k0 + kk;
                            *o++ = (k < in_dim) ? W[(size_t)row * in_dim + k] : 0.0;
                        }
    };
    double* o = out;
    packA(env_W[0], 30, 32, o);
    packA(env_W[1], 256, 256, o);
    packA(env_W[2], 256, 256, o);
    packA(env_W[3], 256, 256, o);
    packA(self_W[0], 21, 32, o);
    for (int ch = 0; ch < 4; ch++)
        for (int kk = 0; kk < MLP_KC; kk++)
            for (int g = 0; g < 4; g++)
                for (int i = 0; i < 4; i++)
                    for (int ty = 0; ty < 8; ty++)
                        for (int e = 0; e < 2; e++) {
                            int row = i * 16 + ty * 2 + e, k = g * 64 + ch * MLP_KC + kk;
                            *o++ = self_W[1][(size_t)row * 256 + k];
                        }
}

}  // namespace mpcc
