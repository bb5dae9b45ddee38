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
// weights stream from L2 through shared memory with cp.async in chunks pre-packed on the host in exactly
// the order the lanes read them.  A warp only ever reads the A fragments of its own neurons, so every
// warp owns a private double-buffered ring (2 x 4 KB) that it fills and waits for by itself: the CTA
// only synchronises where the activation tile changes hands (twice per layer), not once per chunk.
// The contraction runs on the FP64 tensor path, mma.sync.m8n8k4.f64 (DMMA): tcgen05 has no f64 kind,
// and on sm_100 DMMA and DFMA share one pipe of the same peak (tools/probes/fp64_pipes_probe.cu) --
// what DMMA buys is operand traffic: a warp owns 32 neurons x 64 columns as 4 x 8 fragments and needs
// 6 LDS.128 per 32 DMMA (= 256 MAC per lane), where an 8x8 DFMA register tile needs 32 LDS.128 for the
// same 256 DFMA and loses ~20 % of the pipe to them (tools/probes/mlp_loop_probe.cu, mlp_mma_probe.cu).
// An n-fragment is one column KIND (value or tangent j) of the 8 samples, so a lane's accumulators for
// a neuron hold all 8 columns of its two samples and the ReLU mask stays thread-local.
//
// Shared layout of X ("fragment order"): element (k, c, s) = neuron k, column c, sample s sits at double
//     (((k >> 2) * 4 + (c >> 1)) * 32 + slot(k & 3, s)) * 2 + (c & 1),   slot(q, s) = 8 (s >> 1) + ((q + 4 (s & 1) + 2 (s >> 1)) & 7)
// so the B fragments of 4 k-steps x 2 columns are ONE LDS.128 per lane, and both that read (a quarter-warp holds
// q = 0..3 of two adjacent samples) and the epilogue's D-fragment write-back (a quarter-warp holds two q of the four
// samples of one parity) touch eight distinct 16-byte banks.
#pragma once
#include "mpcc_types.h"
#include <cuda_runtime.h>

namespace mpcc {

constexpr int MLP_THREADS = 256;
constexpr int MLP_TILE_S = 8;                 // samples per tile
constexpr int MLP_KC = 16;                    // k-steps per weight chunk of a 256-neuron layer
constexpr int MLP_CHUNK_D = MLP_KC * 256;     // doubles per chunk (32 KB)
// chunk sequence of one tile: env L0 (2) L1 L2 L3 (16 each) | self L0 (2) L1 (4 chunks of 64 k-steps x 64 neurons)
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

// D (8x8) += A (8x4, row) * B (4x8, col) in fp64.  Lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2], D[l >> 2][2 (l & 3) + {0, 1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// double2 slot of X element (k, column pair j, sample s) in the fragment-ordered tile
__device__ __forceinline__ int xslot(int q, int s) { return 8 * (s >> 1) + ((q + 4 * (s & 1) + 2 * (s >> 1)) & 7); }
__device__ __forceinline__ int xl2(int k, int j, int s) { return ((k >> 2) * 4 + j) * 32 + xslot(k & 3, s); }

// one 16-k-step chunk of a 256-neuron layer: the warp's 32 neurons (4 m-fragments) x 64 columns (8 n-fragments).
//   Wc: the warp's slice of the chunk as double2 [kb 4][m-pair 2][lane 32]   Xs: activation tile, kb0 = first 4-k block of the chunk
//   bslot = xslot(lane & 3, lane >> 2): where this lane's B-fragment element sits in a 32-slot row of Xs
__device__ __forceinline__ void mlp_chunk(const double2* __restrict__ Wc, const double2* __restrict__ Xs, int kb0, int lane, int bslot,
                                          double (&acc)[4][8][2]) {
#pragma unroll
    for (int kb = 0; kb < MLP_KC / 4; kb++) {
        const double2 a01 = Wc[(kb * 2 + 0) * 32 + lane], a23 = Wc[(kb * 2 + 1) * 32 + lane];
        double2 b[4];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = Xs[((kb0 + kb) * 4 + j) * 32 + bslot];
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                dmma884(acc[mb][2 * j][0], acc[mb][2 * j][1], a[mb], b[j].x);
                dmma884(acc[mb][2 * j + 1][0], acc[mb][2 * j + 1][1], a[mb], b[j].y);
            }
    }
}

// First layer of a network, one 16-row chunk of its (zero-padded) 32 encoded inputs, WITHOUT the tensor path: the encoded input
// z = [x, sin x, cos x] has a diagonal Jacobian (input j only moves rows j, NIN + j, 2 NIN + j), so the 7 tangent columns of this
// layer cost 3 MAC per neuron and joint instead of the 32 a dense 8-column product spends: 51 (env) / 42 (self) DFMA per neuron and
// sample instead of 256 MAC on DMMA -- 5 x less work on the shared FP64 pipe for 7 % of the kernel's former DMMA count.
//   Wc: the warp's slice of the chunk as [k 16][row 32] (row = 8 mb + fr: the lane's D-fragment rows)    Zs: [k 32][sample 8]
// The accumulators have the D-fragment ownership of mlp_chunk (acc[mb][column][e]: row 32 warp + 8 mb + fr, sample 2 fq + e).
template <int NIN, int CH>
__device__ __forceinline__ void mlp_layer0_chunk(const double* __restrict__ Wc, const double* __restrict__ Zs, int fr, int fq, double (&acc)[4][8][2]) {
#pragma unroll
    for (int kk = 0; kk < MLP_KC; kk++) {
        constexpr int K0 = CH * MLP_KC;
        const int k = K0 + kk;
        if (k < 3 * NIN) {
            const int kind = k / NIN, src = k - kind * NIN;
            double w[4];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) w[mb] = Wc[kk * 32 + mb * 8 + fr];
            const double z0 = Zs[k * 8 + 2 * fq], z1 = Zs[k * 8 + 2 * fq + 1];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) { acc[mb][0][0] = fma(w[mb], z0, acc[mb][0][0]); acc[mb][0][1] = fma(w[mb], z1, acc[mb][0][1]); }
            if (src < 7) {
                // d z_k / d q_src: 1 (x), cos q (sin row), -sin q (cos row)
                double t0 = 1.0, t1 = 1.0;
                if (kind == 1) { t0 = Zs[(2 * NIN + src) * 8 + 2 * fq]; t1 = Zs[(2 * NIN + src) * 8 + 2 * fq + 1]; }
                if (kind == 2) { t0 = -Zs[(NIN + src) * 8 + 2 * fq]; t1 = -Zs[(NIN + src) * 8 + 2 * fq + 1]; }
#pragma unroll
                for (int mb = 0; mb < 4; mb++) {
                    if (kind == 0) { acc[mb][1 + src][0] += w[mb]; acc[mb][1 + src][1] += w[mb]; }
                    else { acc[mb][1 + src][0] = fma(w[mb], t0, acc[mb][1 + src][0]); acc[mb][1 + src][1] = fma(w[mb], t1, acc[mb][1 + src][1]); }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) k_mlp(MlpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* Xs = reinterpret_cast<double2*>(smem_raw);                              // fragment-ordered activation tile
    double2* Wbuf = reinterpret_cast<double2*>(smem_raw + 256 * 64 * sizeof(double));  // 2 x [2048 double2]
    double* Xd = reinterpret_cast<double*>(Xs);
    // env output layer (9 x 256), resident, as A fragments: rows 0..7 [kb 64][lane 32] | row 8 [256]
    double* Wo0 = reinterpret_cast<double*>(smem_raw + (256 * 64 + 2 * MLP_CHUNK_D) * sizeof(double));
    double* Wo1 = Wo0 + 8 * 256;
    for (int i = threadIdx.x; i < 8 * 256; i += MLP_THREADS) Wo0[i] = a.w_out_env[((i & 31) >> 2) * 256 + (i >> 5) * 4 + (i & 3)];
    for (int i = threadIdx.x; i < 256; i += MLP_THREADS) Wo1[i] = a.w_out_env[8 * 256 + i];  // visible after the first __syncthreads() below

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fq = lane & 3;  // fragment row / quad index
    const int bslot = xslot(fq, fr);
    const int tx = tid & 7, ty = tid >> 3;    // input staging: sample, encoded row

    int p = 0;    // position in the chunk sequence
    int buf = 0;  // ring slot holding chunk p
    constexpr int SLICE2 = MLP_CHUNK_D / 2 / 8;  // double2 per warp and chunk (4 KB)
    double2* Wmine = Wbuf + warp * (2 * SLICE2);  // this warp's two ring slots
    auto prefetch = [&](int chunk, int slot) {
        const double2* src = reinterpret_cast<const double2*>(a.wpack) + ((size_t)chunk * 8 + warp) * SLICE2;
        double2* dst = Wmine + slot * SLICE2;
#pragma unroll
        for (int i = 0; i < SLICE2 / 32; i++) cp_async16(dst + lane + i * 32, src + lane + i * 32);
        cp_async_commit();
    };
    auto next_chunk = [&]() {  // the warp's slice of chunk p has landed; start fetching p + 1 into the slot the previous chunk used
        cp_async_wait_all();
        __syncwarp();
        prefetch((p + 1) % MLP_NCHUNK, buf ^ 1);
    };
    auto advance = [&]() { buf ^= 1; p = (p + 1) % MLP_NCHUNK; };
    if ((int)blockIdx.x < a.n_tiles) prefetch(0, 0);

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int s0 = tile * MLP_TILE_S;
        for (int net = 0; net < 2; net++) {  // 0: env, 1: self
            // ---- stage the encoded inputs z = [x, sin x, cos x] of the tile's 8 samples: Zs [k 32][sample 8] (in the idle X tile) ----
            __syncthreads();  // previous users of Xs are done
            double* Zs = Xd;
            {
                const int nin = (net == 0) ? 10 : 7;  // raw inputs; encoded rows = 3 * nin
                const int src = tid >> 3, sidx = tid & 7, n = s0 + sidx;
                if (src < nin) {
                    double xin = 0.0;
                    if (n < a.NS) xin = (src < 7) ? a.qs[(size_t)src * a.NS + n] : a.obs[(size_t)(n / a.S) * 4 + (src - 7)];
                    double sn, cs;
                    sincos(xin, &sn, &cs);
                    if (n >= a.NS) { xin = 0.0; sn = 0.0; cs = 0.0; }   // samples beyond the batch: zero inputs (results are not stored)
                    Zs[src * 8 + sidx] = xin; Zs[(nin + src) * 8 + sidx] = sn; Zs[(2 * nin + src) * 8 + sidx] = cs;
                }
            }
            __syncthreads();  // Z is in place
            const int n_hidden = (net == 0) ? 4 : 1;  // layers producing 256 neurons
            for (int layer = 0; layer < n_hidden; layer++) {
                double acc[4][8][2];
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[mb][c][0] = acc[mb][c][1] = 0.0;
                if (layer == 0) {
                    // encoded input -> 256 neurons on DFMA, exploiting the diagonal encoding Jacobian (mlp_layer0_chunk)
                    next_chunk();
                    if (net == 0) mlp_layer0_chunk<10, 0>(reinterpret_cast<const double*>(Wmine + buf * SLICE2), Zs, fr, fq, acc);
                    else mlp_layer0_chunk<7, 0>(reinterpret_cast<const double*>(Wmine + buf * SLICE2), Zs, fr, fq, acc);
                    advance();
                    next_chunk();
                    if (net == 0) mlp_layer0_chunk<10, 1>(reinterpret_cast<const double*>(Wmine + buf * SLICE2), Zs, fr, fq, acc);
                    else mlp_layer0_chunk<7, 1>(reinterpret_cast<const double*>(Wmine + buf * SLICE2), Zs, fr, fq, acc);
                    advance();
                } else {
                    for (int ch = 0; ch < 16; ch++) {
                        next_chunk();
                        mlp_chunk(Wmine + buf * SLICE2, Xs, ch * (MLP_KC / 4), lane, bslot, acc);
                        advance();
                    }
                }
                // bias + ReLU mask, then the tile becomes the next layer's input
                const double* bias = a.bias + ((net == 0) ? (MLP_BIAS_ENV + layer * 256) : MLP_BIAS_SELF0);
                double bv[4];
#pragma unroll
                for (int mb = 0; mb < 4; mb++) bv[mb] = bias[warp * 32 + mb * 8 + fr];
                __syncthreads();  // everyone finished reading Xs
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int row = warp * 32 + mb * 8 + fr, s = 2 * fq + e;
                        const double pre = acc[mb][0][e] + bv[mb];
                        const bool on = pre > 0.0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const double v0 = (j == 0) ? pre : acc[mb][2 * j][e], v1 = acc[mb][2 * j + 1][e];
                            Xs[xl2(row, j, s)] = on ? make_double2(v0, v1) : make_double2(0.0, 0.0);
                        }
                    }
                __syncthreads();  // the layer's output is in place
            }
            if (net == 0) {
                // ---- env output layer: 9 x 256 as two m-fragments (rows 0..7, row 8); warp = column kind ----
                double o[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};  // [k parity][m-fragment][sample]
                const double* xb = Xd + ((warp >> 1) * 32 + bslot) * 2 + (warp & 1);
#pragma unroll 4
                for (int kb = 0; kb < 64; kb += 2) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const double bx = xb[(kb + h) * 256];
                        const double a0 = Wo0[(kb + h) * 32 + lane], a1 = (lane < 4) ? Wo1[(kb + h) * 4 + lane] : 0.0;
                        dmma884(o[h][0][0], o[h][0][1], a0, bx);
                        dmma884(o[h][1][0], o[h][1][1], a1, bx);
                    }
                }
#pragma unroll
                for (int mb = 0; mb < 2; mb++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int l = mb * 8 + fr, ns = s0 + 2 * fq + e;
                        if (l < 9 && ns < a.NS) {
                            const double v = o[0][mb][e] + o[1][mb][e];
                            if (warp == 0) a.rb[(size_t)(RB_ENV + l) * a.NS + ns] = v + a.bias[MLP_BIAS_ENV_OUT + l];
                            else a.rb[(size_t)(RB_DENV + l * 7 + (warp - 1)) * a.NS + ns] = v;
                        }
                    }
                if (tid < 8 && (s0 + tid) < a.NS) {
                    const int n2 = s0 + tid;
                    a.rb[(size_t)RB_OBSR * a.NS + n2] = a.obs[(size_t)(n2 / a.S) * 4 + 3];
                }
            } else {
                // ---- self layer 1: 64 x 256; warp = m-fragment (8 neurons) x all 8 column kinds, full K ----
                double acc[8][2];
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c][0] = acc[c][1] = 0.0;
                for (int ch = 0; ch < 4; ch++) {
                    next_chunk();
                    const double2* Wc = Wmine + buf * SLICE2;  // [kb pair 8][lane 32] -> {kb even, kb odd}
#pragma unroll 2
                    for (int kp = 0; kp < 8; kp++) {
                        const double2 a2 = Wc[kp * 32 + lane];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int kb = ch * 16 + kp * 2 + h;
                            double2 b[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) b[j] = Xs[(kb * 4 + j) * 32 + bslot];
                            const double av = h ? a2.y : a2.x;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                dmma884(acc[2 * j][0], acc[2 * j][1], av, b[j].x);
                                dmma884(acc[2 * j + 1][0], acc[2 * j + 1][1], av, b[j].y);
                            }
                        }
                    }
                    advance();
                }
                const double bv = a.bias[MLP_BIAS_SELF1 + warp * 8 + fr];
                __syncthreads();  // done reading Xs: rows 0..63 become the layer's output
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int row = warp * 8 + fr, s = 2 * fq + e;
                    const double pre = acc[0][e] + bv;
                    const bool on = pre > 0.0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const double v0 = (j == 0) ? pre : acc[2 * j][e], v1 = acc[2 * j + 1][e];
                        Xs[xl2(row, j, s)] = on ? make_double2(v0, v1) : make_double2(0.0, 0.0);
                    }
                }
                __syncthreads();
                // ---- self output layer: 1 x 64, row 0 of one m-fragment; warp = column kind ----
                {
                    double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                    const double* xb = Xd + ((warp >> 1) * 32 + bslot) * 2 + (warp & 1);
#pragma unroll
                    for (int kb = 0; kb < 16; kb += 2)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const double av = (lane < 4) ? __ldg(a.w_out_self + (kb + h) * 4 + lane) : 0.0;
                            dmma884(o[h][0], o[h][1], av, xb[(kb + h) * 256]);
                        }
                    if (lane < 4) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int ns = s0 + 2 * lane + e;
                            if (ns < a.NS) {
                                const double v = o[0][e] + o[1][e];
                                if (warp == 0) a.rb[(size_t)RB_SEL * a.NS + ns] = v + a.bias[MLP_BIAS_SELF_OUT];
                                else a.rb[(size_t)(RB_DSEL + warp - 1) * a.NS + ns] = v;
                            }
                        }
                    }
                }
            }
        }
    }
    cp_async_wait_all();
}

#endif  // __CUDACC__

// Host-side packing of both networks' hidden-layer weights into the chunk stream k_mlp consumes (A fragments of
// mma.m8n8k4: lane l holds neuron l >> 2 of its m-fragment at k-step l & 3 of the 4-k block).
//   256-neuron layer, chunk = 16 k-steps:  [warp 8][kb 4][m-pair 2][lane 32][e 2]   = W[32 warp + (2 mp + e) 8 + (l >> 2)][k0 + 4 kb + (l & 3)]
//   self layer 1 (64 neurons), chunk = 64 k-steps: [warp 8][kb pair 8][lane 32][h 2] = W[8 warp + (l >> 2)][64 ch + 4 (2 kp + h) + (l & 3)]
// (a warp's slice of a chunk is contiguous: 512 doubles)
// Layer 0 of each net is zero-padded from 30 / 21 encoded inputs to K = 32.
inline void pack_mlp_weights(const double* const env_W[5], const double* const self_W[3], double* out) {
    auto pack256 = [&](const double* W, int in_dim, int k_pad, double*& o) {
        for (int k0 = 0; k0 < k_pad; k0 += MLP_KC)
            for (int warp = 0; warp < 8; warp++)
                for (int kb = 0; kb < MLP_KC / 4; kb++)
                    for (int mp = 0; mp < 2; mp++)
                        for (int l = 0; l < 32; l++)
                            for (int e = 0; e < 2; e++) {
                                const int row = 32 * warp + (2 * mp + e) * 8 + (l >> 2), k = k0 + 4 * kb + (l & 3);
                                *o++ = (k < in_dim) ? W[(size_t)row * in_dim + k] : 0.0;
                            }
    };
    // first layers (direct DFMA path, mlp_layer0_chunk): chunk = 16 encoded inputs: [warp 8][k 16][row 32] = W[32 warp + row][k0 + k], zero-padded to 32
    auto pack_l0 = [&](const double* W, int in_dim, double*& o) {
        for (int k0 = 0; k0 < 32; k0 += MLP_KC)
            for (int warp = 0; warp < 8; warp++)
                for (int kk = 0; kk < MLP_KC; kk++)
                    for (int row = 0; row < 32; row++) {
                        const int k = k0 + kk;
                        *o++ = (k < in_dim) ? W[(size_t)(32 * warp + row) * in_dim + k] : 0.0;
                    }
    };
    double* o = out;
    pack_l0(env_W[0], 30, o);
    pack256(env_W[1], 256, 256, o);
    pack256(env_W[2], 256, 256, o);
    pack256(env_W[3], 256, 256, o);
    pack_l0(self_W[0], 21, o);
    for (int ch = 0; ch < 4; ch++)
        for (int warp = 0; warp < 8; warp++)
            for (int kp = 0; kp < 8; kp++)
                for (int l = 0; l < 32; l++)
                    for (int h = 0; h < 2; h++) {
                        const int row = 8 * warp + (l >> 2), k = 64 * ch + 4 * (2 * kp + h) + (l & 3);
                        *o++ = self_W[1][(size_t)row * 256 + k];
                    }
}

}  // namespace mpcc
