// The collision networks with the three 256 x 256 layers of the env net (88 % of the MACs) on the 5th-generation tensor cores:
// tcgen05.mma kind::i8 (s8 x s8 -> s32 accumulators in TMEM) on an error-free integer split of the fp64 operands ("Ozaki scheme").
// Same interface, inputs and outputs as k_mlp (mlp_kernel.cuh); the other layers (first layers on DFMA with the diagonal encoding
// Jacobian, 64-neuron layer and output layers on DMMA) are k_mlp's.
//
// Arithmetic of one split layer  Y = W X  (W 256 x 256, X = 256 neurons x 64 columns = 8 samples x [value | 7 joint tangents]):
//   W[r,k] = 2^(E_r - 7S) qw,  qw = sum_i a_i[r,k] 128^(S-1-i),   X[k,c] ~ 2^(e_c - 7S) qx,  qx = rint(X 2^(7S - e_c)) = sum_j b_j[k,c] 128^(S-1-j)
//   with S signed 7-bit digits a_i, b_j in [-64, 64] (int8), row exponents E_r fixed at pack time, column exponents e_c = exponent of the
//   column's largest entry + 2, found by the producing layer's epilogue.  The digit products are exact in int32:
//       G_g = sum_{i+j=g} sum_k a_i[r,k] b_j[k,c]      (|G_g| <= (g+1) 2^20),      g = 0 .. S-1   (products with i + j >= S are dropped)
//   and  Y[r,c] = 2^(E_r + e_c - 7S - 7) (((G_0 128 + G_1) 128 + ...) + G_(S-1))   with the bracket exact in int64 and ONE rounding to fp64.
//   Dropped terms: <= (S+1) 2^(-7S) of |W|_max |X|_max K, i.e. 7e-15 relative to full scale at S = 7 (an fp64 dot product of length 256
//   carries ~3e-14 in the same norm); what is not kept is the low bits of entries far below their row / column maximum.
//
// Mapping.  One persistent CTA per SM, tiles of 8 samples as in k_mlp.  Per layer:
//   split     thread = neuron k: its 64 fp64 activations -> S digit planes of the B operand, [n 64][k 256] int8 MN-major core matrices
//             (16 consecutive columns of one neuron are one 16-byte store); the fp64 tile rows 128..255 are overwritten by the planes;
//   2 passes  (neurons 0..127, 128..255 = M of the MMA): one elected thread issues S (S+1) / 2 x 8 MMAs of 128 x 64 x 32; the digit planes of
//             W stream from L2 in 8 KB chunks (128 rows x 64 k, canonical K-major core matrices, packed on the host in consumption order)
//             through a ring filled by cp.async.bulk and released by tcgen05.commit; S accumulators of 64 TMEM columns;
//   epilogue  8 warps: tcgen05.ld of the S accumulators, int64 Horner, scale, bias, ReLU mask, fp64 tile write-back, column maxima.
// Measured (tools/probes/oz_umma_probe.cu): 53 cycles per 128 x 64 x 32 MMA with both operands in shared memory (the 4 KB A read per MMA is
// the bound: 116 B/cycle; the tensor floor is 32), TMEM read-back 200 B/cycle with 4 warps.
#pragma once
#include "mlp_kernel.cuh"
#include <cstdint>
#include <cmath>
#include <cstring>

namespace mpcc {

#ifndef OZ_DIGITS
#define OZ_DIGITS 7
#endif
constexpr int OZ_S = OZ_DIGITS;                        // int8 digits per operand (7: 49-bit operands, fp64-equivalent; -DOZ_DIGITS=6: 42 bits)
#ifndef OZ_CHUNK_BYTES
#define OZ_CHUNK_BYTES 8192
#endif
constexpr int OZ_CHUNK = OZ_CHUNK_BYTES;               // bytes of a chunk: 128 rows x OZ_KCH k
constexpr int OZ_NSLOT = 49152 / OZ_CHUNK;             // ring slots for the weight-digit chunks
constexpr int OZ_KCH = OZ_CHUNK / 128;                 // k per chunk (64: two MMA k-steps)
constexpr int OZ_CPP = 256 / OZ_KCH;                   // chunks per plane
constexpr int OZ_CHUNKS_PER_PASS = OZ_S * OZ_CPP;
constexpr int OZ_CHUNKS_PER_LAYER = 2 * OZ_CHUNKS_PER_PASS;
constexpr int OZ_CHUNKS_PER_TILE = 3 * OZ_CHUNKS_PER_LAYER;
constexpr int OZ_PLANE = 64 * 256;                     // bytes of one digit plane of the activation tile
constexpr int OZ_OFF_PLANES = 65536;                   // the planes start on the fp64 tile's rows 128..255
constexpr int OZ_OFF_RING = (OZ_OFF_PLANES + OZ_S * OZ_PLANE > 131072) ? OZ_OFF_PLANES + OZ_S * OZ_PLANE : 131072;
constexpr int OZ_RING_BYTES = OZ_NSLOT * OZ_CHUNK;     // also holds the per-warp rings of the DMMA layers (8 x 2 x 2 KB)
constexpr int OZ_OFF_MISC = OZ_OFF_RING + OZ_RING_BYTES;
constexpr int OZ_NBARS = 2 * OZ_NSLOT + OZ_S + 1;     // full[NSLOT] | empty[NSLOT] | accumulator g complete [S] | accumulators read
constexpr int OZ_MISC_COLMAX = 0, OZ_MISC_SC = 256, OZ_MISC_COLSCALE = 768, OZ_MISC_BARS = 1280, OZ_MISC_TMEM = 1280 + 8 * OZ_NBARS;
constexpr size_t OZ_SMEM_BYTES = OZ_OFF_MISC + 1536;
static_assert(OZ_NSLOT == 6 && OZ_CHUNK == 8192, "three producer warps with two ring slots each");
static_assert(OZ_S >= 4 && OZ_S <= 7, "int64 Horner of the accumulators holds up to 7 digits");
static_assert(OZ_S * 64 + 4 * (OZ_CHUNK_BYTES / 512) <= 512, "accumulators and the four A buffers must fit TMEM");
static_assert(OZ_RING_BYTES >= 32768, "the DMMA layers' per-warp rings live in the ring region");
static_assert(OZ_SMEM_BYTES <= 232448, "shared memory");
// fp64 chunk stream of one tile: env L0 (4 chunks of 8 encoded inputs) | self L0 (4); 16 KB each
constexpr int OZ_DCHUNK_D = 2048;
constexpr int OZ_NDCHUNK = 8;
constexpr int OZ_WOUT_D = 9 * 256;      // env output layer, appended to the chunk stream: rows 0..7 in A-fragment order | row 8
// The self net runs in REVERSE mode (it has one output: one adjoint sweep instead of seven tangent columns, 44 k instead of 142 k MAC per sample).
// Its 64 x 256 layer goes to shared memory once per tile, row-major with a padded stride, and is read from there as the A operand of both the
// forward product (A[neuron][k]) and the adjoint product (A[k][neuron], the transpose): both reads are bank-conflict free at stride 260.
constexpr int OZ_W1_D = 64 * 256;       // self layer 1, raw row-major, appended after the output layer
constexpr int OZ_W0T_D = 8 * 3 * 8 * 32; // self layer 0 transposed, as A fragments [warp 8][m-fragment 3 (21 encoded inputs, padded to 24)][k-step 8][lane 32]
constexpr int OZ_DOFF_WOUT = OZ_NDCHUNK * OZ_DCHUNK_D, OZ_DOFF_W1 = OZ_DOFF_WOUT + OZ_WOUT_D, OZ_DOFF_W0T = OZ_DOFF_W1 + OZ_W1_D;
constexpr size_t OZ_DPACK_D = (size_t)OZ_DOFF_W0T + OZ_W0T_D;   // doubles of the packed fp64 weights (MlpOzArgs::m.wpack)
// the self net's shared memory, in doubles from the start of the (then idle) activation tile: encoded inputs | layer-1 activations, later their adjoints,
// [k-step 64][sample 8][k 4] (a B fragment is 32 consecutive doubles) | layer-2 adjoints [16][8][4] | per-warp partial outputs | per-warp partial input adjoints | W1
constexpr int OZ_SF_H1 = 256, OZ_SF_A2 = OZ_SF_H1 + 2048, OZ_SF_SELP = OZ_SF_A2 + 512, OZ_SF_GZP = OZ_SF_SELP + 64, OZ_SF_W1 = 4608, OZ_SF_W1_LD = 260;
static_assert(OZ_SF_GZP + 8 * 24 * 8 <= OZ_SF_W1 && (OZ_SF_W1 + 64 * OZ_SF_W1_LD) * 8 <= OZ_OFF_RING, "self-net scratch must fit below the ring");
constexpr int OZ_OFF_WOUT = 131072;     // its place in shared memory: above the fp64 tile, in the (by then dead) digit planes
static_assert(OZ_OFF_WOUT + OZ_WOUT_D * 8 <= OZ_OFF_RING, "output-layer weights must fit between the tile and the ring");
// weight plane consumed at position pi of a pass: heavy (low i: S - i products) and light planes alternate, so that the weight stream's demand
// per unit of tensor time is even (ascending order leaves the light planes at the end of every pass, where the stream cannot keep up)
constexpr int oz_order(int pi) { return (pi & 1) ? OZ_S - 1 - pi / 2 : pi / 2; }
constexpr int oz_pos_of_plane(int i) { return (2 * i < OZ_S) ? 2 * i : 2 * (OZ_S - 1 - i) + 1; }
constexpr long long oz_bias_const() {  // 64 on each of the lower S - 1 digits: makes them unsigned fields (the top digit stays signed)
    long long c = 0;
    for (int t = 0; t < OZ_S - 1; t++) c += 64LL << (7 * t);
    return c;
}


// ---- scalar arithmetic of the split, shared by the kernel and the host emulation of the CPU test tier (tests/emul) ----
#if defined(__CUDACC__)
#define OZ_HD __host__ __device__ __forceinline__
#else
#define OZ_HD inline
#endif
OZ_HD double oz_bits_to_double(unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d; std::memcpy(&d, &b, 8); return d;
#endif
}
OZ_HD long long oz_double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    long long b; std::memcpy(&b, &d, 8); return b;
#endif
}
// element (k, sample s) of a [k][8 samples] operand stored so that the B fragment of k-step ks (lane l holds k = 4 ks + (l & 3), sample l >> 2) is 32 consecutive doubles
OZ_HD int oz_bfrag(int k, int s) { return (k >> 2) * 32 + s * 4 + (k & 3); }
// column scales from the high word of the column's largest |entry|: sc = 2^(7 S - e_c), cs = 2^(e_c), e_c = exponent + 2 (so |x| 2^(-e_c) < 1/2);
// an all-zero (or denormal-range) column gets sc = 0: every digit 0
OZ_HD void oz_col_scales(uint32_t hi, double& sc, double& cs) {
    const int be = (int)(hi >> 20);  // biased exponent
    sc = 0.0; cs = 1.0;
    if (be >= 7 * OZ_S && be <= 2040) {
        sc = oz_bits_to_double((unsigned long long)(2044 + 7 * OZ_S - be) << 52);
        cs = oz_bits_to_double((unsigned long long)(be + 2) << 52);
    }
}
// q' = rint(x sc) + bias constant: rint by the magic-number addition (|x sc| < 2^48), so the lower S - 1 digits are the unsigned 7-bit fields of
// q' minus 64 and the top digit is its arithmetic shift
OZ_HD long long oz_quantize(double x, double sc) {
#if defined(__CUDA_ARCH__)
    const double t = fma(x, sc, 6755399441055744.0);
#else
    const double t = std::fma(x, sc, 6755399441055744.0);
#endif
    return oz_double_to_bits(t) - 0x4338000000000000LL + oz_bias_const();
}
OZ_HD int oz_digit(long long q, int t) { return (t < OZ_S - 1) ? (int)((q >> (7 * t)) & 127) - 64 : (int)(q >> (7 * (OZ_S - 1))); }  // digit of weight 128^t
// where digit plane i of W[r][k] sits in the packed stream (byte index)
inline size_t oz_wq_index(int layer, int r, int k, int i) {
    const int mb = r / 128, rr = r % 128, kk = k % OZ_KCH;
    const size_t chunk = (size_t)layer * OZ_CHUNKS_PER_LAYER + (size_t)(mb * OZ_S + oz_pos_of_plane(i)) * OZ_CPP + k / OZ_KCH;
    return chunk * OZ_CHUNK + (rr % 8) * 16 + (rr / 8) * 128 + (kk % 16) + (kk / 16) * 2048;
}

struct MlpOzArgs {
    MlpArgs m;               // m.wpack: the fp64 weights packed by pack_mlp_oz_weights (OZ_DPACK_D doubles: chunk stream | env output layer | self W1 | self W0 transposed)
    const uint8_t* wq;       // OZ_CHUNKS_PER_TILE x OZ_CHUNK bytes: digit planes of env layers 1..3 in consumption order
    const double* rowscale;  // [3][256]: 2^(E_r - 7 S - 7)
    int dbg_flags;           // experiments (timing only, results wrong): bit 0 = no weight stream (MMAs on whatever the ring holds), bit 1 = no proxy fence per chunk, bit 2 = every pass issued twice (cold / warm instruction fetch)
    long long* dbg;          // optional: cycles of CTA 0 per phase [first layers + staging | split | MMA pass (issue .. accumulators ready) | epilogue | env output | self net | tiles]
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t oz_desc(uint32_t addr, int lbo, int sbo) {  // shared-memory matrix descriptor, no swizzle, version 1
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void oz_mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {  // A operand in TMEM
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void oz_utccp(uint32_t tmem_dst, uint64_t sdesc) {  // 128 rows x 32 bytes, shared memory (canonical K-major core matrices) -> 8 TMEM columns
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void oz_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool oz_elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}\n" : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ bool oz_elect_mask(uint32_t mask) {
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}\n" : "+r"(laneid), "+r"(pred) : "r"(mask));
    return pred != 0;
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);  // waiting warps must not flood the shared-memory pipe the tensor core reads its operands through
        if (clock64() - t0 > 4000000000LL) __trap();  // two seconds: a protocol error must end the kernel, not hang the GPU
    }
}
// the same wait as one PTX block (no C++ control flow): inside the MMA issuer's elected region a C++ wait loop makes the compiler wrap every
// tcgen05.mma in an election loop (+40 cycles per MMA)
__device__ __forceinline__ void oz_mbar_wait_asm(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tOZ_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra OZ_DONE;\n\tbra OZ_WAIT;\n\tOZ_DONE:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// the same load without the wait (several loads in flight; oz_tmem_wait_ld() before the first use of any of them)
__device__ __forceinline__ void oz_tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void oz_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// first layer, one chunk of 8 encoded inputs (K0 .. K0 + 7) on DFMA: mlp_layer0_chunk of mlp_kernel.cuh with an 8-row chunk
//   Wc: the warp's slice [k 8][row 32]    Zs: [k 32][sample 8]
template <int NIN, int K0>
__device__ __forceinline__ void oz_layer0_chunk(const double* __restrict__ Wc, const double* __restrict__ Zs, int fr, int fq, double (&acc)[4][8][2]) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
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


// the same chunk, value column only (self net: its derivatives come from the adjoint sweep)
template <int NIN, int K0>
__device__ __forceinline__ void oz_layer0_chunk_value(const double* __restrict__ Wc, const double* __restrict__ Zs, int fr, int fq, double (&acc)[4][2]) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        const int k = K0 + kk;
        if (k < 3 * NIN) {
            const double z0 = Zs[k * 8 + 2 * fq], z1 = Zs[k * 8 + 2 * fq + 1];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                const double w = Wc[kk * 32 + mb * 8 + fr];
                acc[mb][0] = fma(w, z0, acc[mb][0]); acc[mb][1] = fma(w, z1, acc[mb][1]);
            }
        }
    }
}


// One MMA pass, issued by one thread: the S (S + 1) / 2 digit products of 128 neurons x 64 columns x 256 k, chunk by chunk as the weight digits
// arrive, planes in ascending order; after the last chunk of plane i a commit on `group` barrier i tells the epilogue warps that accumulator i is
// complete.  Fully unrolled: every descriptor offset, accumulator address and accumulate flag is an immediate (rolled: +20 %).  Called by the whole of warp 0,
// converged; the elections inside tell the compiler that exactly one thread issues (otherwise it wraps every MMA in an election loop, +40 cycles).
__device__ __noinline__ long long oz_issue_pass(uint32_t tmem, uint32_t ring_addr, uint32_t planes_addr, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_group, uint32_t bar_tfree,
                                          uint32_t n, uint32_t tfree_wait, bool no_stream, bool no_fence) {
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // s32 += s8 (K-major) x s8 (MN-major), M 128, N 64
    // warp-uniform copies of the arguments (tcgen05 operands live in uniform registers)
    tmem = __shfl_sync(0xffffffffu, tmem, 0); ring_addr = __shfl_sync(0xffffffffu, ring_addr, 0); planes_addr = __shfl_sync(0xffffffffu, planes_addr, 0);
    bar_full = __shfl_sync(0xffffffffu, bar_full, 0); bar_empty = __shfl_sync(0xffffffffu, bar_empty, 0); bar_group = __shfl_sync(0xffffffffu, bar_group, 0);
    bar_tfree = __shfl_sync(0xffffffffu, bar_tfree, 0); n = __shfl_sync(0xffffffffu, n, 0); tfree_wait = __shfl_sync(0xffffffffu, tfree_wait, 0);
    // One elected lane walks the chunks, waits and issues; the other lanes park at the final __syncwarp.  After every wait the lane re-elects
    // itself with a one-lane mask: a wait loop inside a single-thread region otherwise makes the compiler wrap every tcgen05.mma in an election
    // loop (+40 cycles per MMA).  (Measured: the whole warp waiting converged and electing per chunk, the canonical shape, is 10 % slower here.)
    long long waited = 0;
    if (oz_elect_one()) {
        const uint32_t self_mask = 1u << (threadIdx.x & 31);
        if (tfree_wait) oz_mbar_wait_asm(bar_tfree, tfree_wait - 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        uint32_t slot = n % OZ_NSLOT, ph = (n / OZ_NSLOT) & 1;
        const uint64_t bd0 = oz_desc(planes_addr, 128, 4096);
#pragma unroll
        for (int pi = 0; pi < OZ_S; pi++) {
            const int i = oz_order(pi);
#pragma unroll
            for (int kc = 0; kc < OZ_CPP; kc++) {  // (rolling this loop -- a quarter of the code -- is 10 % slower, rolling all of them 20 %)
                if (!no_stream) oz_mbar_wait_asm(bar_full + 8 * slot, ph);
                if (!no_fence) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the chunk was written through the generic proxy (cp.async)
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint64_t ad = oz_desc(ring_addr + slot * OZ_CHUNK, 2048, 128);
                if (oz_elect_mask(self_mask)) {
                    // the chunk goes to TMEM once (tcgen05.cp, 8 columns per 32-k step; 4 rotating buffers behind the accumulators) and the MMAs take
                    // A from there: shared memory is read once per chunk instead of once per MMA, and the ring slot is free as soon as the copy is
                    // done.  Copies and MMAs execute in issue order, so a buffer is not overwritten before the MMAs issued earlier have read it.
                    const uint32_t ta = tmem + OZ_S * 64 + ((pi * OZ_CPP + kc) & 3) * (OZ_KCH / 4);
#pragma unroll
                    for (int ks = 0; ks < OZ_KCH / 32; ks++) oz_utccp(ta + ks * 8, ad + (uint64_t)((ks * 4096) >> 4));
                    oz_commit(bar_empty + 8 * slot);
#pragma unroll
                    for (int j = 0; j + i < OZ_S; j++) {  // accumulator g = i + j
                        const uint64_t bd = bd0 + (uint64_t)((j * OZ_PLANE + kc * (OZ_KCH * 16)) >> 4);
#pragma unroll
                        for (int ks = 0; ks < OZ_KCH / 32; ks++)
                            oz_mma_i8_ts(tmem + (i + j) * 64, ta + ks * 8, bd + (uint64_t)((ks * 512) >> 4), IDESC, (pi > 0 || kc > 0 || ks > 0) ? 1u : 0u);  // plane 0 comes first and starts every accumulator
                    }
                    if (kc == OZ_CPP - 1) oz_commit(bar_group + 8 * pi);  // the last position's commit says: pass complete
                }
                slot = (slot + 1 == (uint32_t)OZ_NSLOT) ? 0u : slot + 1;
                ph ^= (slot == 0) ? 1u : 0u;
            }
        }
    }
    __syncwarp();
    for (int o = 16; o; o >>= 1) { const long long w2 = __shfl_xor_sync(0xffffffffu, waited, o); waited = w2 > waited ? w2 : waited; }
    return waited;
}


// Weight-digit stream of one producer warp (pw = 0..2): its chunks m = pw (mod 3) in [begin, end) of the CTA's stream, global -> ring slot
// m % NSLOT by 16-byte cp.async (LDGSTS), fully asynchronous: after the copies of a chunk every lane issues cp.async.mbarrier.arrive.noinc on
// the slot's `full` barrier, which arrives when that lane's copies have landed -- no wait and no registers on the producer side, both of the
// warp's slots in flight.  (The pattern of CUTLASS' sm100 cp.async + UMMA mainloop; the MMA issuer adds a proxy fence after its wait.)
// Measured alternatives: cp.async.bulk (one 8 / 16 KB bulk copy per chunk completes every ~1300 / ~1700 cycles however many are outstanding:
// 6 .. 10 B/cycle, a third of what the MMAs consume); wait_group + fence + arrive by the copying threads (one chunk in flight per warp, six
// producer warps); copies through registers in an out-of-line function (the ABI leaves it half the register file: the loads were spilled at once).
__device__ __forceinline__ void oz_produce_range(const uint8_t* __restrict__ wq, unsigned char* ring, uint32_t bar_full, uint32_t bar_empty, uint32_t begin, uint32_t end,
                                                uint32_t tile_first, uint32_t pw, uint32_t lane) {
    for (uint32_t m = begin + (pw + 3u - (begin - tile_first) % 3u) % 3u; m < end; m += 3) {
        const uint32_t slot = m % OZ_NSLOT, use = m / OZ_NSLOT;
        if (use > 0) oz_mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);  // the slot's previous tenant has been copied to TMEM
        const unsigned char* src = wq + (size_t)(m % OZ_CHUNKS_PER_TILE) * OZ_CHUNK + lane * 16;
        unsigned char* dst = ring + slot * OZ_CHUNK + lane * 16;
#pragma unroll
        for (int i = 0; i < OZ_CHUNK / 512; i++) cp_async16(dst + i * 512, src + i * 512);
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar_full + 8 * slot) : "memory");
    }
}

__device__ __forceinline__ uint32_t oz_abs_hi(double v) { return (uint32_t)__double2hiint(v) & 0x7FFFFFFFu; }

// Split of one layer's input: thread = neuron k; its 64 fp64 activations -> S digit planes of the B operand ([n 64][k 256] int8, MN-major core
// matrices: 16 consecutive columns of a neuron are one 16-byte store).  Column exponents come from the maxima the producing layer left in colmax.
// All 256 threads call it (it synchronises the CTA).
__device__ __forceinline__ void oz_split(const double2* Xs, unsigned char* planes, uint32_t* colmax, double* s_sc, double* s_colscale, int tid) {
    if (tid < 64) {
        double sc, cs;
        oz_col_scales(colmax[tid], sc, cs);
        s_sc[tid] = sc;
        s_colscale[tid] = cs;
        colmax[tid] = 0;
    }
    __syncthreads();
    uint32_t w[OZ_S][16];
#pragma unroll
    for (int i = 0; i < OZ_S; i++)
#pragma unroll
        for (int q = 0; q < 16; q++) w[i][q] = 0;
    const int k = tid;
#pragma unroll
    for (int s = 0; s < 8; s++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double2 x2 = Xs[xl2(k, j, s)];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int c = 8 * s + 2 * j + h;             // column; 16-column group c >> 4, byte c & 15
                const int word = (c >> 4) * 4 + ((c & 15) >> 2), sh = 8 * (c & 3);
                const long long q = oz_quantize(h ? x2.y : x2.x, s_sc[c]);
#pragma unroll
                for (int t = 0; t < OZ_S - 1; t++) w[OZ_S - 1 - t][word] |= ((uint32_t)(q >> (7 * t)) & 127u) << sh;
                w[0][word] |= ((uint32_t)(q >> (7 * (OZ_S - 1))) & 255u) << sh;   // top digit: signed
            }
        }
    __syncthreads();  // every thread holds its row: the planes may overwrite rows 128..255 of the tile
#pragma unroll
    for (int i = 0; i < OZ_S; i++)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint4 v = make_uint4(w[i][4 * g], w[i][4 * g + 1], w[i][4 * g + 2], w[i][4 * g + 3]);
            if (i > 0) {  // unsigned field u -> signed digit u - 64, per byte
                v.x = ((v.x | 0x80808080u) - 0x40404040u) ^ 0x80808080u; v.y = ((v.y | 0x80808080u) - 0x40404040u) ^ 0x80808080u;
                v.z = ((v.z | 0x80808080u) - 0x40404040u) ^ 0x80808080u; v.w = ((v.w | 0x80808080u) - 0x40404040u) ^ 0x80808080u;
            }
            *reinterpret_cast<uint4*>(planes + i * OZ_PLANE + g * 4096 + (k >> 3) * 128 + (k & 7) * 16) = v;
        }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> tensor-core (async-proxy) reads
    __syncthreads();
}

// Epilogue of one pass, all eight warps: warp w reads TMEM lane quadrant w & 3 (neuron = lane) and columns 32 (w >> 2) .. + 31 of the S
// accumulators, folds them into int64 Horner sums, scales, adds the bias, applies the ReLU mask of the sample's value column to its 8 columns
// and writes the fp64 tile; the column maxima for the next split are reduced over the warp and kept by the lane that owns the column, one
// shared-memory atomic per lane at the end.
__device__ __forceinline__ void oz_epilogue_pass(double2* Xs, uint32_t* colmax, const double* s_colscale, const double* __restrict__ rowscale, const double* __restrict__ bias,
                                                uint32_t tmem, int mb, int warp, int lane, bool want_colmax) {
    const int qd = warp & 3, hh = warp >> 2;
    const int row = mb * 128 + qd * 32 + lane;
    const double rs = __ldg(rowscale + row), bv = __ldg(bias + row);
    // Horner with the accumulators taken in pairs: G_g 128 + G_(g+1) still fits int32 (|G_g| <= (g + 1) 2^20), so the 64-bit steps are halved;
    // the two TMEM loads of a pair are in flight together (one wait per pair instead of one per accumulator)
    long long acc[32];
#pragma unroll
    for (int g = 0; g < OZ_S; g += 2) {
        uint32_t va[32], vb[32];
        const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + g * 64 + hh * 32;
        oz_tmem_ld32_nowait(ta, va);
        if (g + 1 < OZ_S) oz_tmem_ld32_nowait(ta + 64, vb);
        oz_tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; c++) {
            if (g + 1 < OZ_S) {
                const int p = (int)va[c] * 128 + (int)vb[c];
                acc[c] = (g == 0) ? (long long)p : acc[c] * 16384 + (long long)p;
            } else acc[c] = (g == 0) ? (long long)(int)va[c] : acc[c] * 128 + (long long)(int)va[c];
        }
    }
    uint32_t mymax = 0;  // lane c: largest |entry| (high word) of column 32 hh + c over this warp's 32 neurons
#pragma unroll
    for (int sl = 0; sl < 4; sl++) {
        const int s = 4 * hh + sl;
        double y[8];
#pragma unroll
        for (int c = 0; c < 8; c++) y[c] = (__ll2double_rn(acc[8 * sl + c]) * rs) * s_colscale[8 * s + c];
        y[0] += bv;
        const bool on = y[0] > 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) Xs[xl2(row, j, s)] = on ? make_double2(y[2 * j], y[2 * j + 1]) : make_double2(0.0, 0.0);
        if (want_colmax) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint32_t m = __reduce_max_sync(0xffffffffu, on ? oz_abs_hi(y[c]) : 0u);
                if (lane == 8 * sl + c) mymax = m;
            }
        }
    }
    if (want_colmax) atomicMax(&colmax[32 * hh + lane], mymax);
}

__global__ void __launch_bounds__(MLP_THREADS, 1) k_mlp_oz(MlpOzArgs oa) {
    const MlpArgs& a = oa.m;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2* Xs = reinterpret_cast<double2*>(smem_raw);      // fragment-ordered fp64 activation tile (k_mlp's layout)
    double* Xd = reinterpret_cast<double*>(Xs);
    unsigned char* planes = smem_raw + OZ_OFF_PLANES;
    unsigned char* ring = smem_raw + OZ_OFF_RING;
    unsigned char* misc = smem_raw + OZ_OFF_MISC;
    uint32_t* colmax = reinterpret_cast<uint32_t*>(misc + OZ_MISC_COLMAX);   // [64] high word of the largest |entry| of each column of the next split
    double* s_sc = reinterpret_cast<double*>(misc + OZ_MISC_SC);             // [64] 2^(7 S - e_c)
    double* s_colscale = reinterpret_cast<double*>(misc + OZ_MISC_COLSCALE); // [64] 2^(e_c)
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc + OZ_MISC_BARS);       // full[NSLOT] | empty[NSLOT] | accumulator complete [S] | accumulators read
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + OZ_MISC_TMEM);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // the shuffle makes the warp index provably warp-uniform for the compiler
    const int fr = lane >> 2, fq = lane & 3;
    const int bslot = xslot(fq, fr);
    const uint32_t bar_full = oz_smem_u32(bars), bar_empty = oz_smem_u32(bars + OZ_NSLOT), bar_group = oz_smem_u32(bars + 2 * OZ_NSLOT),
                   bar_tfree = oz_smem_u32(bars + 2 * OZ_NSLOT + OZ_S);

    if (tid == 0) {
        for (int i = 0; i < OZ_NBARS; i++)  // arrivals: full = the 32 lanes of the copying warp, empty / group = one tcgen05.commit, read = the 128 epilogue threads
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_full + 8 * i), "r"(i < OZ_NSLOT ? 32 : (i == OZ_NBARS - 1 ? 128 : 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 64) colmax[tid] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(oz_smem_u32(tmem_ptr_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_ptr_s;

    // ---- per-warp rings of the DMMA / DFMA layers (k_mlp's scheme with three slots of 2 KB per warp, inside the ring region): while chunk p is
    //      being used, p + 1 and p + 2 are in flight.  The stream pauses after the env net's first layer (chunk 3 of a tile's 16): the ring region
    //      then belongs to the weight-digit stream until the split layers are done.
    constexpr int SLICE2 = OZ_DCHUNK_D / 2 / 8;  // double2 per warp and chunk
    constexpr uint32_t DSTG = 3;
    static_assert(8 * DSTG * SLICE2 * 16 <= OZ_RING_BYTES, "per-warp rings must fit the ring region");
    double2* Wmine = reinterpret_cast<double2*>(ring) + warp * (DSTG * SLICE2);
    uint32_t p_abs = 0, pi_abs = 0;  // chunks consumed / requested so far (chunk id = counter % OZ_NDCHUNK, slot = counter % DSTG)
    auto d_topup = [&]() {
        const uint32_t r = p_abs % OZ_NDCHUNK, base = p_abs - r;
        const uint32_t lim = (r < 4) ? base + 4 : base + OZ_NDCHUNK + 4;  // never across the pause point
        while (pi_abs < lim && pi_abs < p_abs + DSTG) {
            const double2* src = reinterpret_cast<const double2*>(a.wpack) + ((size_t)(pi_abs % OZ_NDCHUNK) * 8 + warp) * SLICE2;
            double2* dst = Wmine + (pi_abs % DSTG) * SLICE2;
#pragma unroll
            for (int i = 0; i < SLICE2 / 32; i++) cp_async16(dst + lane + i * 32, src + lane + i * 32);
            cp_async_commit();
            pi_abs++;
        }
    };
    auto d_wait = [&]() -> const double2* {  // chunk p_abs has landed
        const uint32_t younger = pi_abs - p_abs - 1;
        if (younger == 0) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        else if (younger == 1) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        else asm volatile("cp.async.wait_group 2;\n" ::: "memory");
        __syncwarp();
        return Wmine + (p_abs % DSTG) * SLICE2;
    };
    auto d_done = [&](bool pause_after) {  // the warp has read chunk p_abs: its slot may be refilled
        __syncwarp();
        p_abs++;
        if (!pause_after) d_topup();
    };
    if ((int)blockIdx.x < a.n_tiles) d_topup();

    // ---- weight-digit stream (used by the elected lane of warp 0 only) ----
    uint32_t tile_iter = 0;   // tiles this CTA has started: chunk numbers of the stream follow from it (no per-thread state: any lane may be elected)
    uint32_t n_pass = 0;      // passes done so far (parity of the accumulator barriers; every thread counts)
    uint32_t n_layer = 0;     // split layers done so far (parity of the `accumulators read` barrier)
    const uint32_t ring_addr = oz_smem_u32(ring), planes_addr = oz_smem_u32(planes);
    const bool no_stream = (oa.dbg_flags & 1) != 0;
    // Roles while the split layers run: warp 0 issues the MMAs, warps 1..3 stream the weight digits, warps 4..7 (one TMEM lane quadrant each) are
    // the epilogue.  Producer warp pw copies the chunks m = pw (mod 3) of the stream -- ring slots pw and pw + 3 -- through registers: 16 x 16
    // bytes per lane and chunk, two chunks in flight per warp (loads are issued two chunks ahead of their stores).  A chunk is handed to the
    // tensor core by its 32 lanes: stores done, proxy fence (generic-proxy writes -> async-proxy reads), arrive on the slot's `full` barrier.
    // Measured alternatives: cp.async.bulk (one 8 / 16 KB bulk copy per chunk completes every ~1300 / ~1700 cycles however many are outstanding:
    // 6 .. 10 B/cycle, a third of what the MMAs consume); cp.async with several groups in flight per thread (the proxy fence waits for all of them).
    const bool producer = warp >= 1 && warp <= 3;
    long long dbg_t[7] = {0, 0, 0, 0, 0, 0, 0}, dbg_c = 0, dbg_wait = 0;
    const bool dbg_on = oa.dbg != nullptr && blockIdx.x == 0 && tid == 0;
#define OZ_MARK(w, e) if (oa.dbg != nullptr && blockIdx.x == 0 && tile_iter == 1 && lane == 0) oa.dbg[8 + (w) * 16 + (e)] = clock64();
#define OZ_DBG(i) if (dbg_on) { const long long t_ = clock64(); dbg_t[i] += t_ - dbg_c; dbg_c = t_; }
    if (dbg_on) { dbg_c = clock64(); oa.dbg[40] = 0; oa.dbg[41] = 0; }
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, tile_iter++) {
        const int s0 = tile * MLP_TILE_S;
        dbg_t[6]++;
        const uint32_t tile_first = tile_iter * OZ_CHUNKS_PER_TILE, tile_end = tile_first + OZ_CHUNKS_PER_TILE;
        for (int net = 0; net < 2; net++) {  // 0: env, 1: self
            // ---- encoded inputs z = [x, sin x, cos x] of the 8 samples: Zs [k 32][sample 8] (in the idle X tile) ----
            if (warp == 0) { OZ_MARK(1, 5 * net + 0) }
            __syncthreads();
            double* Zs = Xd;
            if (net == 1) {
                // self layer 1 (64 x 256) -> shared memory, rows padded to OZ_SF_W1_LD.  The copies are NOT committed here: they join the group of the
                // next ring chunk this thread requests (the last first-layer chunk), so the first-layer chunks before it do not wait for them.
                // (Measured alternative: 64 bulk copies of one row each from one thread onto an mbarrier -- 1.2 k cycles per tile slower, the issue is serial.)
                const double2* src = reinterpret_cast<const double2*>(a.wpack + OZ_DOFF_W1);
                double2* dst = reinterpret_cast<double2*>(Xd + OZ_SF_W1);
#pragma unroll 8
                for (int i = tid; i < 64 * 128; i += MLP_THREADS) cp_async16(dst + (i >> 7) * (OZ_SF_W1_LD / 2) + (i & 127), src + i);
            }
            {
                const int nin = (net == 0) ? 10 : 7;
                const int src = tid >> 3, sidx = tid & 7, n = s0 + sidx;
                if (src < nin) {
                    double xin = 0.0;
                    if (n < a.NS) xin = (src < 7) ? a.qs[(size_t)src * a.NS + n] : a.obs[(size_t)(n / a.S) * 4 + (src - 7)];
                    double sn, cs;
                    sincos(xin, &sn, &cs);
                    if (n >= a.NS) { xin = 0.0; sn = 0.0; cs = 0.0; }
                    Zs[src * 8 + sidx] = xin; Zs[(nin + src) * 8 + sidx] = sn; Zs[(2 * nin + src) * 8 + sidx] = cs;
                }
            }
            __syncthreads();
            if (warp == 0) { OZ_MARK(1, 5 * net + 1) }
            // ---- first layer on DFMA (4 chunks of 8 encoded inputs) ----
            {
                double acc[4][8][2];
                double accv[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};  // self net: value column only
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[mb][c][0] = acc[mb][c][1] = 0.0;
                // after the env net's first layer the ring region belongs to the weight-digit stream: no DMMA prefetch across that boundary
#define OZ_L0_CHUNK(K0, LAST)                                                                              \
    {                                                                                                      \
        const double* Wc0 = reinterpret_cast<const double*>(d_wait());                                     \
        if (net == 0) oz_layer0_chunk<10, K0>(Wc0, Zs, fr, fq, acc);                                       \
        else oz_layer0_chunk_value<7, K0>(Wc0, Zs, fr, fq, accv);                                          \
        d_done((LAST) && net == 0);                                                                        \
    }
                OZ_L0_CHUNK(0, false)
                OZ_L0_CHUNK(8, false)
                OZ_L0_CHUNK(16, false)
                OZ_L0_CHUNK(24, true)
#undef OZ_L0_CHUNK
                if (warp == 0) { OZ_MARK(1, 5 * net + 2) }
                if (net == 1) {
                    // =============== self net, reverse mode: value forwards, ONE adjoint backwards (SelfCollisionModel.cpp:140-250 computes the same
                    //                 Jacobian row forwards); all operands in shared memory, every product on DMMA ===============
                    double* H1 = Xd + OZ_SF_H1;
                    double* A2 = Xd + OZ_SF_A2;
                    const double* W1s = Xd + OZ_SF_W1;
                    // first-layer adjoint weights of this warp's 32 neurons (used last; requested now, they arrive under the other phases)
                    double w0t[3][8];
#pragma unroll
                    for (int mf = 0; mf < 3; mf++)
#pragma unroll
                        for (int t = 0; t < 8; t++) w0t[mf][t] = __ldg(a.wpack + OZ_DOFF_W0T + ((warp * 3 + mf) * 8 + t) * 32 + lane);
                    // layer 0 epilogue: h1 = relu(pre); the masks stay in registers (the adjoint of the same rows comes back to this lane)
                    bool m1[4][2];
#pragma unroll
                    for (int mb = 0; mb < 4; mb++) {
                        const int row = warp * 32 + mb * 8 + fr;
                        const double bv0 = a.bias[MLP_BIAS_SELF0 + row];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const double pre = accv[mb][e] + bv0;
                            m1[mb][e] = pre > 0.0;
                            H1[oz_bfrag(row, 2 * fq + e)] = m1[mb][e] ? pre : 0.0;
                        }
                    }
                    __syncthreads();  // h1 complete; W1 has landed (every thread waited for its copies with the last first-layer chunk)
                    if (warp == 0) { OZ_MARK(1, 5 * net + 3) }
                    OZ_DBG(0)
                    // layer 1 forwards: warp = m-fragment (neurons 8 warp .. + 7), two accumulator chains over the 64 k-steps
                    {
                        double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
                        const double* wrow = W1s + (warp * 8 + fr) * OZ_SF_W1_LD + fq;
#pragma unroll 8
                        for (int ks = 0; ks < 64; ks += 2) {
                            dmma884(c0[0], c0[1], wrow[4 * ks], H1[ks * 32 + lane]);
                            dmma884(c1[0], c1[1], wrow[4 * ks + 4], H1[(ks + 1) * 32 + lane]);
                        }
                        const int n = warp * 8 + fr;
                        const double bv1 = a.bias[MLP_BIAS_SELF1 + n], wo = __ldg(a.w_out_self + n);
                        double part[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const double pre = c0[e] + c1[e] + bv1;
                            const bool on = pre > 0.0;
                            A2[oz_bfrag(n, 2 * fq + e)] = on ? wo : 0.0;   // adjoint of the layer-1 pre-activation
                            part[e] = on ? wo * pre : 0.0;
                        }
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) { part[0] += __shfl_xor_sync(0xffffffffu, part[0], o); part[1] += __shfl_xor_sync(0xffffffffu, part[1], o); }
                        if (fr == 0) { Xd[OZ_SF_SELP + warp * 8 + 2 * fq] = part[0]; Xd[OZ_SF_SELP + warp * 8 + 2 * fq + 1] = part[1]; }
                    }
                    __syncthreads();  // layer-2 adjoints and the partial outputs are in place; nobody reads h1 any more
                    // layer 1 backwards: adjoint of h1 for this warp's rows 32 warp + 8 mb + fr (A = W1 transposed, read from the same shared copy)
                    {
                        double g[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
                        const double* wcol = W1s + fq * OZ_SF_W1_LD + warp * 32 + fr;
#pragma unroll 4
                        for (int ks = 0; ks < 16; ks++) {
                            const double b = A2[ks * 32 + lane];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++) dmma884(g[mb][0], g[mb][1], wcol[4 * ks * OZ_SF_W1_LD + 8 * mb], b);
                        }
#pragma unroll
                        for (int mb = 0; mb < 4; mb++)
#pragma unroll
                            for (int e = 0; e < 2; e++) H1[oz_bfrag(warp * 32 + mb * 8 + fr, 2 * fq + e)] = m1[mb][e] ? g[mb][e] : 0.0;   // through the ReLU of layer 0
                    }
                    __syncwarp();
                    // layer 0 backwards: this warp's share (its own 32 neurons = 8 k-steps) of the adjoint of the 21 encoded inputs
                    {
                        double z[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                        for (int t = 0; t < 8; t++) {
                            const double b = H1[(warp * 8 + t) * 32 + lane];
#pragma unroll
                            for (int mf = 0; mf < 3; mf++) dmma884(z[mf][0], z[mf][1], w0t[mf][t], b);
                        }
#pragma unroll
                        for (int mf = 0; mf < 3; mf++)
                            *reinterpret_cast<double2*>(Xd + OZ_SF_GZP + (warp * 24 + mf * 8 + fr) * 8 + 2 * fq) = make_double2(z[mf][0], z[mf][1]);
                    }
                    __syncthreads();
                    // output and its gradient: d z / d q is diagonal (x: 1, sin x: cos x, cos x: -sin x); partial sums added in warp order
                    if (tid < 64) {
                        const int sidx = tid & 7, jj = (tid >> 3) - 1, ns = s0 + sidx;
                        if (ns < a.NS) {
                            if (jj < 0) {
                                double v = 0.0;
#pragma unroll
                                for (int w2 = 0; w2 < 8; w2++) v += Xd[OZ_SF_SELP + w2 * 8 + sidx];
                                a.rb[(size_t)RB_SEL * a.NS + ns] = v + a.bias[MLP_BIAS_SELF_OUT];
                            } else {
                                double gx = 0.0, gs = 0.0, gc = 0.0;
#pragma unroll
                                for (int w2 = 0; w2 < 8; w2++) {
                                    gx += Xd[OZ_SF_GZP + (w2 * 24 + jj) * 8 + sidx];
                                    gs += Xd[OZ_SF_GZP + (w2 * 24 + 7 + jj) * 8 + sidx];
                                    gc += Xd[OZ_SF_GZP + (w2 * 24 + 14 + jj) * 8 + sidx];
                                }
                                const double sn = Zs[(7 + jj) * 8 + sidx], cs = Zs[(14 + jj) * 8 + sidx];
                                a.rb[(size_t)(RB_DSEL + jj) * a.NS + ns] = gx + cs * gs - sn * gc;
                            }
                        }
                    }
                    if (warp == 0) { OZ_MARK(1, 5 * net + 4) }
                    OZ_DBG(5)
                    continue;
                }
                const double* bias = a.bias + ((net == 0) ? MLP_BIAS_ENV : MLP_BIAS_SELF0);
                double bv[4];
#pragma unroll
                for (int mb = 0; mb < 4; mb++) bv[mb] = bias[warp * 32 + mb * 8 + fr];
                __syncthreads();  // Zs and (env) the DMMA ring slots are free
                if (net == 0 && producer && !no_stream) oz_produce_range(oa.wq, ring, bar_full, bar_empty, tile_first, tile_first + OZ_NSLOT, tile_first, (uint32_t)(warp - 1), (uint32_t)lane);  // the weight-digit stream of this tile starts
                uint32_t cm[2][8];
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int c = 0; c < 8; c++) cm[e][c] = 0;
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
                            if (net == 0 && on) { cm[e][2 * j] = max(cm[e][2 * j], oz_abs_hi(v0)); cm[e][2 * j + 1] = max(cm[e][2 * j + 1], oz_abs_hi(v1)); }
                        }
                    }
                if (net == 0) {
                    // column maxima over the warp's 32 neurons: butterfly over the 8 lanes (fr) that hold the same two samples, then lane (fr, fq)
                    // publishes two of the 16 columns of its samples -- two conflict-free shared-memory atomics per warp
                    uint32_t keep0 = 0, keep1 = 0;
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            uint32_t m = cm[e][c];
                            m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
                            m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
                            m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
                            if (8 * e + c == 2 * fr) keep0 = m;
                            if (8 * e + c == 2 * fr + 1) keep1 = m;
                        }
                    // column index 2 fr (+1) of this lane's sample pair: e = fr >> 2, c = (2 fr) & 7 (+1)
                    atomicMax(&colmax[8 * (2 * fq + (fr >> 2)) + ((2 * fr) & 7)], keep0);
                    atomicMax(&colmax[8 * (2 * fq + (fr >> 2)) + ((2 * fr) & 7) + 1], keep1);
                }
                if (warp == 0) { OZ_MARK(1, 5 * net + 3) }
                __syncthreads();  // the first layer's output and its column maxima are in place
                if (warp == 0) { OZ_MARK(1, 5 * net + 4) }
                OZ_DBG(0)
            }
            if (net == 0) {
                // =============== env layers 1..3: int8 split on tcgen05 ===============
#pragma unroll 1
                for (int layer = 0; layer < 3; layer++) {
                    oz_split(Xs, planes, colmax, s_sc, s_colscale, tid);
                    OZ_DBG(1)
                    for (int mb = 0; mb < 2; mb++) {
                        const uint32_t pass_first = tile_first + (uint32_t)((layer * 2 + mb) * OZ_CHUNKS_PER_PASS);
                        if (warp == 0) {
                            // ---- MMA issuer ----
                            if (layer == 0) { OZ_MARK(0, 3 * mb) }
                            dbg_wait += oz_issue_pass(tmem, ring_addr, planes_addr, bar_full, bar_empty, bar_group, bar_tfree, pass_first, (uint32_t)(oa.dbg_flags >> 3) & 1u, no_stream, (oa.dbg_flags & 2) != 0);
                            if (layer == 0) { OZ_MARK(0, 3 * mb + 1) }
                        } else if (producer && !no_stream) {
                            // ---- weight-digit stream: the rest of this pass's chunks and the first NSLOT of the next pass's (they land during the epilogue / split) ----
                            const uint32_t pb = pass_first + OZ_NSLOT, pe = (pb + OZ_CHUNKS_PER_PASS < tile_end) ? pb + OZ_CHUNKS_PER_PASS : tile_end;
                            oz_produce_range(oa.wq, ring, bar_full, bar_empty, pb, pe, tile_first, (uint32_t)(warp - 1), (uint32_t)lane);
                        }
                        if (warp == 0) oz_mbar_wait(bar_group + 8 * (OZ_S - 1), n_pass & 1u);  // the last plane's commit: every MMA of the pass is done
                        if (warp == 0 && layer == 0) { OZ_MARK(0, 3 * mb + 2) }
                        n_pass++;
                        __syncthreads();  // the other warps wait in the hardware barrier, not by polling
                        if (layer == 2 && mb == 1) {  // the digit planes are dead: the output layer's weights (A fragments) land above the tile during the epilogue
                            const double2* src = reinterpret_cast<const double2*>(a.wpack + OZ_DOFF_WOUT);
                            for (int i = tid; i < OZ_WOUT_D / 2; i += MLP_THREADS) cp_async16(reinterpret_cast<double2*>(smem_raw + OZ_OFF_WOUT) + i, src + i);
                            cp_async_commit();
                        }
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        OZ_DBG(2)
                        oz_epilogue_pass(Xs, colmax, s_colscale, oa.rowscale + layer * 256, a.bias + MLP_BIAS_ENV + (layer + 1) * 256, tmem, mb, warp, lane, layer < 2);
                        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                        __syncthreads();  // accumulators read, tile rows written (the next pass / split / output layer may start)
                        OZ_DBG(3)
                    }
                }
                cp_async_wait_all();
                __syncthreads();  // the output layer's weights are in place
                d_topup();  // all MMAs of the tile are done: the ring region is the DMMA layers' again
                // ---- env output layer (9 x 256); warp = column kind.  Rows 0..7: one m-fragment on DMMA.  Row 8 alone would waste seven eighths of a
                //      second fragment (2 k cycles of the FP64 pipe per tile): it rides along as DFMA on the B-fragment elements the lane holds
                //      anyway (k = 4 kb + fq of sample fr), summed over fq at the end ----
                double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // (four accumulator chains instead of two: no change -- the layer is not bound by the DMMA latency)
                double r8[2] = {0.0, 0.0};
                const double bias_fr = (warp == 0) ? __ldg(a.bias + MLP_BIAS_ENV_OUT + fr) : 0.0, bias_8 = (warp == 0) ? __ldg(a.bias + MLP_BIAS_ENV_OUT + 8) : 0.0;   // requested before the loop
                const double* xb = Xd + ((warp >> 1) * 32 + bslot) * 2 + (warp & 1);
                const double* Wo0 = reinterpret_cast<const double*>(smem_raw + OZ_OFF_WOUT);  // rows 0..7 as A fragments [kb 64][lane 32]
                const double* Wo1 = Wo0 + 8 * 256;                                            // row 8 [256]
#pragma unroll 4
                for (int kb = 0; kb < 64; kb += 2) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const double bx = xb[(kb + h) * 256];
                        dmma884(o[h][0], o[h][1], Wo0[(kb + h) * 32 + lane], bx);
                        r8[h] = fma(Wo1[(kb + h) * 4 + fq], bx, r8[h]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int ns = s0 + 2 * fq + e;
                    if (ns < a.NS) {
                        const double v = o[0][e] + o[1][e];
                        if (warp == 0) a.rb[(size_t)(RB_ENV + fr) * a.NS + ns] = v + bias_fr;
                        else a.rb[(size_t)(RB_DENV + fr * 7 + (warp - 1)) * a.NS + ns] = v;
                    }
                }
                {
                    double v = r8[0] + r8[1];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int ns = s0 + fr;   // the lane's B-fragment column is sample fr
                    if (fq == 0 && ns < a.NS) {
                        if (warp == 0) a.rb[(size_t)(RB_ENV + 8) * a.NS + ns] = v + bias_8;
                        else a.rb[(size_t)(RB_DENV + 8 * 7 + (warp - 1)) * a.NS + ns] = v;
                    }
                }
                if (tid < 8 && (s0 + tid) < a.NS) {
                    const int n2 = s0 + tid;
                    a.rb[(size_t)RB_OBSR * a.NS + n2] = a.obs[(size_t)(n2 / a.S) * 4 + 3];
                }
                OZ_DBG(4)
            }
        }
    }
    if (dbg_on) { for (int i = 0; i < 7; i++) oa.dbg[i] = dbg_t[i]; oa.dbg[7] = dbg_wait; }
#undef OZ_DBG
    cp_async_wait_all();
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512) : "memory");
}

#endif  // __CUDACC__

// Host-side packing for k_mlp_oz.
//   dpack  (OZ_DPACK_D doubles): env L0 as 4 chunks [warp 8][k 8][row 32] | self L0 likewise | env output layer (rows 0..7 as A fragments | row 8) |
//          self L1 row-major [64][256] | self L0 transposed as A fragments [warp 8][m-fragment 3][k-step 8][lane 32]
//   qpack  (OZ_CHUNKS_PER_TILE x OZ_CHUNK bytes): for env layers 1..3, neuron halves mb, plane positions pi (plane oz_order(pi)), k chunks kc:
//          128 rows x 64 k of digit plane i as canonical K-major 8 x 16-byte core matrices: byte (r % 8) 16 + (r / 8) 128 + (k % 16) + (k / 16) 2048
//   rowscale [3][256] = 2^(E_r - 7 S - 7),   2^(E_r) > 2 max_k |W[r,k]|
inline void pack_mlp_oz_weights(const double* const env_W[5], const double* const self_W[3], double* dpack, uint8_t* qpack, double* rowscale) {
    auto pack_l0 = [&](const double* W, int in_dim, double*& o) {
        for (int k0 = 0; k0 < 32; k0 += 8)
            for (int warp = 0; warp < 8; warp++)
                for (int kk = 0; kk < 8; kk++)
                    for (int row = 0; row < 32; row++) {
                        const int k = k0 + kk;
                        *o++ = (k < in_dim) ? W[(size_t)(32 * warp + row) * in_dim + k] : 0.0;
                    }
    };
    double* o = dpack;
    pack_l0(env_W[0], 30, o);
    pack_l0(self_W[0], 21, o);
    for (int kb = 0; kb < 64; kb++)
        for (int l = 0; l < 32; l++) *o++ = env_W[4][(size_t)(l >> 2) * 256 + kb * 4 + (l & 3)];
    for (int k = 0; k < 256; k++) *o++ = env_W[4][(size_t)8 * 256 + k];
    for (int i = 0; i < 64 * 256; i++) *o++ = self_W[1][i];
    for (int warp = 0; warp < 8; warp++)
        for (int mf = 0; mf < 3; mf++)
            for (int t = 0; t < 8; t++)
                for (int l = 0; l < 32; l++) {
                    const int r = 8 * mf + (l >> 2), i = 32 * warp + 4 * t + (l & 3);   // A fragment of the transpose: row = encoded input r, k = neuron i
                    *o++ = (r < 21) ? self_W[0][(size_t)i * 21 + r] : 0.0;
                }
    for (int layer = 0; layer < 3; layer++) {
        const double* W = env_W[layer + 1];
        for (int r = 0; r < 256; r++) {
            double mx = 0.0;
            for (int k = 0; k < 256; k++) mx = std::fmax(mx, std::fabs(W[(size_t)r * 256 + k]));
            int e = 0;
            if (mx > 0.0) std::frexp(mx, &e);  // mx < 2^e
            const int Er = e + 1;
            rowscale[layer * 256 + r] = std::ldexp(1.0, Er - 7 * OZ_S - 7);
            for (int k = 0; k < 256; k++) {
                const long long q = std::llrint(std::ldexp(W[(size_t)r * 256 + k], 7 * OZ_S - Er)) + oz_bias_const();
                for (int t = 0; t < OZ_S; t++) qpack[oz_wq_index(layer, r, k, OZ_S - 1 - t)] = (uint8_t)(int8_t)oz_digit(q, t);
            }
        }
    }
}

}  // namespace mpcc
