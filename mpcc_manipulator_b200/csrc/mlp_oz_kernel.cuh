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

namespace mpcc {

constexpr int OZ_S = 7;                                // int8 digits per operand
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
constexpr int OZ_MISC_COLMAX = 0, OZ_MISC_SC = 256, OZ_MISC_COLSCALE = 768, OZ_MISC_BARS = 1280, OZ_MISC_TMEM = 1280 + 8 * (2 * OZ_NSLOT + 1);
constexpr size_t OZ_SMEM_BYTES = OZ_OFF_MISC + 1536;
static_assert(OZ_S >= 4 && OZ_S <= 7, "int64 Horner of the accumulators holds up to 7 digits");
static_assert(OZ_S * 64 + 4 * (OZ_CHUNK_BYTES / 512) <= 512, "accumulators and the four A buffers must fit TMEM");
static_assert(OZ_RING_BYTES >= 32768, "the DMMA layers' per-warp rings live in the ring region");
static_assert(OZ_SMEM_BYTES <= 232448, "shared memory");
// DMMA chunk stream of one tile: env L0 (4 chunks of 8 encoded inputs) | self L0 (4) | self L1 (8 chunks of 32 k-steps x 64 neurons); 16 KB each
constexpr int OZ_DCHUNK_D = 2048;
constexpr int OZ_NDCHUNK = 16;
// weight plane consumed at position pi of a pass: heavy (low i: S - i products) and light planes alternate, so the stream's demand is even
constexpr int oz_order(int pi) { return (pi & 1) ? OZ_S - 1 - pi / 2 : pi / 2; }
constexpr long long oz_bias_const() {  // 64 on each of the lower S - 1 digits: makes them unsigned fields (the top digit stays signed)
    long long c = 0;
    for (int t = 0; t < OZ_S - 1; t++) c += 64LL << (7 * t);
    return c;
}

struct MlpOzArgs {
    MlpArgs m;               // m.wpack: the DMMA chunk stream packed by pack_mlp_oz_weights (OZ_NDCHUNK x OZ_DCHUNK_D doubles)
    const uint8_t* wq;       // OZ_CHUNKS_PER_TILE x OZ_CHUNK bytes: digit planes of env layers 1..3 in consumption order
    const double* rowscale;  // [3][256]: 2^(E_r - 7 S - 7)
    int dbg_flags;           // experiments (timing only, results wrong): bit 0 = no weight stream (MMAs on whatever the ring holds), bit 1 = no per-CTA rotation of the stream
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


// One MMA pass, issued by one thread: the S (S + 1) / 2 digit products of 128 neurons x 64 columns x 256 k, chunk by chunk as the weight digits
// arrive.  Out of line on purpose: the kernel around it runs at the register limit, and a spill reloaded inside this loop costs more than the
// 53 cycles an MMA takes.  Returns the cycles spent waiting for chunks.
__device__ __noinline__ long long oz_issue_pass(uint32_t tmem, uint32_t ring_addr, uint32_t planes_addr, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_acc, uint32_t n,
                                               uint32_t rot, bool no_stream) {
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // s32 += s8 (K-major) x s8 (MN-major), M 128, N 64
    // called by the whole of warp 0, converged: the election inside tells the compiler that exactly one thread issues (no per-MMA election loop)
    long long waited = 0;
    // warp-uniform copies of the arguments: tcgen05.mma takes its operands from uniform registers, and values the compiler cannot prove uniform
    // cost an election loop around every MMA (~40 cycles each)
    tmem = __shfl_sync(0xffffffffu, tmem, 0); ring_addr = __shfl_sync(0xffffffffu, ring_addr, 0); planes_addr = __shfl_sync(0xffffffffu, planes_addr, 0);
    bar_full = __shfl_sync(0xffffffffu, bar_full, 0); bar_empty = __shfl_sync(0xffffffffu, bar_empty, 0); bar_acc = __shfl_sync(0xffffffffu, bar_acc, 0);
    n = __shfl_sync(0xffffffffu, n, 0); rot = __shfl_sync(0xffffffffu, rot, 0);
    if (oz_elect_one()) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t self_mask = 1u << (threadIdx.x & 31);
    uint32_t touched = 0;
    uint32_t posr = rot;
    const uint64_t bd0 = oz_desc(planes_addr, 128, 4096);
    for (uint32_t pos = 0; pos < (uint32_t)OZ_CHUNKS_PER_PASS; pos++, n++) {
        const int i = oz_order((int)(posr / OZ_CPP)), kc = (int)(posr % OZ_CPP);
        posr = (posr + 1 == (uint32_t)OZ_CHUNKS_PER_PASS) ? 0u : posr + 1;
        const uint32_t slot = n % OZ_NSLOT;
        if (!no_stream) oz_mbar_wait_asm(bar_full + 8 * slot, (n / OZ_NSLOT) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint64_t ad = oz_desc(ring_addr + slot * OZ_CHUNK, 2048, 128);
        uint64_t bd = bd0 + (uint64_t)((kc * (OZ_KCH * 16)) >> 4);
        if (oz_elect_mask(self_mask)) {  // re-establishes "one thread" for the compiler after the wait loop (see oz_mbar_wait_asm)
            // the chunk goes to TMEM once (tcgen05.cp, 8 columns per 32-k step; 4 rotating buffers behind the accumulators) and the MMAs take A from
            // there: shared memory is read once per chunk instead of once per MMA (4 KB each), and the ring slot is free as soon as the copy is done.
            // Copies and MMAs execute in issue order, so a buffer is not overwritten before the MMAs issued earlier have read it.
            const uint32_t ta = tmem + OZ_S * 64 + (n & 3u) * (OZ_KCH / 4);
#pragma unroll
            for (int ks = 0; ks < OZ_KCH / 32; ks++) oz_utccp(ta + ks * 8, ad + (uint64_t)((ks * 4096) >> 4));
            oz_commit(bar_empty + 8 * slot);
            for (int g = i; g < OZ_S; g++, bd += (uint64_t)(OZ_PLANE >> 4)) {  // planes j = 0 .. S-1-i, accumulator g = i + j
                oz_mma_i8_ts(tmem + g * 64, ta, bd, IDESC, (touched >> g) & 1u);
#pragma unroll
                for (int ks = 1; ks < OZ_KCH / 32; ks++) oz_mma_i8_ts(tmem + g * 64, ta + ks * 8, bd + (uint64_t)((ks * 512) >> 4), IDESC, 1u);
                touched |= 1u << g;
            }
            if (pos + 1 == (uint32_t)OZ_CHUNKS_PER_PASS) oz_commit(bar_acc);
        }
    }
    }
    __syncwarp();
    return waited;
}

__device__ __forceinline__ uint32_t oz_abs_hi(double v) { return (uint32_t)__double2hiint(v) & 0x7FFFFFFFu; }

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
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc + OZ_MISC_BARS);       // full[NSLOT] | empty[NSLOT] | accumulators ready
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + OZ_MISC_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fq = lane & 3;
    const int bslot = xslot(fq, fr);
    const uint32_t bar_full = oz_smem_u32(bars), bar_empty = oz_smem_u32(bars + OZ_NSLOT), bar_acc = oz_smem_u32(bars + 2 * OZ_NSLOT);

    if (tid == 0) {
        for (int i = 0; i < 2 * OZ_NSLOT + 1; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_full + 8 * i), "r"(i < OZ_NSLOT ? 32 : 1) : "memory");
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

    // ---- per-warp rings of the DMMA layers (k_mlp's scheme, 2 x 2 KB per warp, inside the ring region) ----
    int p = 0, buf = 0;
    constexpr int SLICE2 = OZ_DCHUNK_D / 2 / 8;  // double2 per warp and chunk
    double2* Wmine = reinterpret_cast<double2*>(ring) + warp * (2 * SLICE2);
    auto prefetch = [&](int chunk, int slot) {
        const double2* src = reinterpret_cast<const double2*>(a.wpack) + ((size_t)chunk * 8 + warp) * SLICE2;
        double2* dst = Wmine + slot * SLICE2;
#pragma unroll
        for (int i = 0; i < SLICE2 / 32; i++) cp_async16(dst + lane + i * 32, src + lane + i * 32);
        cp_async_commit();
    };
    auto next_chunk = [&](bool prefetch_next) {
        cp_async_wait_all();
        __syncwarp();
        if (prefetch_next) prefetch((p + 1) % OZ_NDCHUNK, buf ^ 1);
    };
    auto advance = [&]() { buf ^= 1; p = (p + 1) % OZ_NDCHUNK; };
    if ((int)blockIdx.x < a.n_tiles) prefetch(0, 0);

    // ---- weight-digit stream (used by the elected lane of warp 0 only) ----
    uint32_t tile_iter = 0;   // tiles this CTA has started: chunk numbers of the stream follow from it (no per-thread state: any lane may be elected)
    uint32_t n_pass = 0;      // passes waited for so far (parity of the accumulator barrier; every thread counts)
    const uint32_t ring_addr = oz_smem_u32(ring), planes_addr = oz_smem_u32(planes);
    // Every CTA walks the chunks of a pass from its own starting point (the integer accumulation is exact in any order), so the CTAs do not
    // all ask L2 for the same lines at the same moment.
    const uint32_t rot = (oa.dbg_flags & 2) ? 0u : (blockIdx.x * 11u) % OZ_CHUNKS_PER_PASS;
    auto src_chunk = [&](uint32_t ml) -> uint32_t { return (ml / OZ_CHUNKS_PER_PASS) * OZ_CHUNKS_PER_PASS + (ml % OZ_CHUNKS_PER_PASS + rot) % OZ_CHUNKS_PER_PASS; };
    const bool no_stream = (oa.dbg_flags & 1) != 0;
    // Producers: warps 1..NSLOT copy the chunks with 16-byte cp.async (LDGSTS); warp 1 + k owns ring slot k, i.e. the chunks m = k (mod NSLOT) of
    // the stream, one chunk in flight per warp.  A chunk is handed to the tensor core by its 32 copying lanes: own copies landed (wait_group 0),
    // proxy fence (generic-proxy writes -> async-proxy reads), arrive on the slot's `full` barrier.  Measured alternatives: cp.async.bulk (one
    // 8 / 16 KB bulk copy per chunk completes every ~1300 / ~1700 cycles however many are outstanding: 6 .. 10 B/cycle, a third of what the MMAs
    // consume), and 128 threads sharing every chunk with 5 copy groups in flight per thread (the proxy fence then waits for the younger groups too).
    uint32_t pm = 0;      // next chunk of the stream this producer warp copies
    bool pend = false;    // a chunk of this warp is in flight (its slot: (pm - NSLOT) % NSLOT = warp - 1)
    const bool producer = warp >= 1 && warp <= OZ_NSLOT;
    auto produce = [&](uint32_t limit, bool final_call) {
        const uint32_t slot = warp - 1;
        while (true) {
            if (pend) {
                cp_async_wait_all();
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar_full + 8 * slot) : "memory");
                pend = false;
            }
            if (pm >= limit) break;
            const uint32_t use = pm / OZ_NSLOT;
            if (use > 0) oz_mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);  // the MMAs of the slot's previous tenant are done
            const unsigned char* src = oa.wq + (size_t)src_chunk(pm % OZ_CHUNKS_PER_TILE) * OZ_CHUNK + lane * 16;
            unsigned char* dst = ring + slot * OZ_CHUNK + lane * 16;
#pragma unroll
            for (int i = 0; i < OZ_CHUNK / 512; i++) cp_async16(dst + i * 512, src + i * 512);
            cp_async_commit();
            pend = true;
            pm += OZ_NSLOT;
            if (pm >= limit && !final_call) break;  // the last copy stays in flight across the epilogue; the next call completes it first
        }
    };
    long long dbg_t[7] = {0, 0, 0, 0, 0, 0, 0}, dbg_c = 0, dbg_wait = 0;
    const bool dbg_on = oa.dbg != nullptr && blockIdx.x == 0 && tid == 0;
#define OZ_DBG(i) if (dbg_on) { const long long t_ = clock64(); dbg_t[i] += t_ - dbg_c; dbg_c = t_; }
    if (dbg_on) dbg_c = clock64();
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, tile_iter++) {
        const int s0 = tile * MLP_TILE_S;
        dbg_t[6]++;
        const uint32_t tile_first = tile_iter * OZ_CHUNKS_PER_TILE, tile_end = tile_first + OZ_CHUNKS_PER_TILE;
        for (int net = 0; net < 2; net++) {  // 0: env, 1: self
            // ---- encoded inputs z = [x, sin x, cos x] of the 8 samples: Zs [k 32][sample 8] (in the idle X tile) ----
            __syncthreads();
            double* Zs = Xd;
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
            // ---- first layer on DFMA (4 chunks of 8 encoded inputs) ----
            {
                double acc[4][8][2];
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[mb][c][0] = acc[mb][c][1] = 0.0;
                // after the env net's first layer the ring region belongs to the weight-digit stream: no DMMA prefetch across that boundary
#define OZ_L0_CHUNK(K0, LAST)                                                                                            \
    next_chunk(!((LAST) && net == 0));                                                                                   \
    if (net == 0) oz_layer0_chunk<10, K0>(reinterpret_cast<const double*>(Wmine + buf * SLICE2), Zs, fr, fq, acc);         \
    else oz_layer0_chunk<7, K0>(reinterpret_cast<const double*>(Wmine + buf * SLICE2), Zs, fr, fq, acc);                   \
    advance();
                OZ_L0_CHUNK(0, false)
                OZ_L0_CHUNK(8, false)
                OZ_L0_CHUNK(16, false)
                OZ_L0_CHUNK(24, true)
#undef OZ_L0_CHUNK
                const double* bias = a.bias + ((net == 0) ? MLP_BIAS_ENV : MLP_BIAS_SELF0);
                double bv[4];
#pragma unroll
                for (int mb = 0; mb < 4; mb++) bv[mb] = bias[warp * 32 + mb * 8 + fr];
                __syncthreads();  // Zs and (env) the DMMA ring slots are free
                if (net == 0 && producer && !no_stream) {  // the weight-digit stream of this tile starts: one chunk per producer warp
                    pm = tile_first + (uint32_t)(warp - 1);
                    produce(tile_first + OZ_NSLOT, false);
                }
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
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const uint32_t m = __reduce_max_sync(0x11111111u << fq, cm[e][c]);  // the 8 lanes (fr) that hold this sample
                            if (fr == 0) atomicMax(&colmax[8 * (2 * fq + e) + c], m);
                        }
                }
                __syncthreads();  // the first layer's output and its column maxima are in place
                OZ_DBG(0)
            }
            if (net == 0) {
                // =============== env layers 1..3: int8 split on tcgen05 ===============
                for (int layer = 0; layer < 3; layer++) {
                    // ---- split: fp64 tile -> S digit planes ----
                    if (tid < 64) {
                        const uint32_t hi = colmax[tid];
                        const int be = (int)(hi >> 20);  // biased exponent of the column's largest entry
                        double sc = 0.0, cs = 1.0;
                        if (be >= 7 * OZ_S && be <= 2040) {
                            sc = __hiloint2double((2044 + 7 * OZ_S - be) << 20, 0);  // 2^(7 S - e_c), e_c = be - 1021
                            cs = __hiloint2double((be + 2) << 20, 0);                // 2^(e_c)
                        }
                        s_sc[tid] = sc;
                        s_colscale[tid] = cs;
                        colmax[tid] = 0;
                    }
                    __syncthreads();
                    {
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
                                    const long long q = __double2ll_rn((h ? x2.y : x2.x) * s_sc[c]) + oz_bias_const();
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
                        OZ_DBG(1)
                    }
                    for (int mb = 0; mb < 2; mb++) {
                        // ---- MMA pass: neurons 128 mb .. 128 mb + 127 ----
                        if (warp == 0) {
                            dbg_wait += oz_issue_pass(tmem, ring_addr, planes_addr, bar_full, bar_empty, bar_acc,
                                                      tile_first + (uint32_t)((layer * 2 + mb) * OZ_CHUNKS_PER_PASS), rot, no_stream);
                        } else if (producer && !no_stream) {
                            const uint32_t pass_end = tile_first + (uint32_t)((layer * 2 + mb + 1) * OZ_CHUNKS_PER_PASS);
                            if (pass_end == tile_end) produce(tile_end, true);
                            else produce(pass_end + OZ_NSLOT, false);  // incl. this warp's first chunk of the next pass
                        }
                        // ---- epilogue ----
                        oz_mbar_wait(bar_acc, n_pass & 1);
                        n_pass++;
                        OZ_DBG(2)
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        if (layer == 2 && mb == 1) prefetch(p, buf);  // all MMAs of the tile are done: the ring region is the DMMA layers' again
                        {
                            const int qd = warp & 3, hh = warp >> 2;
                            const int row = mb * 128 + qd * 32 + lane;
                            long long acc[32];
#pragma unroll
                            for (int g = 0; g < OZ_S; g++) {
                                uint32_t v[32];
                                oz_tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + g * 64 + hh * 32, v);
#pragma unroll
                                for (int c = 0; c < 32; c++) acc[c] = (g == 0) ? (long long)(int)v[c] : acc[c] * 128 + (long long)(int)v[c];
                            }
                            const double rs = __ldg(oa.rowscale + layer * 256 + row), bv = __ldg(a.bias + MLP_BIAS_ENV + (layer + 1) * 256 + row);
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
                                if (layer < 2) {
#pragma unroll
                                    for (int c = 0; c < 8; c++) {
                                        const uint32_t m = __reduce_max_sync(0xffffffffu, on ? oz_abs_hi(y[c]) : 0u);
                                        if (lane == 0) atomicMax(&colmax[8 * s + c], m);
                                    }
                                }
                            }
                        }
                        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                        __syncthreads();  // accumulators read, tile rows written (the next pass / split / output layer may start)
                        OZ_DBG(3)
                    }
                }
                // ---- env output layer: 9 x 256 as two m-fragments (rows 0..7, row 8) on DMMA; warp = column kind ----
                double o[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
                const double* xb = Xd + ((warp >> 1) * 32 + bslot) * 2 + (warp & 1);
#pragma unroll 4
                for (int kb = 0; kb < 64; kb += 2) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const double bx = xb[(kb + h) * 256];
                        const double a0 = __ldg(a.w_out_env + fr * 256 + (kb + h) * 4 + fq), a1 = (lane < 4) ? __ldg(a.w_out_env + 8 * 256 + (kb + h) * 4 + lane) : 0.0;
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
                OZ_DBG(4)
            } else {
                // ---- self layer 1: 64 x 256 on DMMA; warp = m-fragment (8 neurons) x all 8 column kinds, full K; 8 chunks of 32 k-steps ----
                double acc[8][2];
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c][0] = acc[c][1] = 0.0;
                for (int ch = 0; ch < 8; ch++) {
                    next_chunk(true);
                    const double2* Wc = Wmine + buf * SLICE2;  // [kb pair 4][lane 32] -> {kb even, kb odd}
#pragma unroll 2
                    for (int kp = 0; kp < 4; kp++) {
                        const double2 a2 = Wc[kp * 32 + lane];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int kb = ch * 8 + kp * 2 + h;
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
                __syncthreads();
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
                OZ_DBG(5)
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
//   dpack  (OZ_NDCHUNK x OZ_DCHUNK_D doubles): env L0 as 4 chunks [warp 8][k 8][row 32] | self L0 likewise | self L1 as 8 chunks
//          [warp 8][kb pair 4][lane 32][h 2] = W[8 warp + (l >> 2)][32 ch + 4 (2 kp + h) + (l & 3)]
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
    for (int ch = 0; ch < 8; ch++)
        for (int warp = 0; warp < 8; warp++)
            for (int kp = 0; kp < 4; kp++)
                for (int l = 0; l < 32; l++)
                    for (int h = 0; h < 2; h++) {
                        const int row = 8 * warp + (l >> 2), k = 32 * ch + 4 * (2 * kp + h) + (l & 3);
                        *o++ = self_W[1][(size_t)row * 256 + k];
                    }
    int pos_of_plane[OZ_S];
    for (int pi = 0; pi < OZ_S; pi++) pos_of_plane[oz_order(pi)] = pi;
    for (int layer = 0; layer < 3; layer++) {
        const double* W = env_W[layer + 1];
        for (int r = 0; r < 256; r++) {
            double mx = 0.0;
            for (int k = 0; k < 256; k++) mx = std::fmax(mx, std::fabs(W[(size_t)r * 256 + k]));
            int e = 0;
            if (mx > 0.0) std::frexp(mx, &e);  // mx < 2^e
            const int Er = e + 1;
            rowscale[layer * 256 + r] = std::ldexp(1.0, Er - 7 * OZ_S - 7);
            const int mb = r / 128, rr = r % 128;
            for (int k = 0; k < 256; k++) {
                const long long q = std::llrint(std::ldexp(W[(size_t)r * 256 + k], 7 * OZ_S - Er)) + oz_bias_const();
                for (int t = 0; t < OZ_S; t++) {
                    const int i = OZ_S - 1 - t;
                    const int d = (t < OZ_S - 1) ? (int)((q >> (7 * t)) & 127) - 64 : (int)(q >> (7 * (OZ_S - 1)));
                    const size_t chunk = (size_t)layer * OZ_CHUNKS_PER_LAYER + (size_t)(mb * OZ_S + pos_of_plane[i]) * OZ_CPP + k / OZ_KCH;
                    const int kk = k % OZ_KCH;
                    qpack[chunk * OZ_CHUNK + (rr % 8) * 16 + (rr / 8) * 128 + (kk % 16) + (kk / 16) * 2048] = (uint8_t)(int8_t)d;
                }
            }
        }
    }
}

}  // namespace mpcc
