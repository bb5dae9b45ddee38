// Probe for the int8-split ("Ozaki") route of the 256 x 256 MLP layers on tcgen05 (kind::i8, s8 x s8 -> s32 in TMEM).
//   1. which shared-memory descriptor convention (LBO / SBO roles) reproduces a CPU int32 product, for A K-major and for
//      B MN-major / K-major, no swizzle (8 x 16-byte core matrices);
//   2. cycles per tcgen05.mma at M = 128, N = 64 / 128 / 256, K = 32 with both operands resident in shared memory
//      (is the N = 64 shape shared-memory-bound?), and the TMEM read-back rate of tcgen05.ld.32x32b.x32.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o oz_umma_probe oz_umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Prod { int a_off, b_off, d_col, accumulate; };  // byte offsets of the operand start inside the A / B regions, TMEM column, enable_input_d
constexpr int MAX_PROD = 512;
__constant__ Prod c_prod[MAX_PROD];

struct Params {
    int a_bytes, b_bytes;       // shared-memory regions (A first, then B)
    int a_lbo, a_sbo, b_lbo, b_sbo;  // descriptor fields, bytes
    int b_major;                // 0: K-major, 1: MN-major
    int n;                      // MMA N
    int n_prod;                 // products per repetition
    int reps;                   // repetitions (timing)
    int d_cols;                 // TMEM columns to read back
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int lbo, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}\n" : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_probe(const uint8_t* __restrict__ Ag, const uint8_t* __restrict__ Bg, int32_t* __restrict__ Dg, long long* __restrict__ cycles, Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* As = smem;
    unsigned char* Bs = smem + p.a_bytes;
    for (int i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(As + i) = *reinterpret_cast<const uint4*>(Ag + i);
    for (int i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(Bs + i) = *reinterpret_cast<const uint4*>(Bg + i);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> async-proxy (tensor core) reads
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.b_major << 16) | ((uint32_t)(p.n >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t parity = 0;
    long long t_total = 0;
    for (int rep = 0; rep < p.reps; rep++) {
        long long t0 = 0;
        if (threadIdx.x == 0) {
            t0 = clock64();
            const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
            for (int i = 0; i < p.n_prod; i++) {
                const Prod q = c_prod[i];
                mma_i8(tmem_base + q.d_col, make_desc(a0 + q.a_off, p.a_lbo, p.a_sbo), make_desc(b0 + q.b_off, p.b_lbo, p.b_sbo), idesc, q.accumulate);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1;
        if (threadIdx.x == 0) t_total += clock64() - t0;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    if (threadIdx.x == 0) cycles[0] = t_total;
    // read back: warp w owns lanes 32 w .. 32 w + 31
    long long t1 = clock64();
    for (int c0 = 0; c0 < p.d_cols; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
              "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++) Dg[(size_t)(warp * 32 + lane) * p.d_cols + c0 + j] = (int32_t)v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[1] = clock64() - t1;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}


// Rate kernel: the issue loop of one pass of an S-slice layer with everything but the descriptor start addresses hoisted
// (S (S + 1) / 2 products x 8 k-steps; A chunks of 128 x 128 in a 3-slot ring, B planes [N x 256] MN-major), then the TMEM read-back
// timed on its own (tcgen05.ld.32x32b.x32 + wait, XOR-reduced so that nothing but the loads is in the timed region).
template <int N, int S>
__global__ void __launch_bounds__(128, 1) k_rate(const uint8_t* __restrict__ Ag, const uint8_t* __restrict__ Bg, long long* __restrict__ cycles, int a_bytes, int b_bytes, int reps, uint32_t* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* As = smem;
    unsigned char* Bs = smem + a_bytes;
    for (int i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(As + i) = *reinterpret_cast<const uint4*>(Ag + i);
    for (int i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(Bs + i) = *reinterpret_cast<const uint4*>(Bg + i);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    constexpr int GROUPS = (512 / N < S) ? 512 / N : S;
    uint32_t parity = 0;
    long long t_total = 0;
    for (int rep = 0; rep < reps; rep++) {
        long long t0 = clock64();
        if (warp == 0 && elect_one()) {
            const uint64_t ad0 = make_desc(smem_u32(As), 2048, 128), bd0 = make_desc(smem_u32(Bs), 128, 32 * 128);
            int chunk = 0;
            for (int i = 0; i < S; i++)
                for (int kc = 0; kc < 2; kc++, chunk++) {
                    const uint64_t ad = ad0 + (uint64_t)(((chunk % 3) * 16384) >> 4);
                    for (int j = 0; j + i < S; j++) {
                        const uint64_t bd = bd0 + (uint64_t)(((j % (b_bytes / (N * 256))) * N * 256 + kc * 4 * 512) >> 4);
                        const uint32_t d = tmem_base + ((i + j) % GROUPS) * N;
                        const uint32_t acc0 = (i > 0 || kc > 0 || (i + j) >= GROUPS) ? 1u : 0u;
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) mma_i8(d, ad + (uint64_t)((ks * 4096) >> 4), bd + (uint64_t)((ks * 512) >> 4), idesc, ks ? 1u : acc0);
                    }
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1;
        if (threadIdx.x == 0) t_total += clock64() - t0;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    if (threadIdx.x == 0) cycles[0] = t_total;
    __syncthreads();
    long long t1 = clock64();
    uint32_t x = 0;
    for (int c0 = 0; c0 < 512; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
              "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++) x ^= v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[1] = clock64() - t1;
    sink[threadIdx.x] = x;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int N, int S>
static void rate(int planesB) {
    const int a_bytes = 3 * 16384, b_bytes = planesB * N * 256;
    std::vector<uint8_t> A(a_bytes), B(b_bytes);
    for (auto& v : A) v = (uint8_t)(rand() % 129 - 64);
    for (auto& v : B) v = (uint8_t)(rand() % 129 - 64);
    uint8_t *dA, *dB; long long* dC; uint32_t* dS;
    CK(cudaMalloc(&dA, a_bytes)); CK(cudaMalloc(&dB, b_bytes)); CK(cudaMalloc(&dC, 16)); CK(cudaMalloc(&dS, 512));
    CK(cudaMemcpy(dA, A.data(), a_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), b_bytes, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_rate<N, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, a_bytes + b_bytes));
    const int reps = 8, n_mma = S * (S + 1) / 2 * 8;
    k_rate<N, S><<<1, 128, a_bytes + b_bytes>>>(dA, dB, dC, a_bytes, b_bytes, reps, dS);
    CK(cudaDeviceSynchronize());
    long long c[2];
    CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
    const double per = (double)c[0] / (reps * (double)n_mma);
    printf("rate: M=128 N=%d K=32 int8, S=%d: %lld cycles for %d x %d MMAs -> %.1f cycles / MMA (tensor floor %d; operands %d B/MMA -> %.0f B/cycle); TMEM read-back of 512 columns x 128 lanes by 4 warps: %lld cycles (%.0f B/cycle)\n",
           N, S, c[0], reps, n_mma, per, N / 2, 4096 + N * 32, (4096.0 + N * 32) / per, c[1], 128.0 * 512 * 4 / (double)c[1]);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS);
}

// A from TMEM: tcgen05.cp.128x256b stages one k-step of A (128 rows x 32 bytes, the same canonical K-major core matrices and descriptor as the
// shared-memory operand) into 8 TMEM columns, and the MMA takes [tmem] as its A operand.  Correctness against the CPU product, then the rate of
// one pass of an S-slice layer: per 8 KB chunk two copies (16 columns, 4 rotating buffers beside the S x 64 accumulator columns) and 2 (S - i) MMAs.
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void utccp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
template <int S>
__global__ void __launch_bounds__(512, 1) k_ts(const uint8_t* __restrict__ Ag, const uint8_t* __restrict__ Bg, int32_t* __restrict__ Dg, uint32_t* __restrict__ Adump,
                                               long long* __restrict__ cycles, int a_bytes, int b_bytes, int mode, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    unsigned char* As = (mode & 16) ? smem + 180224 : smem;
    unsigned char* Bs = (mode & 16) ? smem + 65536 : smem + a_bytes;
    for (int i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(As + i) = *reinterpret_cast<const uint4*>(Ag + i);
    for (int i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(Bs + i) = *reinterpret_cast<const uint4*>(Bg + i);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar2)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t tmem_a0 = tmem_base + 448;
    uint32_t parity = 0;
    if ((mode & 1) == 0) {  // one product, K = 64: two copies + two MMAs
        if (warp == 0 && elect_one()) {
            const uint64_t ad0 = make_desc(smem_u32(As), 2048, 128), bd0 = make_desc(smem_u32(Bs), 128, 8 * 128);
            for (int ks = 0; ks < 2; ks++) utccp_128x256b(tmem_a0 + ks * 8, ad0 + (uint64_t)((ks * 4096) >> 4));
            for (int ks = 0; ks < 2; ks++) mma_i8_ts(tmem_base, tmem_a0 + ks * 8, bd0 + (uint64_t)((ks * 512) >> 4), idesc, ks);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), parity);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                  "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 32; j++) Dg[(size_t)(warp * 32 + lane) * 64 + c0 + j] = (int32_t)v[j];
        }
        {   // what the copy left in TMEM: 16 columns of A
            uint32_t v[32];
            const uint32_t taddr = tmem_a0 + ((uint32_t)(warp * 32) << 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                  "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 16; j++) Adump[(warp * 32 + lane) * 16 + j] = v[j];
        }
    } else {          // rate of a pass
        long long t_total = 0;
        for (int rep = 0; rep < reps; rep++) {
            long long t0 = clock64();
            if (warp == 0 && elect_one()) {
                const uint64_t ad0 = make_desc(smem_u32(As), 2048, 128), bd0 = make_desc(smem_u32(Bs), 128, 32 * 128);
                int chunk = 0;
                for (int i = 0; i < S; i++)
                    for (int kc = 0; kc < 4; kc++, chunk++) {
                        const uint64_t ad = ad0 + (uint64_t)(((chunk % 6) * 8192) >> 4);
                        const uint32_t ta = tmem_a0 + (chunk & 3) * 16;
                        if (mode & 4) asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        utccp_128x256b(ta, ad);
                        utccp_128x256b(ta + 8, ad + (uint64_t)(4096 >> 4));
                        if (mode & 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar2)) : "memory");
                        for (int j = 0; j + i < S; j++) {
                            const uint64_t bd = bd0 + (uint64_t)((j * 64 * 256 + kc * 1024) >> 4);
                            const uint32_t d = tmem_base + (i + j) * 64;
                            mma_i8_ts(d, ta, bd, idesc, (i > 0 || kc > 0) ? 1u : 0u);
                            mma_i8_ts(d, ta + 8, bd + (uint64_t)(512 >> 4), idesc, 1u);
                        }
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
            }
            mbar_wait(smem_u32(&bar), parity);
            parity ^= 1;
            if (threadIdx.x == 0) t_total += clock64() - t0;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        if (threadIdx.x == 0) cycles[0] = t_total;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

// canonical no-swizzle layouts (bytes), int8
static int off_kmajor(int mn, int k, int lbo, int sbo) { return (mn % 8) * 16 + (mn / 8) * sbo + (k % 16) + (k / 16) * lbo; }
static int off_mnmajor(int mn, int k, int lbo, int sbo) { return (mn % 16) + (mn / 16) * sbo + (k % 8) * 16 + (k / 8) * lbo; }

static long long run(const std::vector<uint8_t>& A, const std::vector<uint8_t>& B, const std::vector<Prod>& prods, Params p, std::vector<int32_t>& D, long long* ld_cycles = nullptr) {
    uint8_t *dA, *dB;
    int32_t* dD;
    long long* dC;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, (size_t)128 * p.d_cols * 4));
    CK(cudaMalloc(&dC, 16));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, (size_t)128 * p.d_cols * 4));
    CK(cudaMemcpyToSymbol(c_prod, prods.data(), prods.size() * sizeof(Prod)));
    p.a_bytes = (int)A.size();
    p.b_bytes = (int)B.size();
    p.n_prod = (int)prods.size();
    const int smem = p.a_bytes + p.b_bytes;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_probe<<<1, 128, smem>>>(dA, dB, dD, dC, p);
    CK(cudaDeviceSynchronize());
    D.resize((size_t)128 * p.d_cols);
    long long c[2];
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    if (ld_cycles) *ld_cycles = c[1];
    return c[0];
}

int main() {
    srand(1);
    // ---------------- 1. descriptor conventions: one product D[128 x 64] = A[128 x K] B[64 x K]', K = 64 (two MMAs) ----------------
    const int M = 128, N = 64, K = 64;
    std::vector<int8_t> a(M * K), b(N * K);
    for (auto& v : a) v = (int8_t)(rand() % 129 - 64);
    for (auto& v : b) v = (int8_t)(rand() % 129 - 64);
    std::vector<int32_t> ref(M * N);
    for (int r = 0; r < M; r++)
        for (int n = 0; n < N; n++) {
            int s = 0;
            for (int k = 0; k < K; k++) s += (int)a[r * K + k] * (int)b[n * K + k];
            ref[r * N + n] = s;
        }
    // memory layouts as designed: A K-major (k16 columns 2048 B apart, 8-row groups 128 B apart), B MN-major (k8 groups 128 B apart, 16-column
    // groups K/8 * 128 B apart) or B K-major (k16 columns N/8*128 B apart, 8-column groups 128 B apart)
    const int A_LBO = 2048, A_SBO = 128;
    for (int bmaj = 1; bmaj >= 0; bmaj--) {
        const int B_LBO = bmaj ? 128 : (N / 8) * 128, B_SBO = bmaj ? (K / 8) * 128 : 128;
        std::vector<uint8_t> A((size_t)M * K), B((size_t)N * K);
        for (int r = 0; r < M; r++)
            for (int k = 0; k < K; k++) A[off_kmajor(r, k, A_LBO, A_SBO)] = (uint8_t)a[r * K + k];
        for (int n = 0; n < N; n++)
            for (int k = 0; k < K; k++) B[bmaj ? off_mnmajor(n, k, B_LBO, B_SBO) : off_kmajor(n, k, B_LBO, B_SBO)] = (uint8_t)b[n * K + k];
        for (int variant = 0; variant < 1; variant++) {  // the swapped-field variants address outside shared memory (illegal access): the designed roles are the right ones
            Params p{};
            p.a_lbo = (variant & 1) ? A_SBO : A_LBO;
            p.a_sbo = (variant & 1) ? A_LBO : A_SBO;
            p.b_lbo = (variant & 2) ? B_SBO : B_LBO;
            p.b_sbo = (variant & 2) ? B_LBO : B_SBO;
            p.b_major = bmaj;
            p.n = N;
            p.reps = 1;
            p.d_cols = N;
            std::vector<Prod> prods;
            for (int ks = 0; ks < K / 32; ks++) {
                // K-major: a 32-wide k-step is two 16-byte k-columns (LBO apart); MN-major: four 8-row k-groups (LBO apart)
                Prod q{ks * 2 * A_LBO, bmaj ? ks * 4 * B_LBO : ks * 2 * B_LBO, 0, ks > 0};
                prods.push_back(q);
            }
            std::vector<int32_t> D;
            run(A, B, prods, p, D);
            int bad = 0;
            for (int i = 0; i < M * N; i++) bad += D[i] != ref[i];
            printf("descriptor test: B %s-major, A fields %s, B fields %s: %d / %d mismatches  (D[0]=%d ref %d, D[1]=%d ref %d, D[64]=%d ref %d)\n", bmaj ? "MN" : "K",
                   (variant & 1) ? "swapped" : "as designed", (variant & 2) ? "swapped" : "as designed", bad, M * N, D[0], ref[0], D[1], ref[1], D[N], ref[N]);
        }
    }
    // ---------------- 2. rates ----------------
    rate<64, 7>(7);
    rate<64, 6>(6);
    rate<128, 7>(4);
    rate<256, 7>(2);
    rate<32, 7>(7);

    // ---------------- 3. A operand from TMEM (tcgen05.cp + .ts MMA) ----------------
    {
        const int A_LBO2 = 2048, A_SBO2 = 128;
        std::vector<uint8_t> A((size_t)M * K), B((size_t)N * K);
        for (int r = 0; r < M; r++)
            for (int k = 0; k < K; k++) A[off_kmajor(r, k, A_LBO2, A_SBO2)] = (uint8_t)a[r * K + k];
        for (int n = 0; n < N; n++)
            for (int k = 0; k < K; k++) B[off_mnmajor(n, k, 128, (K / 8) * 128)] = (uint8_t)b[n * K + k];
        uint8_t *dA, *dB; int32_t* dD; uint32_t* dAd; long long* dC;
        CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dAd, 128 * 16 * 4)); CK(cudaMalloc(&dC, 16));
        CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, M * N * 4));
        CK(cudaFuncSetAttribute(k_ts<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230912));
        k_ts<7><<<1, 128, A.size() + B.size()>>>(dA, dB, dD, dAd, dC, (int)A.size(), (int)B.size(), 0, 1);
        CK(cudaDeviceSynchronize());
        std::vector<int32_t> D(M * N); std::vector<uint32_t> Ad(128 * 16);
        CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(Ad.data(), dAd, 128 * 16 * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < M * N; i++) bad += D[i] != ref[i];
        printf("A from TMEM (tcgen05.cp.128x256b + .ts MMA): %d / %d mismatches (D[0]=%d ref %d)\n", bad, M * N, D[0], ref[0]);
        // how the copy laid A out: expect lane r, column c = bytes a[r][4c .. 4c+3] (k-step 0: columns 0..7, k-step 1: columns 8..15)
        int lay_bad = 0;
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < 16; c++) {
                uint32_t w = 0;
                for (int q = 0; q < 4; q++) w |= (uint32_t)(uint8_t)a[r * K + 4 * c + q] << (8 * q);
                lay_bad += Ad[r * 16 + c] != w;
            }
        printf("TMEM image of A: %d / 2048 words differ from [lane = row][column = k / 4]; lane 0: %08x %08x %08x %08x  expected %02x%02x%02x%02x ...; lane 1: %08x; lane 8: %08x\n", lay_bad, Ad[0], Ad[1],
               Ad[2], Ad[8], (uint8_t)a[3], (uint8_t)a[2], (uint8_t)a[1], (uint8_t)a[0], Ad[16], Ad[128]);
        // rate
        const int a_bytes = 6 * 8192, b_bytes = 7 * 64 * 256;
        std::vector<uint8_t> A2(a_bytes), B2(b_bytes);
        for (auto& v : A2) v = (uint8_t)(rand() % 129 - 64);
        for (auto& v : B2) v = (uint8_t)(rand() % 129 - 64);
        uint8_t *dA2, *dB2;
        CK(cudaMalloc(&dA2, a_bytes)); CK(cudaMalloc(&dB2, b_bytes));
        CK(cudaMemcpy(dA2, A2.data(), a_bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB2, B2.data(), b_bytes, cudaMemcpyHostToDevice));
        k_ts<7><<<1, 128, a_bytes + b_bytes>>>(dA2, dB2, dD, dAd, dC, a_bytes, b_bytes, 1, 8);
        CK(cudaDeviceSynchronize());
        long long c[2];
        CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
        printf("rate, A from TMEM: one pass (S = 7: 28 chunks = 56 copies + 224 MMAs of 128 x 64 x 32): %.0f cycles per pass, %.1f per MMA\n", (double)c[0] / 8, (double)c[0] / 8 / 224);
        {
            k_ts<7><<<1, 128, 230912>>>(dA2, dB2, dD, dAd, dC, a_bytes, b_bytes, 17, 8);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
            printf("  with the kernel's shared-memory map (ring at 176 KB, planes at 64 KB): %.0f cycles per pass\n", (double)c[0] / 8);
        }
        for (int grid : {16, 148}) {
            k_ts<7><<<grid, 128, a_bytes + b_bytes>>>(dA2, dB2, dD, dAd, dC, a_bytes, b_bytes, 1, 64);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
            printf("  the same pass on %d SMs at once (64 repetitions, cycles of whichever CTA wrote last): %.0f cycles per pass\n", grid, (double)c[0] / 64);
        }
        for (int nt : {256, 512}) {
            k_ts<7><<<1, nt, a_bytes + b_bytes>>>(dA2, dB2, dD, dAd, dC, a_bytes, b_bytes, 1, 8);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
            printf("  with %d threads in the CTA (all but one spin on the completion barrier): %.0f cycles per pass\n", nt, (double)c[0] / 8);
        }
        for (int mode : {3, 5, 7}) {
            k_ts<7><<<1, 128, a_bytes + b_bytes>>>(dA2, dB2, dD, dAd, dC, a_bytes, b_bytes, mode, 8);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
            printf("  + per chunk:%s%s: %.0f cycles per pass\n", (mode & 2) ? " tcgen05.commit" : "", (mode & 4) ? " tcgen05.fence::after_thread_sync" : "", (double)c[0] / 8);
        }
    }
    return 0;
}
