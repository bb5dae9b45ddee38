// Does the product's own issue function (oz_issue_pass of csrc/mlp_oz_kernel.cuh) run a pass as fast as the bare loop of oz_umma_probe.cu (9.7 k cycles)?
// One CTA, the kernel's shared-memory map, operands resident (no weight stream), nothing else on the SM.
#include "../../mpcc_manipulator_b200/csrc/mlp_oz_kernel.cuh"
#include <cstdio>
using namespace mpcc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256, 1) k_issue(long long* out, int reps, int nthreads_spin) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* misc = smem_raw + OZ_OFF_MISC;
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc + OZ_MISC_BARS);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + OZ_MISC_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_full = oz_smem_u32(bars), bar_empty = oz_smem_u32(bars + OZ_NSLOT), bar_group = oz_smem_u32(bars + 2 * OZ_NSLOT), bar_tfree = oz_smem_u32(bars + 2 * OZ_NSLOT + OZ_S);
    for (int i = tid * 16; i < OZ_OFF_MISC; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem_raw + i) = make_uint4(0x01020304u * (i & 3), 0x3f013f01u, 0x00010203u, 0x40c040c0u);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (tid == 0) {
        for (int i = 0; i < OZ_NBARS; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_full + 8 * i), "r"(i < OZ_NSLOT ? 32 : (i == OZ_NBARS - 1 ? 128 : 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(oz_smem_u32(tmem_ptr_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_ptr_s;
    const uint32_t ring_addr = oz_smem_u32(smem_raw + OZ_OFF_RING), planes_addr = oz_smem_u32(smem_raw + OZ_OFF_PLANES);
    long long total = 0;
    for (int r = 0; r < reps; r++) {
        const long long t0 = clock64();
        if (warp == 0) {
            oz_issue_pass(tmem, ring_addr, planes_addr, bar_full, bar_empty, bar_group, bar_tfree, (uint32_t)(r * OZ_CHUNKS_PER_PASS), 0u, true, true);
            oz_mbar_wait(bar_group + 8 * (OZ_S - 1), r & 1u);
        }
        if (tid == 0) total += clock64() - t0;
        __syncthreads();
    }
    if (tid == 0) out[0] = total;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM_BYTES));
    for (int nt : {32, 256}) {
        k_issue<<<1, nt, OZ_SMEM_BYTES>>>(d, 8, 0);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
        printf("oz_issue_pass, %d threads in the CTA: %.0f cycles per pass\n", nt, (double)c / 8);
    }
    return 0;
}
