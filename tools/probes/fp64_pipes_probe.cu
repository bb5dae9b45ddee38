// Probe: FP64 throughput of the FMA pipe (DFMA), the tensor pipe (mma.sync.m8n8k4.f64 = DMMA) and both mixed.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: DFMA only, 1: DMMA only, 2: mixed 1:1 instruction ratio
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double f[8], c[16];
    for (int i = 0; i < 8; i++) f[i] = threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 16; i++) c[i] = i * 1e-3;
    const double m = 1.0000001, a = 1e-9, x = 0.5 + threadIdx.x * 1e-6, y = 0.25;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = fma(f[i], m, a);
            }
            if (MODE == 1 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) dmma(c[2 * i], c[2 * i + 1], x, y);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += f[i];
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int sms) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8, iters = 2048;
    k<MODE><<<grid, 256>>>(d, 16);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // per thread per it: 8 u x (8 DFMA = 8 FMA) ; DMMA: 8 u x 8 mma, each mma = 256 MAC per warp = 8 MAC per lane
    double fma_flops = (MODE == 0 || MODE == 2) ? 2.0 * 64 * iters * 256.0 * grid : 0;
    double mma_flops = (MODE == 1 || MODE == 2) ? 2.0 * 64 * 8 * iters * 256.0 * grid : 0;
    printf("%-12s %.3f ms  DFMA %.2f TFLOP/s  DMMA %.2f TFLOP/s  total %.2f TFLOP/s\n", name, best, fma_flops / best / 1e9, mma_flops / best / 1e9, (fma_flops + mma_flops) / best / 1e9);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("DFMA", p.multiProcessorCount);
    run<1>("DMMA", p.multiProcessorCount);
    run<2>("DFMA+DMMA", p.multiProcessorCount);
    return 0;
}
