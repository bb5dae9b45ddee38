// Probe: the k_mlp main loop re-expressed on the FP64 tensor path (mma.sync.m8n8k4.f64), operands from shared memory in
// fragment order.  Per warp: 32 rows x 64 columns (4 m-blocks x 8 n-blocks), per 4 k-steps 2 + 4 LDS.128 for 32 DMMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void chunk_mma(const double2* __restrict__ Wc, const double2* __restrict__ Xs, int kb0, int warp, int lane, double (&acc)[4][8][2]) {
#pragma unroll
    for (int kb = 0; kb < 4; kb++) {
        const double2 a01 = Wc[((kb * 8 + warp) * 2 + 0) * 32 + lane], a23 = Wc[((kb * 8 + warp) * 2 + 1) * 32 + lane];
        double2 b[4];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = Xs[((kb0 + kb) * 4 + j) * 32 + lane];
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                dmma(acc[mb][2 * j][0], acc[mb][2 * j][1], a[mb], b[j].x);
                dmma(acc[mb][2 * j + 1][0], acc[mb][2 * j + 1][1], a[mb], b[j].y);
            }
    }
}

template <int MODE>  // 0: no barrier, 1: barrier per chunk
__global__ void __launch_bounds__(256, 1) k(double* out, int layers) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* Xs = reinterpret_cast<double2*>(smem_raw);
    double2* Wbuf = reinterpret_cast<double2*>(smem_raw + 256 * 64 * sizeof(double));
    for (int i = threadIdx.x; i < 256 * 32 + 2 * 2048; i += 256) Xs[i] = make_double2(1e-3 * (i & 7), 1e-4);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[4][8][2];
    for (int m = 0; m < 4; m++) for (int c = 0; c < 8; c++) acc[m][c][0] = acc[m][c][1] = 0;
    int buf = 0;
    for (int l = 0; l < layers; l++)
        for (int ch = 0; ch < 16; ch++) {
            if (MODE == 1) __syncthreads();
            chunk_mma(Wbuf + buf * 2048, Xs, ch * 4, warp, lane, acc);
            buf ^= 1;
        }
    double s = 0;
    for (int m = 0; m < 4; m++) for (int c = 0; c < 8; c++) s += acc[m][c][0] + acc[m][c][1];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int sms, double* d) {
    const int smem = (256 * 64 + 2 * 4096) * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int layers = 2000;
    k<MODE><<<sms, 256, smem>>>(d, 4);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<MODE><<<sms, 256, smem>>>(d, layers); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 64 * 256.0 * layers * 256 * sms;
    printf("%-28s %7.2f TFLOP/s  (%s)\n", name, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, 8);
    run<0>("DMMA loop, no barrier", p.multiProcessorCount, d);
    run<1>("DMMA loop, barrier/chunk", p.multiProcessorCount, d);
    return 0;
}
