// Probe: the k_mlp main loop (mlp_chunk) alone on a resident shared-memory tile -- no weight streaming, optional barriers.
// Separates the cost of the LDS operand traffic and of the per-chunk barrier from the DFMA roof.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mpcc_manipulator_b200/csrc/mlp_kernel.cuh"
using namespace mpcc;

template <int UNROLL>
__device__ __forceinline__ void chunk_u(const double2* __restrict__ Wb, int wofs, const double2* __restrict__ Xs, int k0, int tx, double (&acc)[8][8]) {
#pragma unroll UNROLL
    for (int kk = 0; kk < 16; kk++) {
        double2 w2[4], x2[4];
#pragma unroll
        for (int i = 0; i < 4; i++) w2[i] = Wb[kk * 128 + wofs + i * 32];
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

template <int MODE>  // 0: no barrier, unroll 4;  1: barrier per chunk;  2: unroll 8;  3: unroll 2; 4: unroll 16 (full)
__global__ void __launch_bounds__(256, 1) k(double* out, int layers) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* Xs = reinterpret_cast<double2*>(smem_raw);
    double2* Wbuf = reinterpret_cast<double2*>(smem_raw + 256 * 64 * sizeof(double));
    for (int i = threadIdx.x; i < 256 * 32 + 2 * 2048; i += 256) Xs[i] = make_double2(1e-3 * (i & 7), 1e-4);
    __syncthreads();
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    double acc[8][8];
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) acc[r][c] = 0;
    int buf = 0;
    for (int l = 0; l < layers; l++)
        for (int ch = 0; ch < 16; ch++) {
            if (MODE == 1) __syncthreads();
            if (MODE == 0 || MODE == 1) chunk_u<4>(Wbuf + buf * 2048, ty, Xs, ch * 16, tx, acc);
            if (MODE == 2) chunk_u<8>(Wbuf + buf * 2048, ty, Xs, ch * 16, tx, acc);
            if (MODE == 3) chunk_u<2>(Wbuf + buf * 2048, ty, Xs, ch * 16, tx, acc);
            if (MODE == 4) chunk_u<16>(Wbuf + buf * 2048, ty, Xs, ch * 16, tx, acc);
            buf ^= 1;
        }
    double s = 0;
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) s += acc[r][c];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int sms, double* d) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int layers = 2000;
    k<MODE><<<sms, 256, MLP_SMEM_BYTES>>>(d, 4);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<MODE><<<sms, 256, MLP_SMEM_BYTES>>>(d, layers); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 64 * 256.0 * layers * 256 * sms;
    printf("%-28s %7.2f TFLOP/s  (%s)\n", name, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, 8);
    run<0>("unroll 4, no barrier", p.multiProcessorCount, d);
    run<1>("unroll 4, barrier/chunk", p.multiProcessorCount, d);
    run<2>("unroll 8, no barrier", p.multiProcessorCount, d);
    run<3>("unroll 2, no barrier", p.multiProcessorCount, d);
    run<4>("unroll 16, no barrier", p.multiProcessorCount, d);
    return 0;
}
