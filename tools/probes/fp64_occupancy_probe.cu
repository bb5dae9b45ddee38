// Probe: DFMA throughput of an 8x8 (or 8x4 / 4x4) register outer-product tile as a function of resident warps per SM
// sub-partition -- the k_mlp main loop without its loads.  Answers: can 2 warps per scheduler saturate the FP64 pipe?
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int C>
__global__ void k(double* out, int iters, const double* in) {
    double acc[R][C], w[R], x[C];
    for (int r = 0; r < R; r++) { w[r] = in[r] + threadIdx.x * 1e-9; for (int c = 0; c < C; c++) acc[r][c] = 0.0; }
    for (int c = 0; c < C; c++) x[c] = in[8 + c];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < C; c++) acc[r][c] = fma(w[r], x[c], acc[r][c]);
#pragma unroll
        for (int r = 0; r < R; r++) w[r] += 1e-12;   // keep the operands live and changing (R DADD per R*C DFMA)
    }
    double s = 0;
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) s += acc[r][c];
    if (s == 123.456) out[0] = s;
}

template <int R, int C>
void run(int sms, int threads, double* d, double* in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14;
    k<R, C><<<sms, threads>>>(d, 64, in);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<R, C><<<sms, threads>>>(d, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * R * C * (double)iters * threads * sms;
    printf("tile %dx%d  warps/SMSP %2d : %7.2f TFLOP/s\n", R, C, threads / 128, flops / best / 1e9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 128); cudaMemset(in, 0, 128);
    const int sms = p.multiProcessorCount;
    for (int t : {128, 256, 384, 512, 1024}) {
        if (t <= 256) run<8, 8>(sms, t, d, in);
        if (t <= 512) run<8, 4>(sms, t, d, in);
        run<4, 4>(sms, t, d, in);
    }
    return 0;
}
