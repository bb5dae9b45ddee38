// ArcLengthSpline::fitSpline for a chunk of tracks, one thread per track (dev_track_fit.cuh).
// This translation unit is compiled with --fmad=false: the fit differentiates nearly equal numbers (knot spacing ~ 0.02, so
// rounding is amplified by ~1 / h^2), and without fused multiply-adds the device executes the same rounded operations as the
// host fit (g++ on x86-64 emits none), which keeps the two tables within 1e-12 of each other.  Speed is irrelevant here.
#include "cycle_args.h"
#include "dev_track_fit.cuh"

namespace mpcc {

__global__ void k_fit_tracks(int n_tracks, int n, const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z,
                             const double* __restrict__ R, double* scratch, TrackTable* out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const size_t o = (size_t)t * n;
    tf_fit_track(n, TArr{(double*)X + o, 1}, TArr{(double*)Y + o, 1}, TArr{(double*)Z + o, 1}, TArr{(double*)R + 9 * o, 1}, TArr{scratch + t, (size_t)n_tracks}, out[t]);
}

void launch_fit_tracks(int n_tracks, int n, const double* X, const double* Y, const double* Z, const double* R, double* scratch, TrackTable* out, cudaStream_t s) {
    k_fit_tracks<<<(n_tracks + 63) / 64, 64, 0, s>>>(n_tracks, n, X, Y, Z, R, scratch, out);
}

}  // namespace mpcc
