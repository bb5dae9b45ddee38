// Closed-loop harness in the shape of the reference's cpp/src/main.cpp:57-114: runMPC -> apply u0 -> simTimeStep.
//   usage: closed_loop <Params dir> <nn dir> [n_sim] [horizon]
#include "../include/mpcc/mpcc.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <Params dir> <nn dir> [n_sim] [horizon]\n", argv[0]); return 2; }
    const std::string P = argv[1], NN = argv[2];
    const int n_sim = argc > 3 ? std::atoi(argv[3]) : 50, N = argc > 4 ? std::atoi(argv[4]) : 10;
    const double Ts = 0.01;
    try {
        mpcc::PathToJson path{P + "/model.json", P + "/cost.json", P + "/bounds.json", P + "/track.json", P + "/normalization.json", P + "/sqp.json"};
        mpcc::MPC mpc(Ts, path, mpcc::ParamValue(), N);
        mpc.loadNetworks(NN + "/self_collision.f64", NN + "/env_collision.f64");
        mpcc::State x0{0, 0, 0, -M_PI / 2, 0, M_PI / 2, M_PI / 4, 0, 0};  // main.cpp:60-63
        mpcc::Input u0; u0.setZero();
        auto ee = mpc.robot_->getEEPosition({x0.q1, x0.q2, x0.q3, x0.q4, x0.q5, x0.q6, x0.q7});
        mpcc::ArcLengthSpline track; track.loadJson(path.track_path, &ee);  // track shifted to start at the EE (track.cpp:58-60)
        mpc.setTrack(track);
        std::printf("track length %.6f, EE (%.4f %.4f %.4f)\n", mpc.getTrackLength(), ee[0], ee[1], ee[2]);
        for (int i = 0; i < n_sim; i++) {
            mpcc::MPCReturn sol;
            bool ok = mpc.runMPC(sol, x0, u0);
            if (!ok) { std::printf("cycle %d: runMPC failed (status %d)\n", i, (int)sol.status); return 1; }
            u0 = sol.u0;
            x0 = mpc.batch().simTimeStep({x0}, {u0}, Ts)[0];
            if (i % 10 == 0 || i == n_sim - 1)
                std::printf("cycle %3d  s %.5f  vs %.4f  iters %d  total %.3f ms (set_qp %.3f solve_qp %.3f get_alpha %.3f)\n", i, x0.s, x0.vs, sol.sqp_iters,
                            1e3 * sol.compute_time.total, 1e3 * sol.compute_time.set_qp, 1e3 * sol.compute_time.solve_qp, 1e3 * sol.compute_time.get_alpha);
        }
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
