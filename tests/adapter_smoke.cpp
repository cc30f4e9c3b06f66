// Compiles the C++ batched cSimAdapter of include/trl_b200.hpp against libtrl_b200.so and drives it from an arg file over a
// trivial rigid-body world (every character stands still). Without a GPU it checks the loud failure path.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "trl_b200.hpp"

struct StillWorld : trl::cWorldHook {
    int E = 0, n = 0, N = 0;
    void Reset(int num_envs) override { E = num_envs; }
    void GetState(std::vector<double>& pose, std::vector<double>& vel, std::vector<double>& body_pos,
                  std::vector<double>& body_vel) override {
        pose.assign((size_t)E * n, 0.0);
        vel.assign((size_t)E * n, 0.0);
        body_pos.assign((size_t)E * N * 3, 0.0);
        body_vel.assign((size_t)E * N * 3, 0.0);
        for (int e = 0; e < E; ++e) {
            pose[(size_t)e * n + 0] = 0.25 * e;
            pose[(size_t)e * n + 1] = 1.0;
            pose[(size_t)e * n + 3] = 1.0;
            for (int j = 0; j < N; ++j) {
                body_pos[((size_t)e * N + j) * 3 + 0] = 0.25 * e;
                body_pos[((size_t)e * N + j) * 3 + 1] = 1.0 - 0.2 * j;
            }
        }
    }
    void ApplyTau(const std::vector<double>& tau, double dt) override { (void)tau; (void)dt; ++applied; }
    void HasFallen(std::vector<unsigned char>& out) override { out.assign((size_t)E, 0); }
    int applied = 0;
};

int main(int argc, char** argv) {
    if (argc < 5) return 2;  // arg_file, relative path, n, N [, gpu]
    StillWorld world;
    world.n = std::atoi(argv[3]);
    world.N = std::atoi(argv[4]);
    const bool want_gpu = argc > 5 && std::strcmp(argv[5], "gpu") == 0;
    const int E = 6;
    trl::cBatchedSimAdapter sim({"train", "-arg_file=", argv[1], "-relative_file_path=", argv[2]}, E, &world);
    sim.setRandomSeed(11);
    if (!sim.init()) {
        std::printf("init failed: %s\n", trl::LastError().c_str());
        return want_gpu ? 4 : 0;  // expected on a CPU-only box: no CPU fallback
    }
    if (sim.getActionSpaceSize() != (size_t)world.n - 7) return 5;
    std::vector<double> action(E * sim.getActionSpaceSize(), 0.1);
    if (!sim.updateAction(action)) return 6;
    sim.handleUpdatedAction();
    if (!sim.update()) { std::printf("update failed: %s\n", trl::LastError().c_str()); return 7; }
    if (world.applied != 20) return 8;
    const std::vector<double> ob = sim.getState();
    if (ob.size() != E * sim.getObservationSpaceSize()) return 9;
    double s = 0;
    for (double v : ob) { if (!std::isfinite(v)) return 10; s += std::fabs(v); }
    for (double v : sim.tau()) if (!std::isfinite(v)) return 11;
    if (sim.agentHasFallen().size() != (size_t)E || !sim.needUpdatedAction()) return 12;
    std::printf("ok obs=%zu sum|ob|=%g\n", sim.getObservationSpaceSize(), s);
    sim.finish();
    return 0;
}
