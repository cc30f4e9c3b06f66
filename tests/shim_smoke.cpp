// Compiles and links the C++ shim (include/trl_b200.hpp) against libtrl_b200.so. Without a GPU it checks the loud
// failure path; with one (argv[2] == "gpu") it runs a tiny stable-PD call through the reference-shaped classes.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "trl_b200.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    trl::cCharModel model;
    if (!model.Init(argv[1])) {
        std::printf("model load failed: %s\n", trl::LastError().c_str());
        return 3;
    }
    const int n = model.GetNumDof(), N = model.GetNumJoints();
    std::printf("joints=%d dofs=%d\n", N, n);
    trl::cBatch batch;
    const int E = 8;
    const bool want_gpu = argc > 2 && std::strcmp(argv[2], "gpu") == 0;
    if (!batch.Init(model, E, 0)) {
        std::printf("no device: %s\n", trl::LastError().c_str());
        return want_gpu ? 4 : 0;  // expected on a CPU-only box: no CPU fallback
    }
    trl::cImpPDController pd;
    pd.Init(&batch);
    std::vector<double> pose((size_t)E * n, 0.0), vel((size_t)E * n, 0.0), tau;
    for (int e = 0; e < E; ++e) {
        pose[(size_t)e * n + 1] = 1.0;
        pose[(size_t)e * n + 3] = 1.0;  // identity root quaternion
    }
    if (!pd.UpdateControlForce(1.0 / 600, pose, vel, tau, true)) {
        std::printf("UpdateControlForce failed: %s\n", trl::LastError().c_str());
        return 5;
    }
    double s = 0;
    for (double t : tau) {
        if (!std::isfinite(t)) return 6;
        s += std::fabs(t);
    }
    const std::vector<double>& M = pd.GetMassMat();
    for (int i = 0; i < n; ++i)
        if (M[(size_t)i * n + 6] != 0.0) return 7;  // dead root slot
    std::printf("ok sum|tau|=%g\n", s);
    return 0;
}
