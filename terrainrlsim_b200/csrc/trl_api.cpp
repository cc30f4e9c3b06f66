// C ABI: batch lifetime, pointer-mode staging and kernel sequencing. No torch types, no CPU fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "trl_internal.h"

int trl_obs_state_size(int N, const trl_obs_cfg& cfg);
double trl_fp64_probe_tflops(void* stream);

namespace {

bool cuda_ok(cudaError_t err, const char* what) {
    if (err == cudaSuccess) return true;
    trl_set_error(std::string(what) + ": " + cudaGetErrorString(err));
    return false;
}

// grow-only device buffer used for HOST pointer mode staging
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (!cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc(staging)")) return false;
        cap = bytes;
        return true;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct trl_batch {
    const trl_model* model = nullptr;
    int E = 0;
    int device = 0;
    int ptr_mode = TRL_PTR_HOST;
    trl_obs_cfg cfg{};
    int F = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // fork/join plumbing of trl_step_fused: the three independent kernel chains of a step (stable PD | kin-char + FK |
    // observation) run concurrently on the main stream and two side streams (captured as parallel graph branches)
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};  // kin chain, observation chain, stable-PD chain (highest priority)
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr}, ev_pv = nullptr;
    cudaStream_t obs_side = nullptr;               // k_obs_feat beside the ground-sample kernel
    cudaEvent_t ev_obs[2] = {nullptr, nullptr};
    // second lane of chain streams: the HOST-pointer step is pipelined over env slices, alternating between the two
    // lanes so one slice's uploads overlap the previous slice's kernels and downloads
    cudaStream_t side2[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_join2[3] = {nullptr, nullptr, nullptr}, ev_pv2 = nullptr;
    int host_slices = 2;  // measured: 2 slices best (13.2 M env-steps/s); 4 and 8 lose to per-copy overheads
    static constexpr int kMaxSlices = 16;
    cudaEvent_t ev_up[kMaxSlices][3] = {}, ev_cmp[kMaxSlices][3] = {};  // per slice: inputs uploaded (per upload stream) / chain finished
    cudaStream_t up_streams[3] = {nullptr, nullptr, nullptr};  // one H2D stream moves ~20 GB/s on this platform, three reach PCIe speed
    int overlap = 1;
    int host_async = 0;  // HOST-pointer trl_step_fused returns after enqueueing; trl_batch_sync() completes it
    DevModel* d_model = nullptr;
    double* d_frames = nullptr;
    int* d_status = nullptr;
    double* d_sample_local = nullptr;  // [G][2] local (x, z) of every ground sample
    double sample_diag = 0.0;          // diagonal of the local bounding box of the ground samples (bounds the staged cell rectangle)
    int sample_table_ok = 0;           // every local sample coordinate is 0 or in [2^-300, 2^300]
    void* d_env_rec = nullptr;         // [E] per-env sampling transform handed from k_obs_env to k_obs_samples
    double* d_pd_scratch = nullptr;    // [E][stride] assembled (M + dt Kd, rhs, tau terms) between the two PD kernels
    // ground
    DevGround ground{};
    float* d_ground_data = nullptr;
    DevPiece* d_pieces = nullptr;
    int* d_env_to_field = nullptr;
    // procedural terrain (trl_ground_generate): generator config and the per-terrain state on the device
    TerrainGenCfg tgen{};
    TerrainState* d_tstate = nullptr;
    int* d_tchanged = nullptr;
    double* d_tbounds = nullptr;  // [2][F][3] staging of the view bounds
    long long launches = 0;
    // trl_step_fused_ex: internal device-resident intermediates
    double* d_zero_nvec = nullptr;  // [E][n] zeros: tar_vel == NULL (the reference's default target velocity)
    DevBuf kin_tmp;                 // [E][2n + 3N] kin pose / vel / body positions the caller did not ask for (reward epilogue)
    DevBuf state_tmp;               // [E][F] double state behind an fp32-only state output
    // HOST-pointer step as a CUDA graph: a caller that steps in a loop passes the same host arrays every time, and the
    // sliced pipeline is ~40 copies / launches / event calls per slice -- the host thread, not PCIe, then limits how fine
    // the pipeline can be cut. The second call with identical arguments is captured (all copies, kernels and
    // dependencies of the call) and replayed from then on with one launch; any change of an argument, of the staging
    // buffers or of the batch's device state drops the graph.
    struct HostGraph {
        cudaGraphExec_t exec = nullptr;
        trl_step_args a{};
        trl_step_opts o{};
        bool has_o = false;
        int seen = 0;                   // calls with these arguments so far (captured on the second one)
        unsigned long long tick = 0;    // last use (the least recently used entry makes room)
        long long launches = 0;
        size_t pool_used = 0;           // staging buffers the captured call was handed (in pool order)
        std::vector<const void*> snap;  // every device address / setting baked into the graph
    };
    static constexpr int kHostGraphs = 4;  // argument sets kept (e.g. the substeps with and without the policy state)
    std::vector<HostGraph> hgs;
    bool hg_failed = false;
    unsigned long long hg_tick = 0;
    cudaEvent_t ev_bv = nullptr;    // body states staged (reward epilogue on the kin stream)
    // staging (host pointer mode): a small pool of grow-only device buffers handed out per call
    std::vector<DevBuf> pool;
    size_t pool_next = 0;

    void* stage_in(const void* host, size_t bytes, cudaStream_t st = nullptr) {
        if (!host) return nullptr;
        if (pool_next >= pool.size()) pool.emplace_back();
        DevBuf& b = pool[pool_next++];
        if (!b.reserve(bytes)) return nullptr;
        if (!cuda_ok(cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, st ? st : stream), "H2D")) return nullptr;
        return b.p;
    }
    void* stage_out(size_t bytes) {
        if (pool_next >= pool.size()) pool.emplace_back();
        DevBuf& b = pool[pool_next++];
        if (!b.reserve(bytes)) return nullptr;
        return b.p;
    }
};

namespace {

struct Guard {  // selects the batch's device for the duration of a call
    int prev = -1;
    explicit Guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~Guard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// Resolves an input array according to the pointer mode. Returns false on error.
template <typename T>
bool in_ptr(trl_batch* b, const T* user, size_t count, const T** out, cudaStream_t st = nullptr) {
    if (!user || count == 0) { *out = nullptr; return true; }
    if (b->ptr_mode == TRL_PTR_DEVICE) { *out = user; return true; }
    *out = static_cast<const T*>(b->stage_in(user, count * sizeof(T), st));
    return *out != nullptr;
}

// Output array: device pointer to write to (+ remembers the host destination and the stream for the copy back).
struct OutCopy { void* host; void* dev; size_t bytes; cudaStream_t st; };

template <typename T>
bool out_ptr(trl_batch* b, T* user, size_t count, T** out, std::vector<OutCopy>& copies, bool preload = false,
             cudaStream_t st = nullptr) {
    if (!user || count == 0) { *out = nullptr; return true; }
    if (b->ptr_mode == TRL_PTR_DEVICE) { *out = user; return true; }
    void* d = preload ? b->stage_in(user, count * sizeof(T), st) : b->stage_out(count * sizeof(T));
    if (!d) return false;
    *out = static_cast<T*>(d);
    copies.push_back({user, d, count * sizeof(T), st});
    return true;
}

int finish(trl_batch* b, std::vector<OutCopy>& copies, int launch_rc) {
    if (launch_rc != 0) {
        trl_set_error(std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)launch_rc));
        return TRL_ERR_CUDA;
    }
    if (b->ptr_mode == TRL_PTR_HOST) {
        for (auto& c : copies)
            if (!cuda_ok(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, c.st ? c.st : b->stream), "D2H")) return TRL_ERR_CUDA;
        if (!cuda_ok(cudaStreamSynchronize(b->stream), "cudaStreamSynchronize")) return TRL_ERR_CUDA;
    }
    return TRL_OK;
}

}  // namespace

extern "C" trl_batch* trl_batch_create(const trl_model* m, int num_envs, int device, const trl_obs_cfg* cfg) {
    if (!m || num_envs < 1 || device < 0) {
        trl_set_error("trl_batch_create: invalid argument (device must be >= 0: there is no CPU fallback)");
        return nullptr;
    }
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count <= device) {
        trl_set_error(std::string("trl_batch_create: no usable CUDA device ") + std::to_string(device) + " (" +
                      (err != cudaSuccess ? cudaGetErrorString(err) : "device ordinal out of range") +
                      "); this library has no CPU fallback");
        cudaGetLastError();
        return nullptr;
    }
    std::unique_ptr<trl_batch> b(new trl_batch());
    b->model = m;
    b->E = num_envs;
    b->device = device;
    if (cfg) b->cfg = *cfg;
    else { b->cfg = trl_obs_cfg{TRL_OBS_TERRAIN_RL, TRL_SAMPLER_LINE, 200, 0, -0.5, 10.0, -1}; }
    if (b->cfg.num_samples < 0 || (b->cfg.sampler == TRL_SAMPLER_GRID && b->cfg.grid_res * b->cfg.grid_res != b->cfg.num_samples)) {
        trl_set_error("trl_batch_create: inconsistent observation config");
        return nullptr;
    }
    b->F = trl_obs_state_size(m->N, b->cfg);
    Guard g(device);
    if (!cuda_ok(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return nullptr;
    b->stream = b->own_stream;
    // Block-scheduling priorities: the stable-PD chain is the longest dependent chain of the fused step (three
    // kernels back to back), so its kernels must not queue behind the thousands of blocks of the observation kernels.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = higher priority
    for (int i = 0; i < 3; ++i) {
        // side[0] = kin chain (middle), side[1] = observation chain and side[2] = stable-PD chain (highest): the measured
        // best of the orderings tried (profiles/README.md: round 1, and again in round 2 after the observation chain's
        // first kernel became short -- kin / obs / PD / body-features as lo-mid-hi-mid 0.0890 ms, mid-hi-hi-mid 0.0870,
        // hi-hi-hi 0.0960, mid-hi-mid 0.0954)
        const int prio = (i == 0) ? (prio_lo + prio_hi) / 2 : prio_hi;
        if (!cuda_ok(cudaStreamCreateWithPriority(&b->side[i], cudaStreamNonBlocking, prio), "cudaStreamCreate")) return nullptr;
        if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_join[i], cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
        if (!cuda_ok(cudaStreamCreateWithPriority(&b->side2[i], cudaStreamNonBlocking, prio), "cudaStreamCreate")) return nullptr;
        if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_join2[i], cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    }
    if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_pv2, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    if (const char* hs = getenv("TRL_HOST_SLICES")) b->host_slices = std::min((int)trl_batch::kMaxSlices, std::max(1, atoi(hs)));  // tuning hook
    for (int i = 0; i < trl_batch::kMaxSlices; ++i) {
        for (int k = 0; k < 3; ++k) {
            if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_up[i][k], cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
            if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_cmp[i][k], cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
        }
    }
    for (int k = 0; k < 3; ++k)
        if (!cuda_ok(cudaStreamCreateWithFlags(&b->up_streams[k], cudaStreamNonBlocking), "cudaStreamCreate")) return nullptr;
    if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    {
        const char* fsd = getenv("TRL_OBS_FEAT_SIDE");  // A/B hook: 0 = body features before the ground samples on one stream
        if (!fsd || atoi(fsd) != 0) {
            // below the observation chain's own priority: at the same (or a higher) one the ground-sample kernel is held
            // back behind feature blocks that wait for registers while the stable-PD tiles fill the SMs
            const int obs_prio = (prio_lo + prio_hi) / 2;
            if (!cuda_ok(cudaStreamCreateWithPriority(&b->obs_side, cudaStreamNonBlocking, obs_prio), "cudaStreamCreate")) return nullptr;
            for (int i = 0; i < 2; ++i)
                if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_obs[i], cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
        }
    }
    if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_pv, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_bv, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    {
        const char* ov = getenv("TRL_STEP_OVERLAP");  // tuning hook: 0 serialises the chains on one stream
        if (ov) b->overlap = atoi(ov) != 0;
    }
    if (!cuda_ok(cudaMalloc((void**)&b->d_model, sizeof(DevModel)), "cudaMalloc(model)")) return nullptr;
    if (!cuda_ok(cudaMemcpy(b->d_model, &m->dev, sizeof(DevModel), cudaMemcpyHostToDevice), "H2D(model)")) return nullptr;
    if (m->num_frames > 0) {
        size_t bytes = m->frames.size() * sizeof(double);
        if (!cuda_ok(cudaMalloc((void**)&b->d_frames, bytes), "cudaMalloc(frames)")) return nullptr;
        if (!cuda_ok(cudaMemcpy(b->d_frames, m->frames.data(), bytes, cudaMemcpyHostToDevice), "H2D(frames)")) return nullptr;
    }
    if (b->cfg.num_samples > 0) {
        // CalcGroundSamplePos in the sampling frame, same expression order as the reference
        // (TerrainRLCharController.cpp:310-319; CtController.cpp:189-211 / BipedController3D.cpp:1932-1949)
        const int G = b->cfg.num_samples;
        std::vector<double> loc((size_t)G * 2 + 4);  // + [min x, max x, min z, max z] of the local sample positions
        for (int s = 0; s < G; ++s) {
            if (b->cfg.sampler == TRL_SAMPLER_LINE) {
                loc[2 * s] = ((b->cfg.view_max - b->cfg.view_min) * s) / (G - 1) + b->cfg.view_min;
                loc[2 * s + 1] = 0.0;
            } else {
                const int res = b->cfg.grid_res;
                const double ox = 0.5 * (b->cfg.view_max + b->cfg.view_min);
                const double w = b->cfg.view_max - b->cfg.view_min;
                const double u = static_cast<double>(s % res) / (res - 1);
                const double v = static_cast<double>(s / res) / (res - 1);
                loc[2 * s] = ox + (u - 0.5) * w;
                loc[2 * s + 1] = 0.0 + (v - 0.5) * w;
            }
        }
        loc[2 * G] = loc[2 * G + 1] = loc[0];
        loc[2 * G + 2] = loc[2 * G + 3] = loc[1];
        for (int s = 0; s < G; ++s) {
            loc[2 * G] = std::min(loc[2 * G], loc[2 * s]);
            loc[2 * G + 1] = std::max(loc[2 * G + 1], loc[2 * s]);
            loc[2 * G + 2] = std::min(loc[2 * G + 2], loc[2 * s + 1]);
            loc[2 * G + 3] = std::max(loc[2 * G + 3], loc[2 * s + 1]);
        }
        b->sample_diag = std::hypot(loc[2 * G + 1] - loc[2 * G], loc[2 * G + 3] - loc[2 * G + 2]);
        b->sample_table_ok = 1;
        for (int s = 0; s < 2 * G; ++s) {
            const double a = std::fabs(loc[s]);
            if (!(a == 0.0 || (a >= 0x1p-300 && a <= 0x1p300))) b->sample_table_ok = 0;
        }
        if (!cuda_ok(cudaMalloc((void**)&b->d_sample_local, sizeof(double) * loc.size()), "cudaMalloc(samples)")) return nullptr;
        if (!cuda_ok(cudaMemcpy(b->d_sample_local, loc.data(), sizeof(double) * loc.size(), cudaMemcpyHostToDevice), "H2D(samples)")) return nullptr;
    }
    {
        const size_t tile = (size_t)trl_pd_scratch_tile();
        const size_t bytes = trl_pd_scratch_doubles(m->dev) * sizeof(double) * (((size_t)num_envs + tile - 1) / tile * tile);
        if (!cuda_ok(cudaMalloc((void**)&b->d_pd_scratch, bytes), "cudaMalloc(pd scratch)")) return nullptr;
        cudaMemset(b->d_pd_scratch, 0, bytes);  // structurally-zero entries of the packed matrix are never written
    }
    if (!cuda_ok(cudaMalloc(&b->d_env_rec, trl_obs_env_rec_bytes() * (size_t)num_envs), "cudaMalloc(env_rec)")) return nullptr;
    {
        const size_t bytes = sizeof(double) * (size_t)num_envs * m->n;
        if (!cuda_ok(cudaMalloc((void**)&b->d_zero_nvec, bytes), "cudaMalloc(zeros)")) return nullptr;
        cudaMemset(b->d_zero_nvec, 0, bytes);
    }
    if (!cuda_ok(cudaMalloc((void**)&b->d_status, sizeof(int) * (size_t)num_envs), "cudaMalloc(status)")) return nullptr;
    cudaMemset(b->d_status, 0, sizeof(int) * (size_t)num_envs);
    std::memset(&b->ground, 0, sizeof(b->ground));
    return b.release();
}

extern "C" void trl_batch_destroy(trl_batch* b) {
    if (!b) return;
    Guard g(b->device);
    cudaStreamSynchronize(b->stream);
    for (auto& p : b->pool) p.release();
    if (b->d_model) cudaFree(b->d_model);
    if (b->d_frames) cudaFree(b->d_frames);
    if (b->d_status) cudaFree(b->d_status);
    if (b->d_sample_local) cudaFree(b->d_sample_local);
    if (b->d_env_rec) cudaFree(b->d_env_rec);
    if (b->d_tstate) cudaFree(b->d_tstate);
    if (b->d_tchanged) cudaFree(b->d_tchanged);
    if (b->d_tbounds) cudaFree(b->d_tbounds);
    if (b->d_pd_scratch) cudaFree(b->d_pd_scratch);
    if (b->d_ground_data) cudaFree(b->d_ground_data);
    if (b->d_pieces) cudaFree(b->d_pieces);
    if (b->d_env_to_field) cudaFree(b->d_env_to_field);
    for (int i = 0; i < 3; ++i) {
        if (b->side[i]) { cudaStreamSynchronize(b->side[i]); cudaStreamDestroy(b->side[i]); }
        if (b->ev_join[i]) cudaEventDestroy(b->ev_join[i]);
        if (b->side2[i]) { cudaStreamSynchronize(b->side2[i]); cudaStreamDestroy(b->side2[i]); }
        if (b->ev_join2[i]) cudaEventDestroy(b->ev_join2[i]);
    }
    if (b->ev_pv2) cudaEventDestroy(b->ev_pv2);
    for (int i = 0; i < trl_batch::kMaxSlices; ++i) {
        for (int k = 0; k < 3; ++k) {
            if (b->ev_up[i][k]) cudaEventDestroy(b->ev_up[i][k]);
            if (b->ev_cmp[i][k]) cudaEventDestroy(b->ev_cmp[i][k]);
        }
    }
    for (int k = 0; k < 3; ++k)
        if (b->up_streams[k]) { cudaStreamSynchronize(b->up_streams[k]); cudaStreamDestroy(b->up_streams[k]); }
    for (auto& g : b->hgs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    for (int i = 0; i < 2; ++i) if (b->ev_obs[i]) cudaEventDestroy(b->ev_obs[i]);
    if (b->obs_side) cudaStreamDestroy(b->obs_side);
    if (b->ev_pv) cudaEventDestroy(b->ev_pv);
    if (b->ev_bv) cudaEventDestroy(b->ev_bv);
    if (b->d_zero_nvec) cudaFree(b->d_zero_nvec);
    b->kin_tmp.release();
    b->state_tmp.release();
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    delete b;
}

extern "C" int trl_batch_set_pointer_mode(trl_batch* b, int mode) {
    if (!b || (mode != TRL_PTR_HOST && mode != TRL_PTR_DEVICE)) return TRL_ERR_INVALID_ARG;
    b->ptr_mode = mode;
    return TRL_OK;
}

extern "C" int trl_batch_set_stream(trl_batch* b, void* cuda_stream) {
    if (!b) return TRL_ERR_INVALID_ARG;
    b->stream = (cuda_stream == TRL_STREAM_OWN) ? b->own_stream : (cudaStream_t)cuda_stream;
    return TRL_OK;
}

extern "C" int trl_batch_sync(trl_batch* b) {
    if (!b) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    return cuda_ok(cudaStreamSynchronize(b->stream), "cudaStreamSynchronize") ? TRL_OK : TRL_ERR_CUDA;
}

extern "C" int trl_poli_state_size(const trl_batch* b) { return b ? b->F : TRL_ERR_INVALID_ARG; }
extern "C" int trl_batch_num_envs(const trl_batch* b) { return b ? b->E : TRL_ERR_INVALID_ARG; }
extern "C" long long trl_batch_launch_count(const trl_batch* b) { return b ? b->launches : 0; }
extern "C" int trl_debug_trace_enable(int on) { return trl_trace_enable(on) == 0 ? TRL_OK : TRL_ERR_CUDA; }
extern "C" int trl_debug_trace_read(unsigned long long* out32) { return trl_trace_read(out32) == 0 ? TRL_OK : TRL_ERR_INVALID_ARG; }

// ------------------------------------------------------------------------------------------------
extern "C" int trl_imp_pd_update_control_force(trl_batch* b, double dt, const double* pose, const double* vel,
                                               const double* tar_pose, const double* tar_vel,
                                               const unsigned char* joint_active, const double* kp, const double* kd,
                                               double* inout_tau, double* out_mass, double* out_bias) {
    if (!b || !pose || !vel || !tar_pose || !tar_vel || !inout_tau) {
        trl_set_error("trl_imp_pd_update_control_force: null argument");
        return TRL_ERR_INVALID_ARG;
    }
    if (!(dt > 0)) return TRL_OK;  // UpdateControlForce does nothing for time_step <= 0 (ImpPDController.cpp:56)
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    StepPtrs p{};
    p.dt = dt;
    p.tau_accumulate = 1;
    p.status = b->d_status;
    if (!in_ptr(b, pose, E * n, &p.pose) || !in_ptr(b, vel, E * n, &p.vel) || !in_ptr(b, tar_pose, E * n, &p.tar_pose) ||
        !in_ptr(b, tar_vel, E * n, &p.tar_vel) || !in_ptr(b, joint_active, E * N, &p.joint_active) ||
        !in_ptr(b, kp, E * N, &p.kp) || !in_ptr(b, kd, E * N, &p.kd) ||
        !out_ptr(b, inout_tau, E * n, &p.tau, copies, true) || !out_ptr(b, out_mass, E * n * n, &p.out_mass, copies) ||
        !out_ptr(b, out_bias, E * n, &p.out_bias, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_imp_pd(b->d_model, b->model->dev, b->E, p, b->d_pd_scratch, b->stream);
    if (rc > 0) { b->launches += rc; rc = 0; } else { rc = -rc; }
    return finish(b, copies, rc);
}

extern "C" int trl_apply_control_forces(trl_batch* b, const double* pose, const double* tau, double* out_tau,
                                        double* out_body_wrench) {
    if (!b || !pose || !tau) { trl_set_error("trl_apply_control_forces: null argument"); return TRL_ERR_INVALID_ARG; }
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double *d_pose = nullptr, *d_tau = nullptr;
    double *d_ot = nullptr, *d_w = nullptr;
    if (!in_ptr(b, pose, E * n, &d_pose) || !in_ptr(b, tau, E * n, &d_tau) || !out_ptr(b, out_tau, E * n, &d_ot, copies) ||
        !out_ptr(b, out_body_wrench, E * N * 6, &d_w, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_apply_tau(b->d_model, b->model->dev, b->E, d_pose, d_tau, d_ot, d_w, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

extern "C" int trl_exp_pd_update_control_force(trl_batch* b, double dt, const double* pose, const double* vel,
                                               const double* tar_pose, const double* tar_vel,
                                               const unsigned char* joint_active, const double* kp, const double* kd,
                                               double* inout_tau) {
    if (!b || !pose || !vel || !tar_pose || !inout_tau) {
        trl_set_error("trl_exp_pd_update_control_force: null argument");
        return TRL_ERR_INVALID_ARG;
    }
    if (!(dt > 0)) return TRL_OK;  // cExpPDController::UpdateControlForce, ExpPDController.cpp:62
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    StepPtrs p{};
    p.dt = dt;
    if (!in_ptr(b, pose, E * n, &p.pose) || !in_ptr(b, vel, E * n, &p.vel) || !in_ptr(b, tar_pose, E * n, &p.tar_pose) ||
        !in_ptr(b, tar_vel, E * n, &p.tar_vel) || !in_ptr(b, joint_active, E * N, &p.joint_active) ||
        !in_ptr(b, kp, E * N, &p.kp) || !in_ptr(b, kd, E * N, &p.kd) ||
        !out_ptr(b, inout_tau, E * n, &p.tau, copies, /*preload=*/true))
        return TRL_ERR_CUDA;
    if (!tar_vel) p.tar_vel = b->d_zero_nvec;
    int rc = trl_launch_exp_pd(b->d_model, b->model->dev, b->E, p, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

extern "C" int trl_rbd_end_effector_jacobian(trl_batch* b, const double* pose, const int* joint_ids, int joint_id_all,
                                             double* out_J) {
    if (!b || !pose || !out_J) { trl_set_error("trl_rbd_end_effector_jacobian: null argument"); return TRL_ERR_INVALID_ARG; }
    if (!joint_ids && (joint_id_all < 0 || joint_id_all >= b->model->N)) {
        trl_set_error("trl_rbd_end_effector_jacobian: joint id out of range");
        return TRL_ERR_INVALID_ARG;
    }
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double* d_pose = nullptr;
    const int* d_ids = nullptr;
    double* d_J = nullptr;
    if (!in_ptr(b, pose, E * n, &d_pose) || !in_ptr(b, joint_ids, E, &d_ids) || !out_ptr(b, out_J, E * 6 * n, &d_J, copies))
        return TRL_ERR_CUDA;
    // the chain's columns of the world Jacobian (k_rbd_extras), everything off the chain zeroed
    int rc = trl_launch_rbd_extras(b->d_model, b->model->dev, b->E, d_pose, d_J, nullptr, nullptr, b->stream);
    if (!rc) rc = trl_launch_ee_mask(b->d_model, b->E, (int)n, d_ids, joint_id_all, d_J, b->stream);
    b->launches += 2;
    return finish(b, copies, rc);
}

extern "C" int trl_rbd_extras(trl_batch* b, const double* pose, double* out_J, double* out_Jcom, double* out_gravity_force) {
    if (!b || !pose) { trl_set_error("trl_rbd_extras: null argument"); return TRL_ERR_INVALID_ARG; }
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double* d_pose = nullptr;
    double *d_J = nullptr, *d_Jc = nullptr, *d_g = nullptr;
    if (!in_ptr(b, pose, E * n, &d_pose) || !out_ptr(b, out_J, E * 6 * n, &d_J, copies) ||
        !out_ptr(b, out_Jcom, E * 6 * n, &d_Jc, copies) || !out_ptr(b, out_gravity_force, E * n, &d_g, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_rbd_extras(b->d_model, b->model->dev, b->E, d_pose, d_J, d_Jc, d_g, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

extern "C" int trl_kin_tree_fk(trl_batch* b, const double* pose, double* out_joint_world, double* out_body_pos) {
    if (!b || !pose) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double* d_pose;
    double *d_jw, *d_bp;
    if (!in_ptr(b, pose, E * n, &d_pose) || !out_ptr(b, out_joint_world, E * N * 12, &d_jw, copies) ||
        !out_ptr(b, out_body_pos, E * N * 3, &d_bp, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_fk(b->d_model, b->model->dev, b->E, d_pose, d_jw, d_bp, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

extern "C" int trl_kin_char_pose(trl_batch* b, const double* time, const double* origin_pos, const double* origin_rot,
                                 double* out_pose, double* out_vel) {
    if (!b || !time || !out_pose) return TRL_ERR_INVALID_ARG;
    if (b->model->num_frames < 2) { trl_set_error("model has no motion"); return TRL_ERR_NO_MOTION; }
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double *d_t, *d_op, *d_or;
    double *d_p, *d_v;
    if (!in_ptr(b, time, E, &d_t) || !in_ptr(b, origin_pos, E * 3, &d_op) || !in_ptr(b, origin_rot, E * 4, &d_or) ||
        !out_ptr(b, out_pose, E * n, &d_p, copies) || !out_ptr(b, out_vel, E * n, &d_v, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_kin_char(b->d_model, b->model->dev, b->d_frames, b->E, d_t, d_op, d_or, d_p, d_v, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

// Shared memory per env for the cell rectangles k_obs_samples stages: the sample window (diagonal D in any heading)
// covers at most D / spacing + 5 cells per axis, + 3 because a staged row starts at a column multiple of 4.
static void set_patch_dims(trl_batch* b, int kind, double min_sx, double min_sz, int pieces_per_field) {
    const double cx = std::floor(b->sample_diag / min_sx) + 8.0;
    const double cz = (kind == 2) ? std::floor(b->sample_diag / min_sz) + 5.0 : 1.0;
    b->ground.patch_w = b->ground.patch_h = b->ground.patch_nbuf = 0;
    b->ground.table_ok = b->sample_table_ok;
    if (b->cfg.num_samples > 0 && cx <= 1024.0 && cz <= 64.0) {
        const int w = ((int)cx + 3) & ~3, h = (int)cz;
        const int nbuf = std::min(pieces_per_field, (int)(8192 / ((size_t)w * h * sizeof(float))));  // <= 8 KB per env
        if (nbuf >= 1) { b->ground.patch_w = w; b->ground.patch_h = h; b->ground.patch_nbuf = nbuf; }
    }
}

static void release_ground(trl_batch* b) {
    if (b->d_ground_data) { cudaFree(b->d_ground_data); b->d_ground_data = nullptr; }
    if (b->d_pieces) { cudaFree(b->d_pieces); b->d_pieces = nullptr; }
    if (b->d_env_to_field) { cudaFree(b->d_env_to_field); b->d_env_to_field = nullptr; }
    if (b->d_tstate) { cudaFree(b->d_tstate); b->d_tstate = nullptr; }
    if (b->d_tchanged) { cudaFree(b->d_tchanged); b->d_tchanged = nullptr; }
    if (b->d_tbounds) { cudaFree(b->d_tbounds); b->d_tbounds = nullptr; }
    std::memset(&b->ground, 0, sizeof(b->ground));
}

extern "C" int trl_ground_set_heightfields(trl_batch* b, int kind, int num_fields, int pieces_per_field,
                                           const float* data, long long data_count, const long long* data_offset,
                                           const int* res, const double* origin, const double* scale,
                                           const double* bounds, const int* env_to_field) {
    if (!b) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    cudaStreamSynchronize(b->stream);
    release_ground(b);
    b->ground.kind = kind;
    if (kind == 0 || kind == 3) return TRL_OK;
    if (data_count >= (1LL << 32)) {
        trl_set_error("trl_ground_set_heightfields: more than 2^32 height values");
        b->ground.kind = 0;
        return TRL_ERR_UNSUPPORTED;
    }
    if ((kind != 1 && kind != 2) || num_fields < 1 || pieces_per_field < 1 || pieces_per_field > 4 || !data ||
        data_count < 1 || !data_offset || !res || !origin || !scale || !bounds) {
        trl_set_error("trl_ground_set_heightfields: invalid argument");
        b->ground.kind = 0;
        return TRL_ERR_INVALID_ARG;
    }
    const size_t np = (size_t)num_fields * pieces_per_field;
    std::vector<DevPiece> pieces(np);
    for (size_t i = 0; i < np; ++i) {
        DevPiece& pc = pieces[i];
        pc.data_offset = data_offset[i];
        pc.res_x = res[2 * i];
        pc.res_z = res[2 * i + 1];
        // var2d reads row 0 only; var3d reads rows j and j + 1, so it needs at least two
        if (pc.res_x < 2 || pc.res_z < (kind == 2 ? 2 : 1) || pc.data_offset < 0 ||
            pc.data_offset + (long long)pc.res_x * pc.res_z > data_count) {
            trl_set_error("trl_ground_set_heightfields: piece outside the data array (or fewer than 2 x 2 / 2 x 1 vertices)");
            b->ground.kind = 0;
            return TRL_ERR_INVALID_ARG;
        }
        for (int k = 0; k < 3; ++k) { pc.origin[k] = origin[3 * i + k]; pc.scale[k] = scale[3 * i + k]; }
        if (!std::isfinite(pc.scale[0]) || pc.scale[0] == 0.0 || !std::isfinite(pc.scale[1]) ||
            !std::isfinite(pc.scale[2]) || pc.scale[2] == 0.0) {
            trl_set_error("trl_ground_set_heightfields: grid spacing must be finite and non-zero");
            b->ground.kind = 0;
            return TRL_ERR_INVALID_ARG;
        }
        for (int k = 0; k < 4; ++k) pc.bounds[k] = bounds[4 * i + k];
        pc.rscale[0] = 1.0 / pc.scale[0];
        pc.rscale[1] = 1.0 / pc.scale[2];
        pc.half[0] = (double)(pc.res_x - 1) * 0.5;
        pc.half[1] = (double)(pc.res_z - 1) * 0.5;
        pc.cmax3[0] = pc.res_x - 1.0 - 0.0001;
        pc.cmax3[1] = pc.res_z - 1.0 - 0.0001;
    }
    // Device layout: pieces on 16-byte boundaries, rows `pitch` floats apart, and a tail pad so that a fixed-size row
    // segment read near the end of the last piece stays inside the allocation (see DevPiece).
    std::vector<float> packed;
    {
        size_t total = 0;
        for (size_t i = 0; i < np; ++i) {
            DevPiece& pc = pieces[i];
            pc.pitch = (kind == 2) ? ((pc.res_x + 3) & ~3) : pc.res_x;  // var2d reads row 0 only: rows stay packed
            total = (total + 3) & ~(size_t)3;
            total += (size_t)pc.pitch * pc.res_z;
        }
        const size_t tail = 4096;
        if (total + tail >= ((size_t)1 << 32)) {
            trl_set_error("trl_ground_set_heightfields: more than 2^32 height values");
            b->ground.kind = 0;
            return TRL_ERR_UNSUPPORTED;
        }
        packed.assign(total + tail, 0.0f);
        size_t at = 0;
        for (size_t i = 0; i < np; ++i) {
            DevPiece& pc = pieces[i];
            at = (at + 3) & ~(size_t)3;
            const float* src = data + pc.data_offset;
            for (int j = 0; j < pc.res_z; ++j)
                std::memcpy(packed.data() + at + (size_t)j * pc.pitch, src + (size_t)j * pc.res_x, sizeof(float) * pc.res_x);
            pc.data_offset = (long long)at;
            at += (size_t)pc.pitch * pc.res_z;
        }
    }
    if (!cuda_ok(cudaMalloc((void**)&b->d_ground_data, sizeof(float) * packed.size()), "cudaMalloc(heightfields)") ||
        !cuda_ok(cudaMemcpy(b->d_ground_data, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice), "H2D(heightfields)") ||
        !cuda_ok(cudaMalloc((void**)&b->d_pieces, sizeof(DevPiece) * np), "cudaMalloc(pieces)") ||
        !cuda_ok(cudaMemcpy(b->d_pieces, pieces.data(), sizeof(DevPiece) * np, cudaMemcpyHostToDevice), "H2D(pieces)")) {
        b->ground.kind = 0;
        return TRL_ERR_CUDA;
    }
    {
        std::vector<int> e2f((size_t)b->E);
        for (int e = 0; e < b->E; ++e) {
            e2f[e] = env_to_field ? env_to_field[e] : (e % num_fields);
            if (e2f[e] < 0 || e2f[e] >= num_fields) {
                trl_set_error("trl_ground_set_heightfields: env_to_field out of range");
                b->ground.kind = 0;
                return TRL_ERR_INVALID_ARG;
            }
        }
        if (!cuda_ok(cudaMalloc((void**)&b->d_env_to_field, sizeof(int) * (size_t)b->E), "cudaMalloc(env_to_field)") ||
            !cuda_ok(cudaMemcpy(b->d_env_to_field, e2f.data(), sizeof(int) * (size_t)b->E, cudaMemcpyHostToDevice), "H2D(env_to_field)")) {
            b->ground.kind = 0;
            return TRL_ERR_CUDA;
        }
    }
    {
        double min_sx = HUGE_VAL, min_sz = HUGE_VAL;
        for (size_t i = 0; i < np; ++i) {
            min_sx = std::min(min_sx, std::fabs(pieces[i].scale[0]));
            min_sz = std::min(min_sz, std::fabs(pieces[i].scale[2]));
        }
        set_patch_dims(b, kind, min_sx, min_sz, pieces_per_field);
    }
    b->ground.num_fields = num_fields;
    b->ground.pieces_per_field = pieces_per_field;
    b->ground.data = b->d_ground_data;
    b->ground.pieces = b->d_pieces;
    b->ground.env_to_field = b->d_env_to_field;
    return TRL_OK;
}

// ------------------------------------------------------------------------------------------------
// procedural terrain on the device (kernels: trl_terrain.cu)
// ------------------------------------------------------------------------------------------------
static int upload_env_to_field(trl_batch* b, int num_fields, const int* env_to_field, const char* who) {
    std::vector<int> e2f((size_t)b->E);
    for (int e = 0; e < b->E; ++e) {
        e2f[e] = env_to_field ? env_to_field[e] : (e % num_fields);
        if (e2f[e] < 0 || e2f[e] >= num_fields) {
            trl_set_error(std::string(who) + ": env_to_field out of range");
            return TRL_ERR_INVALID_ARG;
        }
    }
    if (!cuda_ok(cudaMalloc((void**)&b->d_env_to_field, sizeof(int) * (size_t)b->E), "cudaMalloc(env_to_field)") ||
        !cuda_ok(cudaMemcpy(b->d_env_to_field, e2f.data(), sizeof(int) * (size_t)b->E, cudaMemcpyHostToDevice), "H2D(env_to_field)"))
        return TRL_ERR_CUDA;
    return TRL_OK;
}

extern "C" int trl_ground_generate(trl_batch* b, const trl_terrain_cfg* cfg, int num_fields, const unsigned long long* seeds,
                                   const double* bound_min, const double* bound_max, const int* env_to_field) {
    if (!b || !cfg) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    cudaStreamSynchronize(b->stream);
    release_ground(b);
    if (cfg->type == TRL_TERRAIN_PLANE) { b->ground.kind = 3; return TRL_OK; }
    const bool is3d = cfg->type == TRL_TERRAIN_VAR3D_FLAT;
    if (cfg->type < TRL_TERRAIN_VAR2D_FLAT || cfg->type > TRL_TERRAIN_VAR3D_FLAT) {
        trl_set_error("trl_ground_generate: unsupported terrain type");
        return TRL_ERR_UNSUPPORTED;
    }
    if (num_fields < 1 || !bound_min || !bound_max || (!is3d && !seeds) || !(cfg->world_scale > 0) ||
        !(cfg->ground_width > 0) || (is3d && (!(cfg->spacing_x > 0) || !(cfg->spacing_z > 0)))) {
        trl_set_error("trl_ground_generate: invalid argument");
        return TRL_ERR_INVALID_ARG;
    }
    TerrainGenCfg& t = b->tgen;
    t.type = cfg->type;
    t.pieces = is3d ? 4 : 2;
    for (int i = 0; i < TRL_TERRAIN_NUM_PARAMS; ++i) t.params[i] = cfg->params[i];
    t.world_scale = cfg->world_scale;
    t.spacing_x = cfg->spacing_x;
    t.spacing_z = cfg->spacing_z;
    if (is3d) {
        t.ground_width = 40;  // cGroundVar3D::Init forces mGroundWidth (sim/GroundVar3D.cpp:24-36)
        const long long rx = (long long)std::ceil(t.ground_width / t.spacing_x - 0.000001) + 1;
        const long long rz = (long long)std::ceil(t.ground_width / t.spacing_z - 0.000001) + 1;
        t.slot_floats = ((rx + 3) & ~3LL) * rz;
    } else {
        // a builder appends whole features until the width is reached, after a flat lead-in of up to the whole width when
        // the segment spans the origin (AddPadding): room for twice the width and the longest single feature
        t.ground_width = cfg->ground_width;
        const double* p = t.params;
        const double feature = std::max({p[1] + p[3], p[7] + p[9], p[13], p[21] + std::max(1.0, p[29]) * (p[23] + p[25]),
                                         p[31] + std::max(0.0, p[36]) * 0.1 + 0.5, 1.0});
        const double verts = std::ceil((2.0 * t.ground_width + 1.0 + feature) / 0.025) + 8.0 * (std::max(1.0, p[29]) + std::max(0.0, p[36])) + 64.0;
        if (!(verts < 1.0e6)) { trl_set_error("trl_ground_generate: terrain parameters out of range"); return TRL_ERR_INVALID_ARG; }
        t.slot_floats = ((long long)verts + 3) & ~3LL;
    }
    const size_t F = (size_t)num_fields, P = (size_t)t.pieces;
    const size_t floats = F * P * (size_t)t.slot_floats + 4096;
    if (floats >= ((size_t)1 << 32)) { trl_set_error("trl_ground_generate: more than 2^32 height values"); return TRL_ERR_UNSUPPORTED; }
    std::vector<unsigned long long> zero_seeds;
    if (!seeds) { zero_seeds.assign(F, 0ULL); seeds = zero_seeds.data(); }
    unsigned long long* d_seeds = nullptr;
    bool ok = cuda_ok(cudaMalloc((void**)&b->d_ground_data, sizeof(float) * floats), "cudaMalloc(heightfields)") &&
              cuda_ok(cudaMemset(b->d_ground_data, 0, sizeof(float) * floats), "cudaMemset(heightfields)") &&
              cuda_ok(cudaMalloc((void**)&b->d_pieces, sizeof(DevPiece) * F * P), "cudaMalloc(pieces)") &&
              cuda_ok(cudaMalloc((void**)&b->d_tstate, sizeof(TerrainState) * F), "cudaMalloc(terrain state)") &&
              cuda_ok(cudaMemset(b->d_tstate, 0, sizeof(TerrainState) * F), "cudaMemset(terrain state)") &&
              cuda_ok(cudaMalloc((void**)&b->d_tchanged, sizeof(int) * F), "cudaMalloc(terrain flags)") &&
              cuda_ok(cudaMalloc((void**)&b->d_tbounds, sizeof(double) * 6 * F), "cudaMalloc(terrain bounds)") &&
              cuda_ok(cudaMalloc((void**)&d_seeds, sizeof(unsigned long long) * F), "cudaMalloc(seeds)") &&
              cuda_ok(cudaMemcpy(d_seeds, seeds, sizeof(unsigned long long) * F, cudaMemcpyHostToDevice), "H2D(seeds)") &&
              cuda_ok(cudaMemcpy(b->d_tbounds, bound_min, sizeof(double) * 3 * F, cudaMemcpyHostToDevice), "H2D(bounds)") &&
              cuda_ok(cudaMemcpy(b->d_tbounds + 3 * F, bound_max, sizeof(double) * 3 * F, cudaMemcpyHostToDevice), "H2D(bounds)");
    int rc = TRL_OK;
    if (ok) {
        const int lrc = trl_launch_terrain(t, num_fields, d_seeds, b->d_tbounds, b->d_tbounds + 3 * F, b->d_tstate, b->d_ground_data,
                                           b->d_pieces, b->d_tchanged, 1, b->stream);
        b->launches += 1;
        ok = lrc == 0 && cuda_ok(cudaStreamSynchronize(b->stream), "terrain generation");
        if (lrc) trl_set_error(std::string("trl_ground_generate: ") + cudaGetErrorString((cudaError_t)lrc));
    }
    if (d_seeds) cudaFree(d_seeds);
    if (ok) {  // a piece that did not fit its slot is an error, like an out-of-memory in the reference
        std::vector<TerrainState> st(F);
        ok = cuda_ok(cudaMemcpy(st.data(), b->d_tstate, sizeof(TerrainState) * F, cudaMemcpyDeviceToHost), "D2H(terrain state)");
        for (size_t f = 0; ok && f < F; ++f)
            if (st[f].overflow) { trl_set_error("trl_ground_generate: a terrain piece exceeds its slot"); rc = TRL_ERR_UNSUPPORTED; ok = false; }
    }
    if (ok) rc = upload_env_to_field(b, num_fields, env_to_field, "trl_ground_generate");
    if (!ok || rc != TRL_OK) {
        release_ground(b);
        return rc != TRL_OK ? rc : TRL_ERR_CUDA;
    }
    b->ground.kind = is3d ? 2 : 1;
    const double sx = is3d ? (double)(float)(t.spacing_x * t.world_scale) / t.world_scale : (double)(0.025f * (float)1.0) * 1.0;
    const double sz = is3d ? (double)(float)(t.spacing_z * t.world_scale) / t.world_scale : 1.0;
    set_patch_dims(b, b->ground.kind, std::fabs(sx) * 0.999, std::fabs(sz) * 0.999, t.pieces);
    b->ground.num_fields = num_fields;
    b->ground.pieces_per_field = t.pieces;
    b->ground.data = b->d_ground_data;
    b->ground.pieces = b->d_pieces;
    b->ground.env_to_field = b->d_env_to_field;
    return TRL_OK;
}

extern "C" int trl_ground_update(trl_batch* b, const double* bound_min, const double* bound_max, int* out_changed) {
    if (!b || !bound_min || !bound_max) return TRL_ERR_INVALID_ARG;
    if (!b->d_tstate) { trl_set_error("trl_ground_update: no generated terrain (trl_ground_generate)"); return TRL_ERR_INVALID_ARG; }
    Guard g(b->device);
    const size_t F = (size_t)b->ground.num_fields;
    const double *d_min = bound_min, *d_max = bound_max;
    int* d_changed = out_changed;
    if (b->ptr_mode == TRL_PTR_HOST) {
        if (!cuda_ok(cudaMemcpyAsync(b->d_tbounds, bound_min, sizeof(double) * 3 * F, cudaMemcpyHostToDevice, b->stream), "H2D(bounds)") ||
            !cuda_ok(cudaMemcpyAsync(b->d_tbounds + 3 * F, bound_max, sizeof(double) * 3 * F, cudaMemcpyHostToDevice, b->stream), "H2D(bounds)"))
            return TRL_ERR_CUDA;
        d_min = b->d_tbounds;
        d_max = b->d_tbounds + 3 * F;
        d_changed = b->d_tchanged;
    }
    const int lrc = trl_launch_terrain(b->tgen, (int)F, nullptr, d_min, d_max, b->d_tstate, b->d_ground_data, b->d_pieces, d_changed,
                                       0, b->stream);
    b->launches += 1;
    if (lrc) { trl_set_error(std::string("trl_ground_update: ") + cudaGetErrorString((cudaError_t)lrc)); return TRL_ERR_CUDA; }
    if (b->ptr_mode == TRL_PTR_HOST) {
        std::vector<int> flags(F);
        if (!cuda_ok(cudaMemcpyAsync(flags.data(), b->d_tchanged, sizeof(int) * F, cudaMemcpyDeviceToHost, b->stream), "D2H(flags)") ||
            !cuda_ok(cudaStreamSynchronize(b->stream), "cudaStreamSynchronize"))
            return TRL_ERR_CUDA;
        bool overflow = false;
        for (size_t f = 0; f < F; ++f) {
            overflow = overflow || flags[f] == 2;
            if (out_changed) out_changed[f] = flags[f];
        }
        if (overflow) { trl_set_error("trl_ground_update: a terrain piece exceeds its slot"); return TRL_ERR_UNSUPPORTED; }
    }
    return TRL_OK;
}

extern "C" long long trl_ground_read_field(trl_batch* b, int field, float* out_data, long long capacity, long long* data_offset,
                                           int* res, double* origin, double* scale, double* bounds) {
    if (!b || field < 0 || field >= b->ground.num_fields || (b->ground.kind != 1 && b->ground.kind != 2)) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    if (!cuda_ok(cudaStreamSynchronize(b->stream), "cudaStreamSynchronize")) return TRL_ERR_CUDA;
    const int P = b->ground.pieces_per_field;
    std::vector<DevPiece> pcs((size_t)P);
    if (!cuda_ok(cudaMemcpy(pcs.data(), b->d_pieces + (size_t)field * P, sizeof(DevPiece) * P, cudaMemcpyDeviceToHost), "D2H(pieces)"))
        return TRL_ERR_CUDA;
    long long at = 0;
    std::vector<float> row;
    for (int s = 0; s < P; ++s) {
        const DevPiece& pc = pcs[s];
        const long long count = (long long)pc.res_x * pc.res_z;
        if (data_offset) data_offset[s] = at;
        if (res) { res[2 * s] = pc.res_x; res[2 * s + 1] = pc.res_z; }
        for (int k = 0; k < 3; ++k) {
            if (origin) origin[3 * s + k] = pc.origin[k];
            if (scale) scale[3 * s + k] = pc.scale[k];
        }
        for (int k = 0; k < 4; ++k)
            if (bounds) bounds[4 * s + k] = pc.bounds[k];
        if (out_data) {
            if (at + count > capacity) { trl_set_error("trl_ground_read_field: capacity too small"); return TRL_ERR_INVALID_ARG; }
            // device rows are `pitch` floats apart; a var2d segment keeps one row on the device (the reference's 3 are identical)
            const int dev_rows = (b->ground.kind == 1) ? 1 : pc.res_z;
            row.resize((size_t)pc.pitch * dev_rows);
            if (!cuda_ok(cudaMemcpy(row.data(), b->d_ground_data + pc.data_offset, sizeof(float) * row.size(), cudaMemcpyDeviceToHost), "D2H(heights)"))
                return TRL_ERR_CUDA;
            for (int j = 0; j < pc.res_z; ++j)
                std::memcpy(out_data + at + (long long)j * pc.res_x, row.data() + (size_t)(dev_rows == 1 ? 0 : j) * pc.pitch, sizeof(float) * pc.res_x);
        }
        at += count;
    }
    return at;
}

extern "C" int trl_ground_sample_height(trl_batch* b, const double* pos, int S, double* out_h,
                                        unsigned char* out_valid, int* out_cell) {
    if (!b || !pos || !out_h || S < 1) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    const size_t E = b->E;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double* d_pos;
    double* d_h;
    unsigned char* d_v;
    int* d_c;
    if (!in_ptr(b, pos, E * S * 3, &d_pos) || !out_ptr(b, out_h, E * S, &d_h, copies) ||
        !out_ptr(b, out_valid, E * S, &d_v, copies) || !out_ptr(b, out_cell, E * S * 3, &d_c, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_sample_height(b->ground, b->E, d_pos, S, d_h, d_v, d_c, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

extern "C" int trl_build_poli_state(trl_batch* b, const double* pose, const double* vel, const double* body_pos,
                                    const double* body_vel, const double* body_quat, const double* body_angvel,
                                    const double* ground_vel, const double* phase, const unsigned char* flip,
                                    double* out_state, int* out_cell) {
    if (!b || !pose || !body_pos || !body_vel || !out_state) return TRL_ERR_INVALID_ARG;
    if (b->cfg.obs_kind == TRL_OBS_CT_3D && (!body_quat || !body_angvel)) {
        trl_set_error("TRL_OBS_CT_3D needs body_quat and body_angvel");
        return TRL_ERR_INVALID_ARG;
    }
    if (b->cfg.obs_kind == TRL_OBS_TERRAIN_RL_HOPPER && !vel) {
        trl_set_error("TRL_OBS_TERRAIN_RL_HOPPER needs vel");
        return TRL_ERR_INVALID_ARG;
    }
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double *d_pose, *d_vel, *d_bp, *d_bv, *d_bq, *d_ba, *d_gv, *d_ph;
    const unsigned char* d_fl;
    double* d_state;
    int* d_cell;
    if (!in_ptr(b, pose, E * n, &d_pose) || !in_ptr(b, vel, E * n, &d_vel) || !in_ptr(b, body_pos, E * N * 3, &d_bp) ||
        !in_ptr(b, body_vel, E * N * 3, &d_bv) || !in_ptr(b, body_quat, E * N * 4, &d_bq) ||
        !in_ptr(b, body_angvel, E * N * 3, &d_ba) || !in_ptr(b, ground_vel, E * 3, &d_gv) || !in_ptr(b, phase, E, &d_ph) ||
        !in_ptr(b, flip, E, &d_fl) || !out_ptr(b, out_state, E * (size_t)b->F, &d_state, copies) ||
        !out_ptr(b, out_cell, E * (size_t)b->cfg.num_samples * 3, &d_cell, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_poli_state(b->d_model, b->model->dev, b->ground, b->cfg, b->F, b->E, d_pose, d_vel, d_bp, d_bv,
                                   d_bq, d_ba, d_gv, d_ph, d_fl, b->d_sample_local, b->d_env_rec, d_state, d_cell, b->stream);
    b->launches += (b->cfg.num_samples > 0) ? 3 : 2;
    return finish(b, copies, rc);
}

// HOST-pointer step, pipelined over env slices. The step is bound by PCIe (biped3D x 16384 envs: 20 MB up, 50 MB down
// per step against ~0.1 ms of kernels), so the batch is cut into `host_slices` ranges of whole 32-env tiles and run as
// a three-stage pipeline: one upload stream copies slice after slice, the three chain streams compute a slice as soon
// as its inputs have landed, and one download stream copies results back as each chain of a slice finishes. The copy
// engines of the two directions and the SMs all work on different slices at the same time; the critical path is one
// slice's upload + kernels plus the whole download.
static int step_fused_host_enqueue(trl_batch* b, const trl_step_args* a, const trl_step_opts* o, bool want_pd, bool want_kin,
                                   bool want_obs, bool capturing) {
    const size_t E = b->E, n = b->model->n, N = b->model->N, F = (size_t)b->F;
    b->pool_next = 0;
    auto dev = [&](const void* host, size_t bytes_per_env) -> char* {
        return host ? static_cast<char*>(b->stage_out(E * bytes_per_env)) : nullptr;
    };
    bool ok = true;
    auto need = [&](const void* host, char* d) { if (host && !d) ok = false; };
    char* d_pose = dev(a->pose, n * 8);            need(a->pose, d_pose);
    char* d_vel = dev(a->vel, n * 8);              need(a->vel, d_vel);
    char* d_tp = want_pd ? dev(a->tar_pose, n * 8) : nullptr;
    char* d_tv = want_pd ? (a->tar_vel ? dev(a->tar_vel, n * 8) : reinterpret_cast<char*>(b->d_zero_nvec)) : nullptr;
    const bool want_reward = o && o->out_reward, want_f32 = o && o->out_state_f32;
    auto scratch = [&](bool wanted, size_t bytes_per_env) -> char* {
        return wanted ? static_cast<char*>(b->stage_out(E * bytes_per_env)) : nullptr;
    };
    char* d_ja = want_pd ? dev(a->joint_active, N) : nullptr;
    char* d_tau = want_pd ? dev(a->out_tau, n * 8) : nullptr;
    char* d_time = want_kin ? dev(a->kin_time, 8) : nullptr;
    char* d_op = want_kin ? dev(a->origin_pos, 24) : nullptr;
    char* d_or = want_kin ? dev(a->origin_rot, 32) : nullptr;
    // kin outputs the caller did not ask for stay on the device (the reward epilogue reads them there)
    char* d_kp = scratch(want_kin && (a->out_kin_pose || want_reward), n * 8);
    char* d_kv = scratch(want_kin && (a->out_kin_vel || want_reward), n * 8);
    char* d_kbp = scratch(want_kin && a->out_kin_body_pos, N * 24);
    char* d_bp = want_obs ? dev(a->body_pos, N * 24) : nullptr;
    char* d_bv = (want_obs || want_reward) ? dev(a->body_vel, N * 24) : nullptr;
    char* d_ph = want_obs ? dev(a->phase, 8) : nullptr;
    char* d_fl = want_obs ? dev(a->flip, 1) : nullptr;
    char* d_st = scratch(want_obs, F * 8);
    char* d_st32 = scratch(want_f32, F * 4);
    char* d_rw = scratch(want_reward, 8);
    char* d_fallen = want_reward ? dev(o->fallen, 1) : nullptr;
    if (want_pd) { need(a->tar_pose, d_tp); need(a->tar_vel, d_tv); need(a->joint_active, d_ja); need(a->out_tau, d_tau); }
    if (want_kin) { need(a->kin_time, d_time); need(a->origin_pos, d_op); need(a->origin_rot, d_or); need(a->out_kin_pose, d_kp); need(a->out_kin_vel, d_kv); need(a->out_kin_body_pos, d_kbp); }
    if (want_obs) { need(a->body_pos, d_bp); need(a->body_vel, d_bv); need(a->phase, d_ph); need(a->flip, d_fl); need(a->body_pos, d_st); }
    if (want_reward) { need(a->body_vel, d_bv); need(o->out_reward, d_rw); need(o->out_reward, d_kp); need(o->out_reward, d_kv); need(o->fallen, d_fallen); }
    if (want_f32) need(o->out_state_f32, d_st32);
    if (!ok) { trl_set_error("trl_step_fused: device staging allocation failed"); return TRL_ERR_CUDA; }

    const size_t tile = (size_t)trl_pd_scratch_tile();
    const size_t S = (size_t)std::min(b->host_slices, (int)trl_batch::kMaxSlices);
    const size_t es = ((E + S - 1) / S + tile - 1) / tile * tile;  // envs per slice, whole tiles
    cudaStream_t sB = b->side[0], sC = b->side[1], sA = b->side[2];  // kin, observation, stable-PD chains
    // uploads: stream 0 = pose / vel (+ the small arrays), 1 = PD targets, 2 = body states; download: one stream
    cudaStream_t s_up0 = b->up_streams[0], s_up1 = b->up_streams[1], s_up2 = b->up_streams[2], s_dn = b->side2[1];
    if (!cuda_ok(cudaEventRecord(b->ev_fork, b->stream), "cudaEventRecord")) return TRL_ERR_CUDA;
    for (cudaStream_t st : {sA, sB, sC, s_up0, s_up1, s_up2, s_dn})
        if (!cuda_ok(cudaStreamWaitEvent(st, b->ev_fork, 0), "cudaStreamWaitEvent")) return TRL_ERR_CUDA;
    int rc = 0;
    // every event / dependency call is checked: the first failure is kept in rc and reported after the join
    auto ev_rec = [&](cudaEvent_t ev, cudaStream_t st) {
        const cudaError_t err = cudaEventRecord(ev, st);
        if (err != cudaSuccess && !rc) rc = (int)err;
    };
    auto st_wait = [&](cudaStream_t st, cudaEvent_t ev) {
        const cudaError_t err = cudaStreamWaitEvent(st, ev, 0);
        if (err != cudaSuccess && !rc) rc = (int)err;
    };
    cudaStream_t s_up = s_up0;
    auto up = [&](char* d, const void* h, size_t bpe, size_t e0, size_t ec) {
        if (d && !rc && cudaMemcpyAsync(d + e0 * bpe, static_cast<const char*>(h) + e0 * bpe, ec * bpe, cudaMemcpyHostToDevice, s_up) != cudaSuccess)
            rc = (int)cudaGetLastError();
    };
    auto down = [&](void* h, const char* d, size_t bpe, size_t e0, size_t ec) {
        if (h && d && !rc && cudaMemcpyAsync(static_cast<char*>(h) + e0 * bpe, d + e0 * bpe, ec * bpe, cudaMemcpyDeviceToHost, s_dn) != cudaSuccess)
            rc = (int)cudaGetLastError();
    };
    // the small per-env arrays go up once, whole (1.3 MB for 16384 envs): fewer copies per slice
    if (want_obs) { up(d_ph, a->phase, 8, 0, E); up(d_fl, a->flip, 1, 0, E); }
    if (want_kin) { up(d_time, a->kin_time, 8, 0, E); up(d_op, a->origin_pos, 24, 0, E); up(d_or, a->origin_rot, 32, 0, E); }
    if (want_pd) up(d_ja, a->joint_active, N, 0, E);
    if (want_reward) up(d_fallen, o->fallen, 1, 0, E);
    const size_t rec_bytes = trl_obs_env_rec_bytes();
    const size_t sc_stride = trl_pd_scratch_doubles(b->model->dev);
    int slice = 0;
    for (size_t e0 = 0; e0 < E && !rc; e0 += es, ++slice) {
        const size_t ec = std::min(es, E - e0);
        s_up = s_up0;
        up(d_pose, a->pose, n * 8, e0, ec);
        up(d_vel, a->vel, n * 8, e0, ec);
        ev_rec(b->ev_up[slice][0], s_up0);
        // submission order = service order on the copy engines: the observation chain's inputs go first (its [E][F]
        // state is 3/4 of the download, which is the long pole), the PD targets after them
        if (want_obs || want_reward) {
            s_up = s_up2;
            if (want_obs) up(d_bp, a->body_pos, N * 24, e0, ec);
            up(d_bv, a->body_vel, N * 24, e0, ec);
            ev_rec(b->ev_up[slice][2], s_up2);
        }
        if (want_pd) {
            s_up = s_up1;
            up(d_tp, a->tar_pose, n * 8, e0, ec);
            if (a->tar_vel) up(d_tv, a->tar_vel, n * 8, e0, ec);
            ev_rec(b->ev_up[slice][1], s_up1);
        }
        auto at = [&](char* d, size_t bpe) { return d ? d + e0 * bpe : nullptr; };
        if (!rc && want_obs) {
            st_wait(sC, b->ev_up[slice][0]);
            st_wait(sC, b->ev_up[slice][2]);
            DevGround g2 = b->ground;
            if (g2.env_to_field) g2.env_to_field += e0;
            rc = trl_launch_poli_state(b->d_model, b->model->dev, g2, b->cfg, b->F, (int)ec, (const double*)at(d_pose, n * 8),
                                       (const double*)at(d_vel, n * 8), (const double*)at(d_bp, N * 24),
                                       (const double*)at(d_bv, N * 24), nullptr, nullptr, nullptr, (const double*)at(d_ph, 8),
                                       (const unsigned char*)at(d_fl, 1), b->d_sample_local,
                                       static_cast<char*>(b->d_env_rec) + e0 * rec_bytes, (double*)at(d_st, F * 8), nullptr, sC);
            b->launches += (b->cfg.num_samples > 0) ? 3 : 2;
            if (!rc && want_f32) {
                rc = trl_launch_f64_to_f32((const double*)at(d_st, F * 8), (float*)at(d_st32, F * 4), ec * F, sC);
                b->launches += 1;
            }
            ev_rec(b->ev_cmp[slice][1], sC);
        }
        if (!rc && want_kin) {
            st_wait(sB, b->ev_up[slice][0]);  // the small arrays went up ahead of slice 0 on stream 0
            rc = trl_launch_kin(b->d_model, b->model->dev, b->d_frames, (int)ec, (const double*)at(d_time, 8),
                                (const double*)at(d_op, 24), (const double*)at(d_or, 32), (double*)at(d_kp, n * 8),
                                (double*)at(d_kv, n * 8), (double*)at(d_kbp, N * 24), sB);
            b->launches += 1;
            if (!rc && want_reward) {  // cScenarioExpImitate::CalcReward on the kin chain's stream: kin pose / vel never leave the device
                st_wait(sB, b->ev_up[slice][2]);
                DevGround g2 = b->ground;
                if (g2.env_to_field) g2.env_to_field += e0;
                rc = trl_launch_imitate_reward(b->d_model, b->model->dev, g2, (int)ec, (const double*)at(d_pose, n * 8),
                                               (const double*)at(d_vel, n * 8), (const double*)at(d_kp, n * 8),
                                               (const double*)at(d_kv, n * 8), (const double*)at(d_bv, N * 24),
                                               (const unsigned char*)at(d_fallen, 1), (double*)at(d_rw, 8), nullptr, sB);
                b->launches += 1;
            }
            ev_rec(b->ev_cmp[slice][0], sB);
        }
        if (!rc && want_pd) {
            st_wait(sA, b->ev_up[slice][0]);
            st_wait(sA, b->ev_up[slice][1]);
            if (a->dt > 0) {
                StepPtrs p{};
                p.dt = a->dt;
                p.pose = (const double*)at(d_pose, n * 8); p.vel = (const double*)at(d_vel, n * 8);
                p.tar_pose = (const double*)at(d_tp, n * 8); p.tar_vel = (const double*)at(d_tv, n * 8);
                p.joint_active = (const unsigned char*)at(d_ja, N);
                p.tau = (double*)at(d_tau, n * 8);
                p.tau_accumulate = 0;
                p.status = b->d_status + e0;
                int k = trl_launch_imp_pd(b->d_model, b->model->dev, (int)ec, p, b->d_pd_scratch + e0 * sc_stride, sA);
                if (k > 0) b->launches += k; else rc = -k;
            } else {
                cudaMemsetAsync(at(d_tau, n * 8), 0, ec * n * 8, sA);
            }
            ev_rec(b->ev_cmp[slice][2], sA);
        }
        // downloads of this slice, in the order the chains finish
        if (!rc && want_kin) {
            st_wait(s_dn, b->ev_cmp[slice][0]);
            down(a->out_kin_pose, d_kp, n * 8, e0, ec);
            down(a->out_kin_vel, d_kv, n * 8, e0, ec);
            down(a->out_kin_body_pos, d_kbp, N * 24, e0, ec);
            if (want_reward) down(o->out_reward, d_rw, 8, e0, ec);
        }
        if (!rc && want_obs) {
            st_wait(s_dn, b->ev_cmp[slice][1]);
            down(a->out_state, d_st, F * 8, e0, ec);
            if (want_f32) down(o->out_state_f32, d_st32, F * 4, e0, ec);
        }
        if (!rc && want_pd) {
            st_wait(s_dn, b->ev_cmp[slice][2]);
            down(a->out_tau, d_tau, n * 8, e0, ec);
        }
    }
    {   // join everything back into the caller's stream (also on error, so the streams stay consistent)
        int k = 0;
        for (cudaStream_t st : {sB, sC, sA}) {
            ev_rec(b->ev_join[k], st);
            st_wait(b->stream, b->ev_join[k]);
            ++k;
        }
        ev_rec(b->ev_join2[0], s_up0);
        st_wait(b->stream, b->ev_join2[0]);
        ev_rec(b->ev_join2[1], s_dn);
        st_wait(b->stream, b->ev_join2[1]);
        ev_rec(b->ev_join2[2], s_up1);
        st_wait(b->stream, b->ev_join2[2]);
        ev_rec(b->ev_pv2, s_up2);
        st_wait(b->stream, b->ev_pv2);
    }
    if (rc) {
        if (!capturing) cudaStreamSynchronize(b->stream);
        trl_set_error(std::string("trl_step_fused: ") + cudaGetErrorString((cudaError_t)rc));
        return TRL_ERR_CUDA;
    }
    return TRL_OK;  // enqueued; the caller synchronises (or not: asynchronous host mode)
}

// everything a captured HOST-pointer step bakes in besides its arguments: the staging buffers it was handed, the batch's
// device arrays and the settings that shape the pipeline
static void host_graph_snapshot(const trl_batch* b, std::vector<const void*>* out) {
    out->clear();
    for (size_t i = 0; i < b->pool_next && i < b->pool.size(); ++i) out->push_back(b->pool[i].p);
    const void* fixed[] = {b->d_model, b->d_frames, b->d_status, b->d_sample_local, b->d_env_rec, b->d_pd_scratch,
                           b->ground.data, b->ground.pieces, b->ground.env_to_field, b->d_zero_nvec, b->kin_tmp.p, b->state_tmp.p,
                           (const void*)b->stream, (const void*)(size_t)b->host_async, (const void*)(size_t)b->host_slices,
                           (const void*)(size_t)b->ground.kind, (const void*)(size_t)b->ground.num_fields,
                           (const void*)(size_t)b->ground.pieces_per_field, (const void*)(size_t)b->ground.patch_nbuf,
                           (const void*)(size_t)b->ground.patch_w, (const void*)(size_t)b->ground.patch_h,
                           (const void*)(size_t)b->ground.table_ok, (const void*)(size_t)b->overlap};
    for (const void* p : fixed) out->push_back(p);
}

static bool host_args_pinned(const trl_step_args* a, const trl_step_opts* o) {
    const void* ptrs[] = {a->pose, a->vel, a->tar_pose, a->tar_vel, a->joint_active, a->kin_time, a->origin_pos, a->origin_rot,
                          a->body_pos, a->body_vel, a->phase, a->flip, a->out_tau, a->out_kin_pose, a->out_kin_vel,
                          a->out_kin_body_pos, a->out_state, o ? (const void*)o->out_reward : nullptr,
                          o ? (const void*)o->fallen : nullptr, o ? (const void*)o->out_state_f32 : nullptr};
    for (const void* p : ptrs) {
        if (!p) continue;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        if (at.type != cudaMemoryTypeHost) return false;  // pageable memory: its copies cannot be captured
    }
    return true;
}

static int step_fused_host_sliced(trl_batch* b, const trl_step_args* a, const trl_step_opts* o, bool want_pd, bool want_kin,
                                  bool want_obs) {
    static const int graph_env = getenv("TRL_HOST_GRAPH") ? atoi(getenv("TRL_HOST_GRAPH")) : 1;  // A/B hook: 0 = direct submission
    auto finish = [&]() -> int {
        if (b->host_async) return TRL_OK;  // the caller completes the step with trl_batch_sync()
        return cuda_ok(cudaStreamSynchronize(b->stream), "cudaStreamSynchronize") ? TRL_OK : TRL_ERR_CUDA;
    };
    trl_batch::HostGraph* g = nullptr;
    if (graph_env && !b->hg_failed && b->stream != nullptr) {
        for (auto& c : b->hgs)
            if (c.has_o == (o != nullptr) && memcmp(&c.a, a, sizeof(*a)) == 0 && (!o || memcmp(&c.o, o, sizeof(*o)) == 0)) g = &c;
        if (!g) {  // first call with these arguments: remember them (the least recently used set makes room)
            if ((int)b->hgs.size() < trl_batch::kHostGraphs) {
                b->hgs.emplace_back();
                g = &b->hgs.back();
            } else {
                g = &b->hgs[0];
                for (auto& c : b->hgs)
                    if (c.tick < g->tick) g = &c;
                if (g->exec) cudaGraphExecDestroy(g->exec);
                *g = trl_batch::HostGraph();
            }
            g->a = *a;
            g->has_o = o != nullptr;
            if (o) g->o = *o;
        }
        g->tick = ++b->hg_tick;
        g->seen += 1;
    }
    auto drop = [&]() {
        if (g && g->exec) cudaGraphExecDestroy(g->exec);
        if (g) g->exec = nullptr;
    };
    if (g && g->exec) {  // replay: same arguments, and nothing the graph points at has moved
        std::vector<const void*> now;
        b->pool_next = g->pool_used;  // the staging pool hands out the same buffers in the same order
        host_graph_snapshot(b, &now);
        if (now == g->snap && cudaGraphLaunch(g->exec, b->stream) == cudaSuccess) {
            b->launches += g->launches;
            return finish();
        }
        cudaGetLastError();
        drop();
    }
    if (g && g->seen >= 2 && host_args_pinned(a, o)) {  // second (or later) call with these arguments: capture it
        const long long l0 = b->launches;
        if (cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            const int rc = step_fused_host_enqueue(b, a, o, want_pd, want_kin, want_obs, true);
            cudaGraph_t graph = nullptr;
            const cudaError_t ee = cudaStreamEndCapture(b->stream, &graph);
            bool ok = rc == TRL_OK && ee == cudaSuccess && graph != nullptr;
            if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                g->launches = b->launches - l0;
                g->pool_used = b->pool_next;
                host_graph_snapshot(b, &g->snap);
                if (cudaGraphLaunch(g->exec, b->stream) == cudaSuccess) return finish();
            }
            cudaGetLastError();
            drop();
            b->launches = l0;
        } else {
            cudaGetLastError();
        }
        b->hg_failed = true;  // this batch keeps to direct submission
    }
    const int rc = step_fused_host_enqueue(b, a, o, want_pd, want_kin, want_obs, false);
    if (rc != TRL_OK) return rc;
    return finish();
}

extern "C" int trl_batch_set_host_async(trl_batch* b, int on) {
    if (!b) return TRL_ERR_INVALID_ARG;
    b->host_async = on ? 1 : 0;
    return TRL_OK;
}

extern "C" int trl_step_fused(trl_batch* b, const trl_step_args* a) { return trl_step_fused_ex(b, a, nullptr); }

extern "C" int trl_step_fused_ex(trl_batch* b, const trl_step_args* a, const trl_step_opts* o) {
    if (!b || !a || !a->pose || !a->vel) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    const bool want_reward = o && o->out_reward, want_f32 = o && o->out_state_f32;
    const bool want_pd = a->out_tau != nullptr, want_kin = a->out_kin_pose != nullptr || want_reward;
    const bool want_obs = a->out_state != nullptr || want_f32;
    if (want_pd && !a->tar_pose) { trl_set_error("trl_step_fused: out_tau needs tar_pose"); return TRL_ERR_INVALID_ARG; }
    if (want_reward && !a->body_vel) { trl_set_error("trl_step_fused: out_reward needs body_vel"); return TRL_ERR_INVALID_ARG; }
    if (want_kin && (b->model->num_frames < 2 || !a->kin_time)) { trl_set_error("trl_step_fused: kin pose needs a motion and kin_time"); return TRL_ERR_NO_MOTION; }
    if (want_obs && (!a->body_pos || !a->body_vel)) { trl_set_error("trl_step_fused: out_state needs body_pos and body_vel"); return TRL_ERR_INVALID_ARG; }
    if (want_obs && b->cfg.obs_kind == TRL_OBS_CT_3D) { trl_set_error("trl_step_fused: TRL_OBS_CT_3D needs body_quat / body_angvel: use trl_build_poli_state"); return TRL_ERR_INVALID_ARG; }
    if (b->ptr_mode == TRL_PTR_HOST && b->overlap && (b->host_async || (b->host_slices > 1 && E >= 2048)))
        return step_fused_host_sliced(b, a, o, want_pd, want_kin, want_obs);
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    StepPtrs p{};
    p.dt = a->dt;
    p.tau_accumulate = 0;
    p.status = b->d_status;

    // Three independent chains: (A) stable PD, (B) kin-char pose + FK, (C) observation. With overlap on they fork
    // onto the batch's side streams (A on the high-priority one) and join the caller's stream before returning. In HOST pointer mode every chain also
    // stages its own inputs and copies its own outputs back on its stream, so H2D, kernels and D2H of different
    // chains overlap (the observation chain, whose [E][F] state is the largest transfer, goes first).
    int dep_rc = 0;  // first failure of an event / dependency call (reported with the launch status)
    auto ev_rec = [&](cudaEvent_t ev, cudaStream_t st) {
        const cudaError_t err = cudaEventRecord(ev, st);
        if (err != cudaSuccess && !dep_rc) dep_rc = (int)err;
    };
    auto st_wait = [&](cudaStream_t st, cudaEvent_t ev) {
        const cudaError_t err = cudaStreamWaitEvent(st, ev, 0);
        if (err != cudaSuccess && !dep_rc) dep_rc = (int)err;
    };
    const bool fork = b->overlap && ((want_pd ? 1 : 0) + (want_kin ? 1 : 0) + (want_obs ? 1 : 0)) > 1;
    cudaStream_t sA = b->stream, sB = b->stream, sC = b->stream;
    if (fork) {
        if (!cuda_ok(cudaEventRecord(b->ev_fork, b->stream), "cudaEventRecord")) return TRL_ERR_CUDA;
        if (want_kin && (want_pd || want_obs)) { sB = b->side[0]; st_wait(sB, b->ev_fork); }
        if (want_obs && want_pd) { sC = b->side[1]; st_wait(sC, b->ev_fork); }
        if (want_pd) { sA = b->side[2]; st_wait(sA, b->ev_fork); }
    }
    const double *d_time = nullptr, *d_op = nullptr, *d_or = nullptr, *d_bp = nullptr, *d_bv = nullptr, *d_ph = nullptr;
    const unsigned char *d_fl = nullptr, *d_fallen = nullptr;
    double *d_kp = nullptr, *d_kv = nullptr, *d_kbp = nullptr, *d_state = nullptr, *d_reward = nullptr;
    float* d_state32 = nullptr;
    bool ok = true;
    if (want_reward && (!a->out_kin_pose || !a->out_kin_vel)) ok = ok && b->kin_tmp.reserve(sizeof(double) * E * 2 * n);
    if (want_f32 && !a->out_state) ok = ok && b->state_tmp.reserve(sizeof(double) * E * (size_t)b->F);
    // pose / vel are shared by chains A and C: stage them on C's stream (C starts first), A waits for them
    const cudaStream_t s_pv = want_obs ? sC : sA;
    ok = ok && in_ptr(b, a->pose, E * n, &p.pose, s_pv) && in_ptr(b, a->vel, E * n, &p.vel, s_pv);
    if (ok && fork && s_pv != sA && b->ptr_mode == TRL_PTR_HOST) {
        ev_rec(b->ev_pv, s_pv);
        st_wait(sA, b->ev_pv);
    }
    if (want_obs) {
        ok = ok && in_ptr(b, a->body_pos, E * N * 3, &d_bp, sC) && in_ptr(b, a->body_vel, E * N * 3, &d_bv, sC) &&
             in_ptr(b, a->phase, E, &d_ph, sC) && in_ptr(b, a->flip, E, &d_fl, sC) &&
             out_ptr(b, a->out_state, E * (size_t)b->F, &d_state, copies, false, sC) &&
             out_ptr(b, o ? o->out_state_f32 : nullptr, E * (size_t)b->F, &d_state32, copies, false, sC);
        if (!d_state) d_state = static_cast<double*>(b->state_tmp.p);
    } else if (want_reward) {
        ok = ok && in_ptr(b, a->body_vel, E * N * 3, &d_bv, sB);
    }
    if (want_kin) {
        ok = ok && in_ptr(b, a->kin_time, E, &d_time, sB) && in_ptr(b, a->origin_pos, E * 3, &d_op, sB) &&
             in_ptr(b, a->origin_rot, E * 4, &d_or, sB) && out_ptr(b, a->out_kin_pose, E * n, &d_kp, copies, false, sB) &&
             out_ptr(b, a->out_kin_vel, E * n, &d_kv, copies, false, sB) &&
             out_ptr(b, a->out_kin_body_pos, E * N * 3, &d_kbp, copies, false, sB);
        if (want_reward) {  // kin pose / vel the caller did not ask for stay on the device
            if (!d_kp) d_kp = static_cast<double*>(b->kin_tmp.p);
            if (!d_kv) d_kv = static_cast<double*>(b->kin_tmp.p) + E * n;
            ok = ok && in_ptr(b, o->fallen, E, &d_fallen, sB) && out_ptr(b, o->out_reward, E, &d_reward, copies, false, sB);
        }
    }
    if (want_pd) {
        ok = ok && in_ptr(b, a->tar_pose, E * n, &p.tar_pose, sA) && in_ptr(b, a->tar_vel, E * n, &p.tar_vel, sA) &&
             in_ptr(b, a->joint_active, E * N, &p.joint_active, sA) && out_ptr(b, a->out_tau, E * n, &p.tau, copies, false, sA);
        if (!a->tar_vel) p.tar_vel = b->d_zero_nvec;  // the reference's default target velocity
    }
    if (ok && want_reward && fork && b->ptr_mode == TRL_PTR_HOST) {  // the reward reads pose / vel / body_vel staged on other streams
        if (s_pv != sB) { ev_rec(b->ev_pv2, s_pv); st_wait(sB, b->ev_pv2); }
        if (want_obs && sC != sB) { ev_rec(b->ev_bv, sC); st_wait(sB, b->ev_bv); }
    }
    int rc = ok ? 0 : (int)cudaErrorMemoryAllocation;
    auto launch_obs = [&]() {
        if (rc || !want_obs) return;
        rc = trl_launch_poli_state(b->d_model, b->model->dev, b->ground, b->cfg, b->F, b->E, p.pose, p.vel, d_bp, d_bv,
                                   nullptr, nullptr, nullptr, d_ph, d_fl, b->d_sample_local, b->d_env_rec, d_state, nullptr, sC,
                                   b->obs_side, b->ev_obs[0], b->ev_obs[1]);
        b->launches += (b->cfg.num_samples > 0) ? 3 : 2;
        if (!rc && d_state32) {
            rc = trl_launch_f64_to_f32(d_state, d_state32, E * (size_t)b->F, sC);
            b->launches += 1;
        }
    };
    auto launch_kin = [&]() {
        if (rc || !want_kin) return;
        rc = trl_launch_kin(b->d_model, b->model->dev, b->d_frames, b->E, d_time, d_op, d_or, d_kp, d_kv, d_kbp, sB);
        b->launches += 1;
        if (!rc && want_reward) {
            rc = trl_launch_imitate_reward(b->d_model, b->model->dev, b->ground, b->E, p.pose, p.vel, d_kp, d_kv, d_bv, d_fallen,
                                           d_reward, nullptr, sB);
            b->launches += 1;
        }
    };
    auto launch_pd = [&]() {
        if (rc || !want_pd) return;
        if (a->dt > 0) {
            rc = trl_launch_imp_pd(b->d_model, b->model->dev, b->E, p, b->d_pd_scratch, sA);
            if (rc > 0) { b->launches += rc; rc = 0; } else { rc = -rc; }
        } else {
            cudaMemsetAsync(p.tau, 0, sizeof(double) * E * n, sA);
        }
    };
    if (b->ptr_mode == TRL_PTR_HOST) {  // largest D2H (the state) first
        launch_obs(); launch_kin(); launch_pd();
    } else {
        // Submission order of the three chains (block-scheduling priority is per stream, see trl_batch_create). The
        // observation chain goes first: its first kernel gates the step's widest one (k_obs_samples). Measured per step:
        // obs-pd-kin 0.1027 ms, pd-obs-kin 0.1051, obs-kin-pd 0.1045, kin-obs-pd 0.1047.
        launch_obs(); launch_pd(); launch_kin();
    }
    if (!rc && b->ptr_mode == TRL_PTR_HOST) {  // copy every chain's outputs back on its own stream
        for (auto& c : copies)
            if (!cuda_ok(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, c.st ? c.st : b->stream), "D2H")) rc = (int)cudaErrorUnknown;
        copies.clear();
    }
    if (fork) {  // join (also on error, so the streams stay consistent / a capture stays closed)
        if (sB != b->stream) { ev_rec(b->ev_join[0], sB); st_wait(b->stream, b->ev_join[0]); }
        if (sC != b->stream) { ev_rec(b->ev_join[1], sC); st_wait(b->stream, b->ev_join[1]); }
        if (sA != b->stream) { ev_rec(b->ev_join[2], sA); st_wait(b->stream, b->ev_join[2]); }
    }
    return finish(b, copies, rc ? rc : dep_rc);
}

extern "C" int trl_imitate_calc_reward(trl_batch* b, const double* pose, const double* vel, const double* kin_pose,
                                       const double* kin_vel, const double* body_vel, const unsigned char* fallen,
                                       double* out_reward, double* out_terms) {
    if (!b || !pose || !vel || !kin_pose || !kin_vel || !body_vel || !out_reward) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    const size_t E = b->E, n = b->model->n, N = b->model->N;
    b->pool_next = 0;
    std::vector<OutCopy> copies;
    const double *d_p0, *d_v0, *d_p1, *d_v1, *d_bv;
    const unsigned char* d_f;
    double *d_r, *d_t;
    if (!in_ptr(b, pose, E * n, &d_p0) || !in_ptr(b, vel, E * n, &d_v0) || !in_ptr(b, kin_pose, E * n, &d_p1) ||
        !in_ptr(b, kin_vel, E * n, &d_v1) || !in_ptr(b, body_vel, E * N * 3, &d_bv) || !in_ptr(b, fallen, E, &d_f) ||
        !out_ptr(b, out_reward, E, &d_r, copies) || !out_ptr(b, out_terms, E * 5, &d_t, copies))
        return TRL_ERR_CUDA;
    int rc = trl_launch_imitate_reward(b->d_model, b->model->dev, b->ground, b->E, d_p0, d_v0, d_p1, d_v1, d_bv, d_f, d_r,
                                       d_t, b->stream);
    b->launches += 1;
    return finish(b, copies, rc);
}

extern "C" int trl_batch_get_status(trl_batch* b, int* out_status) {
    if (!b || !out_status) return TRL_ERR_INVALID_ARG;
    Guard g(b->device);
    cudaMemcpyKind kind = (b->ptr_mode == TRL_PTR_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (!cuda_ok(cudaMemcpyAsync(out_status, b->d_status, sizeof(int) * (size_t)b->E, kind, b->stream), "status copy")) return TRL_ERR_CUDA;
    if (b->ptr_mode == TRL_PTR_HOST && !cuda_ok(cudaStreamSynchronize(b->stream), "sync")) return TRL_ERR_CUDA;
    return TRL_OK;
}

extern "C" int trl_device_fp64_probe(int device, double* out_tflops) {
    if (!out_tflops || device < 0) return TRL_ERR_INVALID_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= device) {
        cudaGetLastError();
        trl_set_error("trl_device_fp64_probe: no usable CUDA device");
        return TRL_ERR_NO_DEVICE;
    }
    Guard g(device);
    const double tf = trl_fp64_probe_tflops(nullptr);
    if (tf < 0) { trl_set_error("fp64 probe failed"); return TRL_ERR_CUDA; }
    *out_tflops = tf;
    return TRL_OK;
}
