// CUDA kernels (sm_100a) for the batched control-and-observation step.
//
// Layout: one warp per environment. All per-env intermediates live in a shared-memory workspace
// (no global scratch), inputs/outputs are env-major in HBM so the lanes of a warp read/write
// consecutive doubles of one env (coalesced).
//
// Stable PD (k_imp_pd), per env:
//   P1  joint-local transforms (lanes <-> joints)                 cKinTree::ChildParentTrans
//   P2  root-frame transforms by tree level (lanes <-> elements)  cRBDUtil::CalcWorldJointTransforms
//   P3  joint subspace columns in the root frame                  cRBDUtil::BuildJointSubspace*
//   P4  velocities / bias accelerations by level                  cRBDUtil::SolveInvDyna forward pass (qdd = 0)
//   P5  rigid-body inertias (10-parameter form) + body forces     BuildInertiaSpatialMat, SolveInvDyna
//   P6  subtree sums (composite inertias, forces)                 BuildMassMat reverse sweep, SolveInvDyna backward
//   P7  mass matrix H and bias C on the live dofs                 BuildMassMat column walks
//   P8  PD errors, right-hand side, H += dt*Kd                    cImpPDController::CalcControlForces
//   P9  LDL^T factorisation + solve (live dofs only)              Eigen LDLT (zero pivot => 0)
//   P10 tau = Kp e_p + Kd (e_v - dt qdd)
// Everything is expressed in the root-joint frame: there, parent->child transforms are pure additions
// for composite inertias and forces, so the only level-serial work is the 3x4 transform composition
// and a 6-vector recurrence. Numbers agree with the reference's joint-local formulation to ~1e-13.
#include <algorithm>
#include <cstdint>
#include <mutex>
#include <cstdlib>
#include <type_traits>
#include <math_constants.h>

#include "trl_device.cuh"

namespace {

// ---- optional kernel timeline (trl_debug_trace_*): first block start / last block end per kernel, %globaltimer ns ----
__constant__ unsigned long long* g_trace = nullptr;  // [kTraceSlots][2] or null. In the constant bank: the test at the start and end of every block is one cached constant load, not a trip to L2 (k_obs_gather runs 16384 short blocks per step)
constexpr int kTraceSlots = 16;
enum { TR_IMP_PD = 0, TR_PD_TREE, TR_OBS_FEAT, TR_PD_SOLVE, TR_KIN, TR_OBS_ENV, TR_OBS_SAMPLES, TR_FK, TR_KIN_CHAR, TR_REWARD };
struct KTrace {
    int id;
    __device__ __forceinline__ explicit KTrace(int id_) : id(id_) {
        if (g_trace && threadIdx.x == 0) {  // pointer first: tracing off = one uniform constant load and a branch
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMin(&g_trace[2 * id], t);
        }
    }
    __device__ __forceinline__ ~KTrace() {
        if (g_trace && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMax(&g_trace[2 * id + 1], t);
        }
    }
};

constexpr int kWarpsPerBlock = 4;
#ifndef TRL_PD_MIN_BLOCKS
#define TRL_PD_MIN_BLOCKS 4
#endif

constexpr int kCst = 21;  // per-joint double constants staged per block by k_imp_pd
__host__ __device__ inline int pd_cst_doubles(int N) { return (kCst * N + 1) & ~1; }

constexpr int kTile = 32;
// Column c of the joint subspace S in the root frame (cRBDUtil::BuildJointSubspace*, sim/RBDUtil.cpp:733-828),
// recomputed from the joint's root-frame transform wherever it is needed instead of being stored.
TRL_DEV void subspace_col(const ModelIdx* X, const double* T0, const double* Rq, int c, double* s) {
    const int j = X->col_joint[c], kind = X->col_kind[c], a = X->col_axis[c];
    const double* T = T0 + 12 * j;
    if (kind == CK_ANG) {
        s[0] = T[a]; s[1] = T[3 + a]; s[2] = T[6 + a];
        cross3(T + 9, s, s + 3);
    } else if (kind == CK_LIN) {
        s[0] = s[1] = s[2] = 0.0;
        s[3] = T[a]; s[4] = T[3 + a]; s[5] = T[6 + a];
    } else {  // CK_ROOT_LIN: row a of R(q_root)
        s[0] = s[1] = s[2] = 0.0;
        s[3] = Rq[3 * a]; s[4] = Rq[3 * a + 1]; s[5] = Rq[3 * a + 2];
    }
}

// ------------------------------------------------------------------------------------------------
// transforms in the root frame: Tl (local) -> T0 (root frame), level by level
// ------------------------------------------------------------------------------------------------
TRL_DEV void compose_levels(const DevModel* __restrict__ M, const ModelIdx* X, const double* Tl, double* T0, int lane,
                            int first_level, int G = 32) {
    const int nlev = M->num_levels;
    for (int lv = first_level; lv < nlev; ++lv) {
        const int beg = X->level_start[lv];
        const int cnt = X->level_start[lv + 1] - beg;
        for (int it = lane; it < cnt * 12; it += G) {
            const int j = X->level_joint[beg + it / 12];
            const int el = it % 12;
            const int p = X->parent[j];
            const double* Rp = T0 + 12 * p;
            const double* L = Tl + 12 * j;
            double v;
            if (el < 9) {
                const int r = el / 3, c = el % 3;
                v = Rp[r * 3] * L[c] + Rp[r * 3 + 1] * L[3 + c] + Rp[r * 3 + 2] * L[6 + c];
            } else {
                const int r = el - 9;
                v = Rp[9 + r] + (Rp[r * 3] * L[9] + Rp[r * 3 + 1] * L[10] + Rp[r * 3 + 2] * L[11]);
            }
            T0[12 * j + el] = v;
        }
        __syncwarp();
    }
}

TRL_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

TRL_DEV double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

constexpr int kTreeHead = 9;  // per env, ahead of the levels: R(q_root), rows = directions of the root translation dofs

TRL_DEV void tree_subspace(int kind, int a, const double* T /* [(k) * 32] strided */, const double* Rq, double* s) {
    if (kind == CK_ANG) {
        s[0] = T[a * kTile]; s[1] = T[(3 + a) * kTile]; s[2] = T[(6 + a) * kTile];
        const double t0 = T[9 * kTile], t1 = T[10 * kTile], t2 = T[11 * kTile];
        s[3] = t1 * s[2] - t2 * s[1];
        s[4] = t2 * s[0] - t0 * s[2];
        s[5] = t0 * s[1] - t1 * s[0];
    } else if (kind == CK_LIN) {
        s[0] = s[1] = s[2] = 0.0;
        s[3] = T[a * kTile]; s[4] = T[(3 + a) * kTile]; s[5] = T[(6 + a) * kTile];
    } else {  // CK_ROOT_LIN: row a of R(q_root)
        s[0] = s[1] = s[2] = 0.0;
        s[3] = Rq[(3 * a) * kTile]; s[4] = Rq[(3 * a + 1) * kTile]; s[5] = Rq[(3 * a + 2) * kTile];
    }
}
#include "trl_pd_block.cuh"
#include "trl_pd_aba.cuh"

// ------------------------------------------------------------------------------------------------
// k_rbd_extras: cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce (sim/RBDUtil.cpp:256-275,294-345,
// 937-982) -- thread per env, same depth-first walk and per-level workspace as k_pd_tree.
// In the root frame every quantity is a by-product of the subtree sums (mass m, first moment h = sum m c):
//   J column c (world)   = X_{world<-root} S_c = [Rw w ; Rw v + p_root x Rw w]
//   Jcom column c        = [(m_sub / M) w_w ; (m_sub / M) v_w - ((Rw h_sub + m_sub p_root) / M) x w_w]
//   tau_g[c]             = S_c . [h_g x g_r ; m_g g_r],  g_r = Rw^T g
// where the Jcom sums run over every valid body below the joint, while the gravity sums follow the reference's
// quirk: an invalid body gets no torque and does not hand its children's forces to its parent.
// ------------------------------------------------------------------------------------------------
constexpr int kXSlotT = 0;   // root-frame transform [R(9) | t(3)]
constexpr int kXSlotC = 12;  // m, h(3) of the valid bodies of the subtree (Jcom)
constexpr int kXSlotG = 16;  // m, h(3) along the reference's gravity recursion
constexpr int kXSlot = 20;

__global__ void __launch_bounds__(128)
k_rbd_extras(const DevModel* __restrict__ M, int E, const double* __restrict__ pose, double* __restrict__ out_J,
             double* __restrict__ out_Jcom, double* __restrict__ out_gforce) {
    extern __shared__ double smem[];
    __shared__ ModelIdx s_ix;
    stage_model_idx(&s_ix, M);
    const ModelIdx* X = &s_ix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e0 = (blockIdx.x * (blockDim.x >> 5) + warp) * kTile;
    if (e0 >= E) return;
    const int envs = min(kTile, E - e0);
    const bool env_ok = lane < envs;
    const int e = e0 + (env_ok ? lane : envs - 1);
    const int N = M->N, n = M->n;
    double* ws = smem + (size_t)warp * ((kTreeHead + M->num_levels * kXSlot) * kTile) + lane;
    double* Rq = ws;
    double* lvl = ws + kTreeHead * kTile;
    const double* q = pose + (size_t)e * n;
    // J / Jcom rows are assembled in shared memory ([lane][6n], odd row stride: conflict-free) and written out coalesced
    // at the end: a thread-per-env kernel writing its own 1.2 KB row directly would touch 32 sectors per store.
    const int jstride = (6 * n) | 1;
    double* jbuf = smem + (size_t)(blockDim.x >> 5) * ((kTreeHead + M->num_levels * kXSlot) * kTile) +
                   (size_t)warp * 2 * kTile * jstride;
    double* Jw = out_J ? jbuf + lane * jstride : nullptr;
    double* Jc = out_Jcom ? jbuf + (kTile + lane) * jstride : nullptr;
    double* tg = out_gforce ? out_gforce + (size_t)e * n : nullptr;
    for (int i = 0; i < 6 * n; ++i) {  // dead dof slots and the columns of joints without dofs stay zero
        if (Jw) Jw[i] = 0.0;
        if (Jc) Jc[i] = 0.0;
    }
    if (env_ok && tg) for (int i = 0; i < n; ++i) tg[i] = 0.0;
    double Rw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pr[3] = {0, 0, 0}, gr[3] = {0, 0, 0};
    const double inv_total = 1.0 / M->total_mass;
    const int nev = 2 * N;
    for (int evi = 0; evi < nev; ++evi) {
        const int code = X->dfs_event[evi];
        if (code >= 0) {
            const int j = code;
            const int lv = X->level[j];
            double* W = lvl + (size_t)lv * kXSlot * kTile;
            const int o = X->offset[j], sz = X->size[j];
            double T[12];
            if (X->parent[j] == -1) {
                double qr[7];
#pragma unroll
                for (int i = 0; i < 7; ++i) qr[i] = q[o + i];
                double rq9[9];
                quat_to_mat(qr[3], qr[4], qr[5], qr[6], rq9);
                mat3_mul(M->attR[j], rq9, Rw);
                mat3_mulv(M->attR[j], qr, pr);
                pr[0] += M->attT[j][0]; pr[1] += M->attT[j][1]; pr[2] += M->attT[j][2];
#pragma unroll
                for (int i = 0; i < 9; ++i) Rq[i * kTile] = rq9[i];
                const double g[3] = {M->gravity[0], M->gravity[1], M->gravity[2]};
                mat3T_mulv(Rw, g, gr);
#pragma unroll
                for (int i = 0; i < 9; ++i) T[i] = (i % 4 == 0) ? 1.0 : 0.0;
                T[9] = T[10] = T[11] = 0.0;
            } else {
                double qj[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) qj[i] = (i < sz) ? q[o + i] : 0.0;
                double R[9], t[3];
                joint_local_trans(X, j, qj, M->attR[j], M->attT[j], R, t);
                const double* Wp = W - kXSlot * kTile;
                double Tp[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) Tp[i] = Wp[(kXSlotT + i) * kTile];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        T[r * 3 + c] = Tp[r * 3] * R[c] + Tp[r * 3 + 1] * R[3 + c] + Tp[r * 3 + 2] * R[6 + c];
                    T[9 + r] = Tp[9 + r] + (Tp[r * 3] * t[0] + Tp[r * 3 + 1] * t[1] + Tp[r * 3 + 2] * t[2]);
                }
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) W[(kXSlotT + i) * kTile] = T[i];
            double m = 0.0, h[3] = {0, 0, 0};
            if (X->valid_body[j]) {
                m = M->mass[j];
                double c[3];
                mat3_mulv(T, M->com[j], c);
                h[0] = m * (c[0] + T[9]); h[1] = m * (c[1] + T[10]); h[2] = m * (c[2] + T[11]);
            }
            W[kXSlotC * kTile] = m; W[kXSlotG * kTile] = m;
#pragma unroll
            for (int i = 0; i < 3; ++i) { W[(kXSlotC + 1 + i) * kTile] = h[i]; W[(kXSlotG + 1 + i) * kTile] = h[i]; }
        } else {
            const int j = -1 - code;
            const int lv = X->level[j];
            double* W = lvl + (size_t)lv * kXSlot * kTile;
            const bool vb = X->valid_body[j] != 0;
            const double mc = W[kXSlotC * kTile], mg = W[kXSlotG * kTile];
            double hc[3], hg[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { hc[i] = W[(kXSlotC + 1 + i) * kTile]; hg[i] = W[(kXSlotG + 1 + i) * kTile]; }
            // world first moment of the subtree about the world origin, over the total mass
            double hw[3];
            mat3_mulv(Rw, hc, hw);
            const double mf = mc * inv_total;
            const double pc[3] = {(hw[0] + mc * pr[0]) * inv_total, (hw[1] + mc * pr[1]) * inv_total, (hw[2] + mc * pr[2]) * inv_total};
            // gravity wrench of the subtree in the root frame: [h x g ; m g]
            double ng[3];
            cross3(hg, gr, ng);
            const int c0 = X->first_col[j], nc = X->ncol[j];
            for (int c = c0; c < c0 + nc; ++c) {
                double sv[6];
                tree_subspace(X->col_kind[c], X->col_axis[c], W + kXSlotT * kTile, Rq, sv);
                double ww[3], vr[3], vw[3], cx[3];
                mat3_mulv(Rw, sv, ww);
                mat3_mulv(Rw, sv + 3, vr);
                cross3(pr, ww, cx);
                vw[0] = vr[0] + cx[0]; vw[1] = vr[1] + cx[1]; vw[2] = vr[2] + cx[2];
                const int di = X->col_dof[c];
                if (Jw) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) { Jw[r * n + di] = ww[r]; Jw[(3 + r) * n + di] = vw[r]; }
                }
                if (Jc) {
                    double px[3];
                    cross3(pc, ww, px);
#pragma unroll
                    for (int r = 0; r < 3; ++r) { Jc[r * n + di] = mf * ww[r]; Jc[(3 + r) * n + di] = mf * vw[r] - px[r]; }
                }
                if (tg && env_ok)
                    tg[di] = vb ? ((sv[0] * ng[0] + sv[1] * ng[1] + sv[2] * ng[2]) +
                                   mg * (sv[3] * gr[0] + sv[4] * gr[1] + sv[5] * gr[2])) : 0.0;
            }
            if (lv > 0) {
                double* Wp = W - kXSlot * kTile;
#pragma unroll
                for (int i = 0; i < 4; ++i) Wp[(kXSlotC + i) * kTile] += W[(kXSlotC + i) * kTile];
                if (vb) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) Wp[(kXSlotG + i) * kTile] += W[(kXSlotG + i) * kTile];
                }
            }
        }
    }
    __syncwarp();
    for (int r = 0; r < envs; ++r) {  // coalesced copy-out, one env row at a time
        if (out_J) {
            double* o = out_J + (size_t)(e0 + r) * 6 * n;
            const double* src = jbuf + r * jstride;
            for (int i = lane; i < 6 * n; i += 32) o[i] = src[i];
        }
        if (out_Jcom) {
            double* o = out_Jcom + (size_t)(e0 + r) * 6 * n;
            const double* src = jbuf + (kTile + r) * jstride;
            for (int i = lane; i < 6 * n; i += 32) o[i] = src[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_apply_tau: cSimCharacter::ApplyControlForces + cJoint::ApplyTau (sim/SimCharacter.cpp:589-603, sim/Joint.cpp:
// 162-221,300-318,536-562) -- thread per env, depth-first walk with the world transforms of the
// current root path in shared memory. Generalized torques -> per-joint (torque, force), clamped by TorqueLim /
// ForceLim, rotated to world coordinates by the joint's world frame and applied +/- to the child / parent body.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_apply_tau(const DevModel* __restrict__ M, int E, const double* __restrict__ pose, const double* __restrict__ tau,
            double* __restrict__ out_tau, double* __restrict__ out_wrench) {
    extern __shared__ double smem[];
    __shared__ ModelIdx s_ix;
    stage_model_idx(&s_ix, M);
    const ModelIdx* X = &s_ix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e0 = (blockIdx.x * (blockDim.x >> 5) + warp) * kTile;
    if (e0 >= E) return;
    const int envs = min(kTile, E - e0);
    const bool env_ok = lane < envs;
    const int e = e0 + (env_ok ? lane : envs - 1);
    const int N = M->N, n = M->n;
    double* lvl = smem + (size_t)warp * (M->num_levels * 12 * kTile) + lane;
    const double* q = pose + (size_t)e * n;
    const double* te = tau + (size_t)e * n;
    double* ot = out_tau ? out_tau + (size_t)e * n : nullptr;
    double* ow = out_wrench ? out_wrench + (size_t)e * N * 6 : nullptr;
    if (env_ok) {
        if (ot) for (int i = 0; i < n; ++i) ot[i] = te[i];
        if (ow) for (int i = 0; i < 6 * N; ++i) ow[i] = 0.0;
    }
    const int nev = 2 * N;
    for (int evi = 0; evi < nev; ++evi) {
        const int j = X->dfs_event[evi];
        if (j < 0) continue;
        const int o = X->offset[j], sz = X->size[j], p = X->parent[j];
        double* W = lvl + (size_t)X->level[j] * 12 * kTile;
        double qj[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) qj[i] = (i < sz) ? q[o + i] : 0.0;
        double R[9], t[3], T[12];
        joint_local_trans(X, j, qj, M->attR[j], M->attT[j], R, t);
        if (p == -1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) T[i] = R[i];
            T[9] = t[0]; T[10] = t[1]; T[11] = t[2];
        } else {
            const double* Wp = W - 12 * kTile;
            double Tp[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) Tp[i] = Wp[i * kTile];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) T[r * 3 + c] = Tp[r * 3] * R[c] + Tp[r * 3 + 1] * R[3 + c] + Tp[r * 3 + 2] * R[6 + c];
                T[9 + r] = Tp[9 + r] + (Tp[r * 3] * t[0] + Tp[r * 3 + 1] * t[1] + Tp[r * 3 + 2] * t[2]);
            }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) W[i * kTile] = T[i];
        if (p == -1 || !X->valid_body[j]) continue;  // cJoint::IsValid: a constraint with a child body
        const int type = X->type[j];
        double tot[6] = {0, 0, 0, 0, 0, 0};  // mTotalTau = [torque ; force]
        bool has_torque = false, has_force = false;
        if (type == JT_REVOLUTE) { tot[2] = te[o]; has_torque = true; }
        else if (type == JT_PLANAR) { tot[3] = te[o]; tot[4] = te[o + 1]; tot[0] = te[o + 2]; has_torque = has_force = true; }  // Joint.cpp:541-546
        else if (type == JT_PRISMATIC) { tot[3] = te[o]; has_force = true; }
        else if (type == JT_SPHERICAL) { tot[0] = te[o]; tot[1] = te[o + 1]; tot[2] = te[o + 2]; has_torque = true; }
        if (has_torque) {
            const double mag = sqrt(tot[0] * tot[0] + tot[1] * tot[1] + tot[2] * tot[2]);
            const double lim = M->torque_lim[j];
            if (mag > lim) { const double sc = lim / mag; tot[0] *= sc; tot[1] *= sc; tot[2] *= sc; }
            if (ow && env_ok) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double w = T[r * 3] * tot[0] + T[r * 3 + 1] * tot[1] + T[r * 3 + 2] * tot[2];
                    ow[6 * p + r] += -w;
                    ow[6 * j + r] += w;
                }
            }
        }
        if (has_force) {
            const double mag = sqrt(tot[3] * tot[3] + tot[4] * tot[4] + tot[5] * tot[5]);
            const double lim = M->force_lim[j];
            if (mag > lim) { const double sc = lim / mag; tot[3] *= sc; tot[4] *= sc; tot[5] *= sc; }
            if (ow && env_ok) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double w = T[r * 3] * tot[3] + T[r * 3 + 1] * tot[4] + T[r * 3 + 2] * tot[5];
                    ow[6 * p + 3 + r] += -w;
                    ow[6 * j + 3 + r] += w;
                }
            }
        }
        if (ot && env_ok) {
            if (type == JT_REVOLUTE) ot[o] = tot[2];
            else if (type == JT_PLANAR) { ot[o] = tot[3]; ot[o + 1] = tot[4]; ot[o + 2] = tot[0]; }
            else if (type == JT_PRISMATIC) ot[o] = tot[3];
            else if (type == JT_SPHERICAL) { ot[o] = tot[0]; ot[o + 1] = tot[1]; ot[o + 2] = tot[2]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_fk: cKinTree::JointWorldTrans for all joints + CalcBodyPartPos -- one warp per env
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_fk(const DevModel* __restrict__ M, int E, const double* __restrict__ pose, double* __restrict__ out_joint_world,
     double* __restrict__ out_body_pos) {
    extern __shared__ double smem[];
    __shared__ ModelIdx s_ix;
    stage_model_idx(&s_ix, M);
    const ModelIdx* X = &s_ix;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int e = blockIdx.x * kWarpsPerBlock + warp;
    if (e >= E) return;
    const int N = M->N, n = M->n;
    double* ws = smem + (size_t)warp * (n + 24 * N);
    double* q = ws;
    double* Tl = ws + n;
    double* Tw = Tl + 12 * N;
    for (int i = lane; i < n; i += 32) q[i] = pose[(size_t)e * n + i];
    __syncwarp();
    if (lane < N) {
        double R[9], t[3];
        joint_local_trans(X, lane, q + X->offset[lane], M->attR[lane], M->attT[lane], R, t);
        double* dst = (lane == 0) ? Tw : (Tl + 12 * lane);
#pragma unroll
        for (int i = 0; i < 9; ++i) dst[i] = R[i];
        dst[9] = t[0]; dst[10] = t[1]; dst[11] = t[2];
    }
    __syncwarp();
    compose_levels(M, X, Tl, Tw, lane, 1);
    if (out_joint_world) {
        double* o = out_joint_world + (size_t)e * N * 12;
        for (int idx = lane; idx < N * 12; idx += 32) {
            const int j = idx / 12, el = idx % 12, r = el / 4, c = el % 4;
            o[idx] = (c < 3) ? Tw[12 * j + r * 3 + c] : Tw[12 * j + 9 + r];
        }
    }
    if (out_body_pos) {
        double* o = out_body_pos + (size_t)e * N * 3;
        for (int idx = lane; idx < N * 3; idx += 32) {
            const int j = idx / 3, r = idx % 3;
            const double* T = Tw + 12 * j;
            const double* c = M->com[j];
            o[idx] = X->valid_body[j] ? (T[r * 3] * c[0] + T[r * 3 + 1] * c[1] + T[r * 3 + 2] * c[2] + T[9 + r]) : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_kin_char: cKinCharacter::Pose(t) -- one thread per (env, joint)
// ------------------------------------------------------------------------------------------------
TRL_DEV void motion_index_phase(const DevModel* __restrict__ M, const double* __restrict__ frames, int fs, double time,
                                int* idx, double* phase) {  // cMotion::CalcIndexPhase, anim/Motion.cpp:240-273
    const double max_time = M->motion_duration;
    const int nf = M->num_frames;
    if (!M->loop) {
        if (time <= 0) { *idx = 0; *phase = 0; return; }
        if (time >= max_time) { *idx = nf - 2; *phase = 1; return; }
    }
    time = fmod(time, max_time);
    if (time < 0) time += max_time;
    int lo = 0, hi = nf;  // upper_bound
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (!(time < __ldg(frames + (size_t)mid * fs))) lo = mid + 1; else hi = mid;
    }
    const int i = lo - 1;
    const double t0 = __ldg(frames + (size_t)i * fs), t1 = __ldg(frames + (size_t)(i + 1) * fs);
    *idx = i;
    *phase = (time - t0) / (t1 - t0);
}

// pose of one joint at `time` (root: with cycle offset and origin transform, anim/KinCharacter.cpp:280-315)
TRL_DEV void kin_joint_pose_at(const DevModel* __restrict__ M, const ModelIdx* X, const double* __restrict__ frames,
                               int fs, int j, double time, int idx, double lerp, const double* opos, const double* orot,
                               double* out /* size[j] */) {
    const int o = X->offset[j], sz = X->size[j];
    const double* f0 = frames + (size_t)idx * fs + 1 + o;
    const double* f1 = frames + (size_t)(idx + 1) * fs + 1 + o;
    const bool is_root = (j == 0);
    const bool has_quat = is_root || X->type[j] == JT_SPHERICAL;
    const int qo = is_root ? 3 : 0;  // where the joint's quaternion sits in its parameter block
    double qr[4] = {1, 0, 0, 0};
    if (has_quat) {  // one slerp call site for the root and the spherical joints (keeps the code small)
        double qa[4], qb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qa[i] = __ldg(f0 + qo + i); qb[i] = __ldg(f1 + qo + i); }
        quat_slerp_norm(qa, lerp, qb, qr);
    }
    if (is_root) {
        double p[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = (1 - lerp) * __ldg(f0 + i) + lerp * __ldg(f1 + i);
        double delta[3] = {0, 0, 0};
        if (M->loop) {  // cMotion::CalcCycleCount, Motion.cpp:231-238
            const int cyc = (int)floor(time / M->motion_duration);
#pragma unroll
            for (int i = 0; i < 3; ++i) delta[i] = cyc * M->cycle_delta[i];
        }
        const double ident[4] = {1, 0, 0, 0};
        double rdr[4], rr[4], rp[3];
        quat_mul(orot, ident, rdr);
        quat_mul(rdr, qr, rr);
        quat_rot_vec(rdr, p, rp);
#pragma unroll
        for (int i = 0; i < 3; ++i) out[i] = rp[i] + (delta[i] + opos[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) out[3 + i] = rr[i];
    } else if (has_quat) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = qr[i];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < sz) out[i] = (1 - lerp) * __ldg(f0 + i) + lerp * __ldg(f1 + i);
    }
}

TRL_DEV void kin_joint_pose(const DevModel* __restrict__ M, const ModelIdx* X, const double* __restrict__ frames, int fs,
                            int j, double time, const double* opos, const double* orot, double* out /* size[j] */) {
    int idx;
    double ph;
    motion_index_phase(M, frames, fs, time, &idx, &ph);
    kin_joint_pose_at(M, X, frames, fs, j, time, idx, clampd(ph, 0.0, 1.0), opos, orot, out);
}

__global__ void __launch_bounds__(128)
k_kin_char(const DevModel* __restrict__ M, const double* __restrict__ frames, int E, const double* __restrict__ time,
           const double* __restrict__ origin_pos, const double* __restrict__ origin_rot, double* __restrict__ out_pose,
           double* __restrict__ out_vel) {
    __shared__ ModelIdx s_ix;
    stage_model_idx(&s_ix, M);
    const ModelIdx* X = &s_ix;
    const int N = M->N, n = M->n;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)E * N) return;
    const int e = (int)(tid / N), j = (int)(tid % N);
    const int fs = n + 1;
    const double t = time[e];
    double opos[3] = {0, 0, 0}, orot[4] = {1, 0, 0, 0};
    if (j == 0) {
        if (origin_pos) { opos[0] = origin_pos[3 * (size_t)e]; opos[1] = origin_pos[3 * (size_t)e + 1]; opos[2] = origin_pos[3 * (size_t)e + 2]; }
        if (origin_rot) { orot[0] = origin_rot[4 * (size_t)e]; orot[1] = origin_rot[4 * (size_t)e + 1]; orot[2] = origin_rot[4 * (size_t)e + 2]; orot[3] = origin_rot[4 * (size_t)e + 3]; }
    }
    double p0[7], p1[7];
    kin_joint_pose(M, X, frames, fs, j, t, opos, orot, p0);
    const int o = X->offset[j], sz = X->size[j];
    for (int i = 0; i < sz; ++i) out_pose[(size_t)e * n + o + i] = p0[i];
    if (out_vel) {
        const double ddt = 1 / 60.0;  // gDiffTimeStep, anim/KinCharacter.cpp:5
        kin_joint_pose(M, X, frames, fs, j, t + ddt, opos, orot, p1);
        double v[7];
        if (j == 0) {  // cKinTree::CalcVel, anim/KinTree.cpp:1470-1508
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = (p1[i] - p0[i]) / ddt;
            quat_vel_rel(p0 + 3, p1 + 3, ddt, v + 3);
        } else if (X->type[j] == JT_SPHERICAL) {
            quat_vel_rel(p0, p1, ddt, v);
        } else {
            for (int i = 0; i < sz; ++i) v[i] = (p1[i] - p0[i]) / ddt;
        }
        for (int i = 0; i < sz; ++i) out_vel[(size_t)e * n + o + i] = v[i];
    }
}

// ------------------------------------------------------------------------------------------------
// k_kin_tile: the kinematic chain with one WARP PER JOINT over a tile of 32 envs (lane = env, block = N warps). Every
// branch depends on the joint only, so each warp runs one uniform code path; the joints' poses / velocities are
// independent and run in parallel; FK then composes each joint's world transform up its (short) ancestor chain from
// the local transforms left in shared memory ([joint][12][32 lanes]). N x the parallelism of k_kin_thread at the same
// instruction count.
// ------------------------------------------------------------------------------------------------
constexpr int kKinWarps = 16;  // joints beyond this many are looped over
__global__ void __launch_bounds__(kKinWarps * 32)
k_kin_tile(const DevModel* __restrict__ M, const __grid_constant__ ModelIdx ixp, const double* __restrict__ frames, int E,
           const double* __restrict__ time, const double* __restrict__ origin_pos, const double* __restrict__ origin_rot,
           double* __restrict__ out_pose, double* __restrict__ out_vel, double* __restrict__ out_body_pos) {
    KTrace trace_(TR_KIN);
    extern __shared__ double smem[];  // local transforms [N][12][32]
    const ModelIdx* X = &ixp;  // index tables in the constant bank (kernel parameter): every lookup is warp-uniform
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int e0 = blockIdx.x * kTile;
    const int envs = min(kTile, E - e0);
    const bool env_ok = lane < envs;
    const int e = e0 + (env_ok ? lane : envs - 1);
    const int N = M->N, n = M->n, fs = n + 1;
    double* Tl = smem + lane;  // element k of joint a: Tl[(a * 12 + k) * 32]
    const double t = time[e];
    const double ddt = 1 / 60.0;  // gDiffTimeStep, anim/KinCharacter.cpp:5
    const double inv_ddt = 1.0 / ddt;  // the reference divides by ddt; one reciprocal + products agree to an ulp
    // frame lookup (an exact fmod + a binary search) once per env and time, by the first two warps, shared through
    // shared memory -- not once per joint
    __shared__ double s_lerp[2][kTile];
    __shared__ int s_idx[2][kTile];
    for (int k = warp; k < (out_vel ? 2 : 1); k += nwarps) {
        int idx;
        double ph;
        motion_index_phase(M, frames, fs, k ? t + ddt : t, &idx, &ph);
        s_idx[k][lane] = idx;
        s_lerp[k][lane] = clampd(ph, 0.0, 1.0);
    }
    __syncthreads();
    const int idx0 = s_idx[0][lane], idx1 = out_vel ? s_idx[1][lane] : 0;
    const double lerp0 = s_lerp[0][lane], lerp1 = out_vel ? s_lerp[1][lane] : 0.0;
    for (int j = warp; j < N; j += nwarps) {  // warp <-> joint
        const bool is_root = X->parent[j] == -1;
        double opos[3] = {0, 0, 0}, orot[4] = {1, 0, 0, 0};
        if (is_root) {
            if (origin_pos) { opos[0] = origin_pos[3 * (size_t)e]; opos[1] = origin_pos[3 * (size_t)e + 1]; opos[2] = origin_pos[3 * (size_t)e + 2]; }
            if (origin_rot) { orot[0] = origin_rot[4 * (size_t)e]; orot[1] = origin_rot[4 * (size_t)e + 1]; orot[2] = origin_rot[4 * (size_t)e + 2]; orot[3] = origin_rot[4 * (size_t)e + 3]; }
        }
        const int o = X->offset[j], sz = X->size[j];
        const bool sph = !is_root && X->type[j] == JT_SPHERICAL;
        double p0[7] = {0, 0, 0, 0, 0, 0, 0};
        kin_joint_pose_at(M, X, frames, fs, j, t, idx0, lerp0, opos, orot, p0);
        if (env_ok) {
#pragma unroll
            for (int i = 0; i < 7; ++i)
                if (i < sz) out_pose[(size_t)e * n + o + i] = p0[i];
        }
        if (out_vel) {
            double p1[7] = {0, 0, 0, 0, 0, 0, 0}, v[7] = {0, 0, 0, 0, 0, 0, 0};
            kin_joint_pose_at(M, X, frames, fs, j, t + ddt, idx1, lerp1, opos, orot, p1);
            if (is_root) {  // cKinTree::CalcVel, anim/KinTree.cpp:1470-1508
#pragma unroll
                for (int i = 0; i < 3; ++i) v[i] = (p1[i] - p0[i]) * inv_ddt;
                quat_vel_rel(p0 + 3, p1 + 3, ddt, v + 3);
            } else if (sph) {
                quat_vel_rel(p0, p1, ddt, v);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = (p1[i] - p0[i]) * inv_ddt;
            }
            if (env_ok) {
#pragma unroll
                for (int i = 0; i < 7; ++i)
                    if (i < sz) out_vel[(size_t)e * n + o + i] = v[i];
            }
        }
        if (out_body_pos) {
            double R[9], tt[3];
            joint_local_trans(X, j, p0, M->attR[j], M->attT[j], R, tt);
#pragma unroll
            for (int i = 0; i < 9; ++i) Tl[(j * 12 + i) * kTile] = R[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) Tl[(j * 12 + 9 + i) * kTile] = tt[i];
        }
    }
    if (!out_body_pos) return;
    __syncthreads();
    for (int j = warp; j < N; j += nwarps) {
        double T[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = Tl[(j * 12 + i) * kTile];
        for (int a = X->parent[j]; a != -1; a = X->parent[a]) {  // T <- Tl[a] o T, up to the root
            double P[12], Tn[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) P[i] = Tl[(a * 12 + i) * kTile];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) Tn[r * 3 + c] = P[r * 3] * T[c] + P[r * 3 + 1] * T[3 + c] + P[r * 3 + 2] * T[6 + c];
                Tn[9 + r] = P[9 + r] + (P[r * 3] * T[9] + P[r * 3 + 1] * T[10] + P[r * 3 + 2] * T[11]);
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = Tn[i];
        }
        if (env_ok) {
            const double* c = M->com[j];
            const bool vb = X->valid_body[j] != 0;
            double* ob = out_body_pos + ((size_t)e * N + j) * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r) ob[r] = vb ? (T[r * 3] * c[0] + T[r * 3 + 1] * c[1] + T[r * 3 + 2] * c[2] + T[9 + r]) : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_imitate_reward<G>: cScenarioExpImitate::CalcReward (scenarios/ScenarioExpImitate.cpp:13-159) -- G lanes per env.
// pose0/vel0 = simulated character, pose1/vel1 = kinematic reference. One FK per character in world coordinates;
// the heading-normalised frames of cKinTree::NormalizePoseHeading are applied to the FK results instead of
// re-running FK on normalised pose vectors (same numbers to rounding).
// ------------------------------------------------------------------------------------------------
TRL_DEV double group_sum(double v, int G) {
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// heading frame of a root pose: R_h = R(y, -heading) as (c, s) with R_h * p = (c px + s pz, py, -s px + c pz),
// and the heading quaternion hq = AxisAngleToQuaternion(y, -heading)
struct HeadingFrame { double c, s, hq[4]; };
TRL_DEV HeadingFrame heading_frame(const double* q) {
    HeadingFrame h;
    const double xdir[3] = {1, 0, 0};
    double rd[3];
    quat_rot_vec(q + 3, xdir, rd);
    const double heading = atan2(-rd[2], rd[0]);
    sincos(-heading, &h.s, &h.c);
    double sh, ch;
    sincos(-heading / 2, &sh, &ch);
    h.hq[0] = ch; h.hq[1] = sh * 0.0; h.hq[2] = sh * 1.0; h.hq[3] = sh * 0.0;
    return h;
}

template <int G>
__global__ void __launch_bounds__(128)
k_imitate_reward(const DevModel* __restrict__ M, DevGround g, int E, const double* __restrict__ pose0,
                 const double* __restrict__ vel0, const double* __restrict__ pose1, const double* __restrict__ vel1,
                 const double* __restrict__ body_vel, const unsigned char* __restrict__ fallen,
                 double* __restrict__ out_reward, double* __restrict__ out_terms) {
    extern __shared__ double smem[];
    __shared__ ModelIdx s_ix;
    stage_model_idx(&s_ix, M);
    const ModelIdx* X = &s_ix;
    constexpr int EPW = 32 / G;
    const int wlane = threadIdx.x & 31;
    const int lane = wlane % G, sub = wlane / G;
    const int warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int e_raw = (blockIdx.x * wpb + warp) * EPW + sub;
    if ((blockIdx.x * wpb + warp) * EPW >= E) return;
    const bool env_ok = e_raw < E;
    const int e = env_ok ? e_raw : E - 1;
    const int N = M->N, n = M->n;
    double* ws = smem + (size_t)(warp * EPW + sub) * (42 * N);
    double* Tl = ws;             // 12N scratch for the level composition
    double* Tw0 = Tl + 12 * N;   // world transforms, simulated character
    double* Tw1 = Tw0 + 12 * N;  // world transforms, kinematic character
    double* V1 = Tw1 + 12 * N;   // world spatial velocities (at the world origin), kinematic character
    const double* q0 = pose0 + (size_t)e * n;
    const double* qd0 = vel0 + (size_t)e * n;
    const double* q1 = pose1 + (size_t)e * n;
    const double* qd1 = vel1 + (size_t)e * n;

    // FK of both characters (world frame)
    for (int pass = 0; pass < 2; ++pass) {
        const double* q = pass ? q1 : q0;
        double* Tw = pass ? Tw1 : Tw0;
        for (int j = lane; j < N; j += G) {
            double R[9], t[3];
            double qq[7];
            const int o = X->offset[j], sz = X->size[j];
            for (int i = 0; i < sz; ++i) qq[i] = q[o + i];
            joint_local_trans(X, j, qq, M->attR[j], M->attT[j], R, t);
            double* dst = (j == 0) ? Tw : (Tl + 12 * j);
#pragma unroll
            for (int i = 0; i < 9; ++i) dst[i] = R[i];
            dst[9] = t[0]; dst[10] = t[1]; dst[11] = t[2];
        }
        __syncwarp();
        compose_levels(M, X, Tl, Tw, lane, 1, G);
    }
    // world spatial velocity of every joint of the kinematic character: V_j = V_p + S_j qd_j (motion vectors at the
    // world origin: omega, v_O). Root: omega_l = qd[3:6], v_l = R(q)^T qd[0:3] in the joint frame (RBDUtil.cpp:770-784).
    if (lane == 0) {
        double Rq[9], vl[3], wl[3] = {qd1[3], qd1[4], qd1[5]}, lin[3] = {qd1[0], qd1[1], qd1[2]};
        quat_to_mat(q1[3], q1[4], q1[5], q1[6], Rq);
        mat3T_mulv(Rq, lin, vl);
        double ww[3], vw[3], pxw[3];
        mat3_mulv(Tw1, wl, ww);
        mat3_mulv(Tw1, vl, vw);
        cross3(Tw1 + 9, ww, pxw);
        V1[0] = ww[0]; V1[1] = ww[1]; V1[2] = ww[2];
        V1[3] = vw[0] + pxw[0]; V1[4] = vw[1] + pxw[1]; V1[5] = vw[2] + pxw[2];
    }
    __syncwarp();
    for (int lv = 1; lv < M->num_levels; ++lv) {
        const int beg = X->level_start[lv], cnt = X->level_start[lv + 1] - beg;
        for (int it = lane; it < cnt; it += G) {
            const int j = X->level_joint[beg + it];
            double s[6] = {0, 0, 0, 0, 0, 0};
            const int c0 = X->first_col[j], nc = X->ncol[j];
            for (int c = c0; c < c0 + nc; ++c) {
                double sc6[6];
                subspace_col(X, Tw1, nullptr, c, sc6);  // non-root columns never read Rq
                const double w = qd1[X->col_dof[c]];
#pragma unroll
                for (int k = 0; k < 6; ++k) s[k] += sc6[k] * w;
            }
            const int pj = X->parent[j];
#pragma unroll
            for (int k = 0; k < 6; ++k) V1[6 * j + k] = V1[6 * pj + k] + s[k];
        }
        __syncwarp();
    }

    const HeadingFrame h0 = heading_frame(q0), h1 = heading_frame(q1);
    // per-joint contributions
    double s_m0[3] = {0, 0, 0}, s_m1[3] = {0, 0, 0}, s_v0[3] = {0, 0, 0}, s_v1[3] = {0, 0, 0};
    double pose_err = 0, vel_err = 0, ee_err = 0, n_ee = 0;
    const bool has_ground = g.kind != 0;
    const DevPiece* pcs = nullptr;
    if (g.kind == 1 || g.kind == 2) pcs = g.pieces + (long long)ground_field_of_env(g, e) * g.pieces_per_field;
    for (int j = lane; j < N; j += G) {
        const double w = M->joint_w[j];
        const int o = X->offset[j], sz = X->size[j];
        if (X->valid_body[j]) {
            const double m = M->mass[j];
            const double* cm = M->com[j];
            // body CoM in the heading-normalised frames: R_h (c - root_xz) (y keeps the world height)
            const double* T0 = Tw0 + 12 * j;
            const double* T1 = Tw1 + 12 * j;
            double c0w[3], c1w[3];
            mat3_mulv(T0, cm, c0w); c0w[0] += T0[9]; c0w[1] += T0[10]; c0w[2] += T0[11];
            mat3_mulv(T1, cm, c1w); c1w[0] += T1[9]; c1w[1] += T1[10]; c1w[2] += T1[11];
            const double d0x = c0w[0] - q0[0], d0z = c0w[2] - q0[2];
            const double d1x = c1w[0] - q1[0], d1z = c1w[2] - q1[2];
            s_m0[0] += m * (h0.c * d0x + h0.s * d0z); s_m0[1] += m * c0w[1]; s_m0[2] += m * (-h0.s * d0x + h0.c * d0z);
            s_m1[0] += m * (h1.c * d1x + h1.s * d1z); s_m1[1] += m * c1w[1]; s_m1[2] += m * (-h1.s * d1x + h1.c * d1z);
            // CoM velocities (world): simulated = body velocities, kinematic = v_O + omega x c
            const double* bv = body_vel + ((size_t)e * N + j) * 3;
            s_v0[0] += m * bv[0]; s_v0[1] += m * bv[1]; s_v0[2] += m * bv[2];
            const double* Vj = V1 + 6 * j;
            double wxc[3];
            cross3(Vj, c1w, wxc);
            s_v1[0] += m * (Vj[3] + wxc[0]); s_v1[1] += m * (Vj[4] + wxc[1]); s_v1[2] += m * (Vj[5] + wxc[2]);
        }
        if (j == 0) {
            // root: CalcRootRotErr / CalcRootAngVelErr on the heading-normalised poses (KinTree.cpp:1358-1363,1422-1426)
            double a[4], b[4], bc[4], d[4];
            quat_mul(h0.hq, q0 + 3, a);
            quat_mul(h1.hq, q1 + 3, b);
            bc[0] = a[0]; bc[1] = -a[1]; bc[2] = -a[2]; bc[3] = -a[3];
            quat_mul(b, bc, d);  // QuatDiff(q0, q1) = q1 * conj(q0)
            if (fabs(d[0]) > 1) quat_normalize(d);
            const double th = 2 * acos(d[0]);
            pose_err += w * (th * th);
            double w0r[3], w1r[3];
            quat_rot_vec(h0.hq, qd0 + 3, w0r);
            quat_rot_vec(h1.hq, qd1 + 3, w1r);
            double sv = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) { const double dv = w1r[i] - w0r[i]; sv += dv * dv; }
            vel_err += w * sv;
        } else {
            double cpe = 0, cve = 0;
            if (X->type[j] == JT_SPHERICAL) {
                double bc[4] = {q0[o], -q0[o + 1], -q0[o + 2], -q0[o + 3]}, d[4];
                quat_mul(q1 + o, bc, d);
                if (fabs(d[0]) > 1) quat_normalize(d);
                const double th = 2 * acos(d[0]);
                cpe = th * th;
            } else {
                for (int i = 0; i < sz; ++i) { const double dv = q1[o + i] - q0[o + i]; cpe += dv * dv; }
            }
            for (int i = 0; i < sz; ++i) { const double dv = qd1[o + i] - qd0[o + i]; cve += dv * dv; }
            pose_err += w * cpe;
            vel_err += w * cve;
        }
    }
    // reduce the CoM sums over the env's lanes (every lane gets the totals)
    const double inv_m = 1.0 / M->total_mass;
    double com0[3], com1[3], cv0[3], cv1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        com0[i] = group_sum(s_m0[i], G) * inv_m;
        com1[i] = group_sum(s_m1[i], G) * inv_m;
        cv0[i] = group_sum(s_v0[i], G) * inv_m;
        cv1[i] = group_sum(s_v1[i], G) * inv_m;
    }
    // end effectors (ScenarioExpImitate.cpp:107-128)
    for (int j = 1 + lane; j < N; j += G) {
        if (!X->end_eff[j]) continue;
        const double* T0 = Tw0 + 12 * j;
        const double* T1 = Tw1 + 12 * j;
        double gh0 = 0.0;
        if (has_ground) {
            int v, c3[3];
            gh0 = ground_sample_pieces(g.kind, g.pieces_per_field, pcs, g.data, T0[9], T0[11], &v, c3);
        }
        const double d0x = T0[9] - q0[0], d0z = T0[11] - q0[2];
        const double d1x = T1[9] - q1[0], d1z = T1[11] - q1[2];
        const double r0x = (h0.c * d0x + h0.s * d0z) - com0[0], r0y = T0[10] - gh0, r0z = (-h0.s * d0x + h0.c * d0z) - com0[2];
        const double r1x = (h1.c * d1x + h1.s * d1z) - com1[0], r1y = T1[10] - 0.0, r1z = (-h1.s * d1x + h1.c * d1z) - com1[2];
        ee_err += (r1x - r0x) * (r1x - r0x) + (r1y - r0y) * (r1y - r0y) + (r1z - r0z) * (r1z - r0z);
        n_ee += 1.0;
    }
    pose_err = group_sum(pose_err, G);
    vel_err = group_sum(vel_err, G);
    ee_err = group_sum(ee_err, G);
    n_ee = group_sum(n_ee, G);
    if (lane == 0 && env_ok) {
        if (n_ee > 0) ee_err /= n_ee;
        double rgh0 = 0.0;
        if (has_ground) {
            int v, c3[3];
            rgh0 = ground_sample_pieces(g.kind, g.pieces_per_field, pcs, g.data, q0[0], q0[2], &v, c3);
        }
        // root velocities in the normalised frames (only through the 0-weighted term, kept for inf/NaN fidelity)
        double rv0[3], rv1[3];
        quat_rot_vec(h0.hq, qd0, rv0);
        quat_rot_vec(h1.hq, qd1, rv1);
        double dv2 = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double dv = rv1[i] - rv0[i]; dv2 += dv * dv; }
        const double hh0 = q0[1] - rgh0, hh1 = q1[1] - 0.0;
        const double root_err = (hh1 - hh0) * (hh1 - hh0) + 0 * 0.01 * dv2;
        double cs = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double dv = cv1[i] - cv0[i]; cs += dv * dv; }
        const double com_err = 0.1 * cs;
        double pose_w = 0.5, vel_w = 0.05, end_eff_w = 0.15, root_w = 0.1, com_w = 0.2;
        const double total_w = pose_w + vel_w + end_eff_w + root_w + com_w;
        pose_w /= total_w; vel_w /= total_w; end_eff_w /= total_w; root_w /= total_w; com_w /= total_w;
        double r = pose_w * exp(-1 * 2 * pose_err) + vel_w * exp(-1 * 0.005 * vel_err) + end_eff_w * exp(-1 * 40 * ee_err) +
                   root_w * exp(-1 * 10 * root_err) + com_w * exp(-1 * 10 * com_err);
        const bool fell = fallen ? (fallen[e] != 0) : false;
        out_reward[e] = fell ? 0.0 : r;
        if (out_terms) {
            double* t = out_terms + (size_t)e * 5;
            t[0] = fell ? 0.0 : pose_err; t[1] = fell ? 0.0 : vel_err; t[2] = fell ? 0.0 : ee_err;
            t[3] = fell ? 0.0 : root_err; t[4] = fell ? 0.0 : com_err;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_sample_height: cGround::SampleHeight(MatrixXd, VectorXd&) with the scalar overload's numerics
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sample_height(DevGround g, int E, const double* __restrict__ pos, int S, double* __restrict__ out_h,
                unsigned char* __restrict__ out_valid, int* __restrict__ out_cell) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)E * S) return;
    const int e = (int)(tid / S);
    int valid, cell[3];
    const double h = ground_sample_height(g, e, pos[3 * tid], pos[3 * tid + 2], &valid, cell);
    out_h[tid] = h;
    if (out_valid) out_valid[tid] = (unsigned char)valid;
    if (out_cell) { out_cell[3 * tid] = cell[0]; out_cell[3 * tid + 1] = cell[1]; out_cell[3 * tid + 2] = cell[2]; }
}

// ------------------------------------------------------------------------------------------------
// k_poli_state: ParseGround + BuildPoliState -- one warp per env
// ------------------------------------------------------------------------------------------------
struct ObsDims { int G, psize, vsize, pos_dim, is_ct, is3d, F; };

__host__ __device__ inline ObsDims obs_dims(int N, const trl_obs_cfg& cfg) {
    ObsDims d;
    d.G = cfg.num_samples;
    d.is_ct = (cfg.obs_kind == TRL_OBS_CT || cfg.obs_kind == TRL_OBS_CT_3D);
    d.is3d = (cfg.obs_kind == TRL_OBS_CT_3D);
    d.pos_dim = d.is3d ? 3 : 2;
    if (cfg.obs_kind == TRL_OBS_CT) { d.psize = N * 2 + 1; d.vsize = N * 2; }
    else if (cfg.obs_kind == TRL_OBS_CT_3D) { d.psize = N * 7 + 1; d.vsize = N * 6; }
    else { d.psize = N * 2 - 1; d.vsize = N * 2; }
    d.F = d.G + d.psize + d.vsize + (cfg.num_phase_bins >= 0 ? 1 + cfg.num_phase_bins : 0);
    return d;
}

// cMathUtil::RotMatToQuaternion (util/MathUtil.cpp:272-307) for a rotation about y by angle a: R = Ry(a)
TRL_DEV void roty_to_quat(double c, double s, double m11, double* q) {
    // R = [[c,0,s],[0,1,0],[-s,0,c]]
    const double m00 = c, m22 = c;
    const double tr = m00 + m11 + m22;
    if (tr > 0) {
        const double S = sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * S; q[1] = (0.0 - 0.0) / S; q[2] = (s - (-s)) / S; q[3] = (0.0 - 0.0) / S;
    } else if ((m00 > m11) && (m00 > m22)) {
        const double S = sqrt(1.0 + m00 - m11 - m22) * 2;
        q[0] = (0.0 - 0.0) / S; q[1] = 0.25 * S; q[2] = (0.0 + 0.0) / S; q[3] = (s + (-s)) / S;
    } else if (m11 > m22) {
        const double S = sqrt(1.0 + m11 - m00 - m22) * 2;
        q[0] = (s - (-s)) / S; q[1] = (0.0 + 0.0) / S; q[2] = 0.25 * S; q[3] = (0.0 + 0.0) / S;
    } else {
        const double S = sqrt(1.0 + m22 - m00 - m11) * 2;
        q[0] = (0.0 - 0.0) / S; q[1] = (s + (-s)) / S; q[2] = (0.0 + 0.0) / S; q[3] = 0.25 * S;
    }
}

// Per-env record handed from k_obs_env to k_obs_samples: the ground-sample transform (InvRigid(origin_trans) with
// y += ground height under the root), the stance flip and how the env's sample window maps onto its terrain pieces.
struct ObsEnvRec {
    double i00, i02, i20, i22, it0, it1, it2;
    int flip;
    int has_ground;
    long long src_off[4];  // per piece: first float (device array) of the cell rectangle the window covers, a multiple of 4
    int pitch[4];          //            floats between its rows
    int ph[4];             //            rows
    int shift[4];          //            cell (i, j) of the piece sits at patch[shift + j * patch_w + i]
    int cand;      // bit mask of the pieces a sample can fall into (all of them when the window's box is unknown)
    int up;        // >= 0: every sample lies in this piece and in no earlier one (the reference's scan stops there)
    int interior;  // bit 0 (`up` only): the window keeps a cell of distance from the grid's edge, the clamps cannot fire;
                   // bit 1: operands in the range where the division-by-spacing sequence needs no guard
    int staged;    // bit 1 of `interior` holds and the rectangles (of `up`, or of all candidates) fit the patch buffers
    // --- the sample kernels read the 160 bytes above; k_obs_feat also needs: ---
    double ground_h;  // ground height under the root as sampled (non-finite outside the terrain, like the reference's)
    double pad_;
};
static_assert(sizeof(ObsEnvRec) == 176 && offsetof(ObsEnvRec, ground_h) == 160, "read with 16-byte loads");
constexpr int kRecSampleQ = 10;  // 16-byte words of the record the sample kernels read

// k_obs_rec: BuildGroundSampleTrans for one env -- heading, origin transform, ground height under the root, and how the
// env's sample window maps onto the pieces of its terrain (the record k_obs_samples / k_obs_gather work from). FOUR LANES
// PER ENV, LANE = TERRAIN PIECE: the reference scans the (<= 4) pieces of a terrain one after the other, and so did the
// first version of this kernel, one dependent load chain per piece (heading -> piece search -> heights -> cell ranges was
// 20 k cycles of latency per env at one warp per sub-partition). Here every lane tests / maps its own piece and the group
// combines the results with one ballot; the per-env arithmetic is simply repeated on the four lanes.
constexpr int kRecG = 4;
static_assert(kRecG == 4, "one lane per terrain piece (TRL pieces per field <= 4)");

__global__ void __launch_bounds__(128)
k_obs_rec(const DevModel* __restrict__ M, DevGround g, trl_obs_cfg cfg, int E, const double* __restrict__ pose,
          const unsigned char* __restrict__ flip_arr, const double* __restrict__ sample_local,
          ObsEnvRec* __restrict__ rec, double* __restrict__ out_state) {
    KTrace trace_(TR_OBS_ENV);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = tid & (kRecG - 1);  // this lane's piece
    const int e_raw = tid / kRecG;
    const bool env_ok = e_raw < E;
    const int e = env_ok ? e_raw : E - 1;
    const unsigned gshift = (unsigned)(threadIdx.x & 31 & ~(kRecG - 1));  // first lane of the env's group in the warp
    const int N = M->N, n = M->n;
    const ObsDims D = obs_dims(N, cfg);
    const double* q = pose + (size_t)e * n;
    const int flip = flip_arr ? (flip_arr[e] != 0) : 0;
    const int gkind = g.kind, gP = g.pieces_per_field;
    const bool has_ground = gkind != 0;
    const bool field = (gkind == 1 || gkind == 2);
    // this lane's piece: bounds and origin are requested now, before the heading arithmetic needs the root pose
    const DevPiece* pcs = field ? g.pieces + (long long)ground_field_of_env(g, e) * gP : nullptr;
    const bool mine = field && s < gP;
    double b0 = 0, b1 = 0, b2 = 0, b3 = 0, pox = 0, poz = 0;
    if (mine) {
        const DevPiece& pc = pcs[s];
        b0 = pc.bounds[0]; b1 = pc.bounds[1]; b2 = pc.bounds[2]; b3 = pc.bounds[3];
        pox = pc.origin[0]; poz = pc.origin[2];
    }

    // cKinTree::CalcHeading / BuildOriginTrans (anim/KinTree.cpp:1604-1639): origin = Ry(-heading) * T(-root_xz)
    const double rx = q[0], rz = q[2];
    const double qr[4] = {q[3], q[4], q[5], q[6]};
    const double xdir[3] = {1, 0, 0};
    double rd[3];
    quat_rot_vec(qr, xdir, rd);
    const double heading = atan2(-rd[2], rd[0]);
    double sh, ch;
    sincos(-heading, &sh, &ch);
    // RotateMat(axis y, theta): [[c,0,s],[0,1,0],[-s,0,c]] (c + y*y*(1-c) on the diagonal, util/MathUtil.cpp:157-174)
    const double o00 = ch, o02 = sh, o20 = -sh, o22 = ch;
    const double o11 = ch + (1 - ch);
    const double oz = 0.0;  // structurally-zero entries, kept in the products so inf/NaN propagate like the reference
    const double nx = -rx, ny = -0.0, nz = -rz;
    const double ot0 = o00 * nx + oz * ny + o02 * nz;
    const double ot1 = oz * nx + o11 * ny + oz * nz;
    const double ot2 = o20 * nx + oz * ny + o22 * nz;

    // ground height under the root (TerrainRLCharController.cpp:261-279,370-380): the first piece (in scan order) that
    // contains the root position -- every lane tests its own piece, the lowest set bit of the group's ballot wins
    double ground_h = 0.0;
    if (field) {
        const bool hit = mine && ((gkind == 1) ? (rx >= b0 && rx <= b1) : (rx >= b0 && rx <= b1 && rz >= b2 && rz <= b3));
        const unsigned hm = (__ballot_sync(0xffffffffu, hit) >> gshift) & 0xfu;
        const int pi = hm ? (__ffs(hm) - 1) : -1;
        int v0, c0[3];
        ground_h = ground_sample_piece(gkind, pcs, pi, g.data, rx, rz, &v0, c0);
    }
    const double ground_h_finite = isfinite(ground_h) ? ground_h : 0.0;

    // ground sample transform = InvRigid(origin_trans), y += ground_h (finite)
    const double i00 = o00, i02 = o20, i20 = o02, i22 = o22, i11 = o11;
    const double it0 = -(i00 * ot0 + oz * ot1 + i02 * ot2);
    const double it1 = -(oz * ot0 + i11 * ot1 + oz * ot2) + ground_h_finite;
    const double it2 = -(i20 * ot0 + oz * ot1 + i22 * ot2);

    // ---- the sample window on the terrain pieces ----
    int cand = 0xf, up = -1, interior = 0, staged = 0;  // candidates: all pieces unless the bounding-box test narrows them
    long long my_off = 0;
    int my_pitch = 0, my_ph = 0, my_shift = 0;
    if (field && cfg.num_samples > 0 && sample_local) {
        // Bounding box of all sample positions: the position is affine (hence weakly monotone, also after
        // rounding) in the local coordinates, so its extremes sit at the corners of the local bounding box. The
        // box is widened by a few ulps because the sample kernel may contract the same expression differently.
        const double* bb = sample_local + 2 * cfg.num_samples;
        const double lx[2] = {__ldg(bb), __ldg(bb + 1)};
        double lz[2] = {__ldg(bb + 2), __ldg(bb + 3)};
        if (flip && (cfg.obs_kind == TRL_OBS_CT || cfg.obs_kind == TRL_OBS_CT_3D)) {
            const double tt = lz[0];
            lz[0] = -lz[1];
            lz[1] = -tt;
        }
        double x0 = CUDART_INF, x1 = -CUDART_INF, z0 = CUDART_INF, z1 = -CUDART_INF;
        bool finite = true;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double px = i00 * lx[a] + i02 * lz[b] + it0;
                const double pz = i20 * lx[a] + i22 * lz[b] + it2;
                finite = finite && isfinite(px) && isfinite(pz);
                x0 = (px < x0) ? px : x0; x1 = (px > x1) ? px : x1;
                z0 = (pz < z0) ? pz : z0; z1 = (pz > z1) ? pz : z1;
            }
        const double ex = 1e-12 * (fabs(x0) + fabs(x1) + 1.0), ez = 1e-12 * (fabs(z0) + fabs(z1) + 1.0);
        x0 -= ex; x1 += ex; z0 -= ez; z1 += ez;
        {   // (`finite` is uniform over an env's four lanes but not over the warp: the ballots below are unconditional,
            // a non-finite box simply touches nothing and the record keeps its defaults)
            // The staged sample code divides by the grid spacing with the unguarded correction sequence of
            // div_by_const, exact for operands x = pos - origin with x == 0 or 2^-900 < |x| < 2^900. Every quantity x
            // is built from being zero or in [2^-300, 2^300] guarantees that (the sample table is checked on the host).
            auto okmag = [](double v) { const double a = fabs(v); return a == 0.0 || (a >= 0x1p-300 && a <= 0x1p300); };
            const bool safe0 = g.table_ok != 0 && okmag(i00) && okmag(i02) && okmag(i20) && okmag(i22) && okmag(it0) && okmag(it2) &&
                               fabs(x0) < 0x1p300 && fabs(x1) < 0x1p300 && fabs(z0) < 0x1p300 && fabs(z1) < 0x1p300;
            // Cell range of the box in a piece's grid. The estimate (plain arithmetic) is within a tiny fraction of a
            // cell of the exact coordinate sequence used per sample; one cell of slack on either side absorbs it, one
            // more on the high side holds the i + 1 / j + 1 vertices.
            auto range = [](double lo_w, double hi_w, double o, double rs, double half, int res, int& first, int& count, bool& clear) {
                const double ca = (lo_w - o) * rs + half, cb = (hi_w - o) * rs + half;
                double lo = (ca < cb) ? ca : cb, hi = (ca < cb) ? cb : ca;
                lo = clampd(lo, -4.0, 1.0e9);
                hi = clampd(hi, -4.0, 1.0e9);
                const int ilo = (int)floor(lo) - 1, ihi = (int)floor(hi) + 2;
                clear = ilo >= 0 && ihi <= res - 1;
                first = max(ilo, 0);
                count = max(min(ihi, res - 1) - first + 1, 0);
            };
            // this lane's piece against the box
            const bool inside = finite && mine && ((gkind == 1) ? (x0 >= b0 && x1 <= b1) : (x0 >= b0 && x1 <= b1 && z0 >= b2 && z1 <= b3));
            const bool apart = !finite || !mine || ((gkind == 1) ? (x1 < b0 || x0 > b1) : (x1 < b0 || x0 > b1 || z1 < b2 || z0 > b3));
            const bool origin_ok = apart || (okmag(pox) && okmag(poz));
            bool p_inner = false, p_fits = false;
            int pi0 = 0, pj0 = 0;
            if (!apart && g.patch_nbuf > 0) {  // cell rectangle, only when k_obs_samples stages
                const DevPiece& pc = pcs[s];
                bool cx, cz = true;
                int first_x, count_x, count_z = 1;
                range(x0, x1, pox, pc.rscale[0], pc.half[0], pc.res_x, first_x, count_x, cx);
                if (gkind == 2) range(z0, z1, poz, pc.rscale[1], pc.half[1], pc.res_z, pj0, count_z, cz);
                pi0 = first_x & ~3;  // staged rows start at a column multiple of 4 (16-byte aligned bulk copies)
                const int pw = first_x + count_x - pi0;
                my_ph = count_z;
                my_pitch = pc.pitch;
                my_off = pc.data_offset + (long long)pj0 * pc.pitch + pi0;
                p_inner = cx && cz;
                p_fits = count_x >= 1 && pw <= g.patch_w && count_z >= 1 && count_z <= g.patch_h;
            }
            // combine over the pieces (the scan order of the reference = lane order)
            const unsigned m_touch = (__ballot_sync(0xffffffffu, !apart) >> gshift) & 0xfu;
            const unsigned m_inside = (__ballot_sync(0xffffffffu, inside) >> gshift) & 0xfu;
            const unsigned m_ok = (__ballot_sync(0xffffffffu, origin_ok) >> gshift) & 0xfu;
            const unsigned m_inner = (__ballot_sync(0xffffffffu, p_inner) >> gshift) & 0xfu;
            const unsigned m_fits = (__ballot_sync(0xffffffffu, p_fits) >> gshift) & 0xfu;
            // `up`: a piece that holds the whole box with no earlier piece touching it (the scan stops there for every
            // sample). Scanning in order, that is the last inside-piece at or before the first touching one.
            if (finite) {
                const int first_touch = m_touch ? (__ffs(m_touch) - 1) : 3;
                const unsigned up_set = m_inside & ((2u << first_touch) - 1u);
                up = up_set ? (31 - __clz(up_set)) : -1;
                cand = (int)m_touch;
                const bool safe = safe0 && m_ok == 0xfu;
                const unsigned smask = (up >= 0) ? (1u << up) : m_touch;
                const int nb = __popc(smask);
                if ((smask >> s) & 1u) my_shift = __popc(smask & ((1u << s) - 1u)) * g.patch_w * g.patch_h - pj0 * g.patch_w - pi0;
                staged = (safe && smask != 0 && (m_fits & smask) == smask && nb <= g.patch_nbuf) ? 1 : 0;
                interior = ((up >= 0 && ((m_inner >> up) & 1u)) ? 1 : 0) | (safe ? 2 : 0);
            }
        }
    }
    if (env_ok) {
        ObsEnvRec* r = rec + e;
        r->src_off[s] = my_off; r->pitch[s] = my_pitch; r->ph[s] = my_ph; r->shift[s] = my_shift;
        if (s == 0) {
            r->i00 = i00; r->i02 = i02; r->i20 = i20; r->i22 = i22; r->it0 = it0; r->it1 = it1; r->it2 = it2;
            r->flip = flip; r->has_ground = has_ground ? 1 : 0;
            r->cand = cand; r->up = up; r->interior = interior; r->staged = staged;
            r->ground_h = ground_h;
        }
        if (!M->bx.all_valid) {  // the reference starts from VectorXd::Zero(psize): slots of shape-less bodies stay zero
            double* opose = out_state + (size_t)e * D.F + D.G;
            for (int i = 1 + s; i < D.psize; i += kRecG) opose[i] = 0.0;
        }
    }
}

// k_obs_feat: BuildPoliStatePose / Vel (+ phase / hopper / stance flip) -- THREAD PER (env, body). Each thread rebuilds the
// origin transform from the env's record (the heading terms are the record's rotation entries) and writes its body's
// feature slots; body 0's thread also writes the root height and the phase block.
__global__ void __launch_bounds__(128)
k_obs_feat(const DevModel* __restrict__ M, trl_obs_cfg cfg, int E, const double* __restrict__ pose,
           const double* __restrict__ vel, const double* __restrict__ body_pos, const double* __restrict__ body_vel,
           const double* __restrict__ body_quat, const double* __restrict__ body_angvel,
           const double* __restrict__ ground_vel, const double* __restrict__ phase,
           const ObsEnvRec* __restrict__ rec, double* __restrict__ out_state) {
    KTrace trace_(TR_OBS_FEAT);
    const int N = M->N, n = M->n;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)E * N) return;
    const int e = (int)(tid / N);
    const int i = (int)(tid - (long long)e * N);  // this thread's body
    const ModelIdx* X = &M->ix;
    const ObsDims D = obs_dims(N, cfg);
    const double* q = pose + (size_t)e * n;
    double* out = out_state + (size_t)e * D.F;
    const ObsEnvRec* r = rec + e;
    const double rx = q[0], ry = q[1], rz = q[2];
    const double qr[4] = {q[3], q[4], q[5], q[6]};
    const int flip = r->flip;
    const double ground_h = r->ground_h;
    // origin transform: the record holds its inverse rotation (i00 = o00, i02 = o20, i20 = o02, i22 = o22)
    const double ch = r->i00, sh = r->i20;
    const double o00 = ch, o02 = sh, o20 = -sh, o22 = ch;
    const double o11 = ch + (1 - ch);
    const double oz = 0.0;  // structurally-zero entries, kept in the products so inf/NaN propagate like the reference
    const double nx = -rx, ny = -0.0, nz = -rz;
    const double ot0 = o00 * nx + oz * ny + o02 * nz;
    const double ot1 = oz * nx + o11 * ny + oz * nz;
    const double ot2 = o20 * nx + oz * ny + o22 * nz;
    // ---- pose block ----
    double* opose = out + D.G;
    double* ovel = opose + D.psize;
    const double zsign = (D.is_ct && flip) ? -1.0 : 1.0;
    const double ryg = ry - ground_h;
    const double rr0 = o00 * rx + oz * ryg + o02 * rz + ot0;
    const double rr1 = oz * rx + o11 * ryg + oz * rz + ot1;
    const double rr2 = zsign * (o20 * rx + oz * ryg + o22 * rz + ot2);
    if (i == 0) opose[0] = rr1;
    const int first = D.is_ct ? 0 : 1;
    const int per = D.pos_dim + (D.is3d ? 4 : 0);
    // cBipedController3D::FlipPoliPoseStance (BipedController3D.cpp:1811-1830) swaps the two leg blocks at the tail
    // of the pose / vel feature vectors; it is applied on the fly as an index permutation of the writes.
    const int nlp = (!D.is_ct && flip) ? ((N - 1) / 2) * 2 : 0;
    auto perm = [nlp](int idx, int size) {
        if (nlp == 0) return idx;
        const int a = size - 1 - idx;          // distance from the end
        if (a < nlp) return idx - nlp;          // second leg block -> first
        if (a < 2 * nlp) return idx + nlp;      // first leg block -> second
        return idx;
    };
    const int slot = (i >= first) ? (D.is_ct ? X->slot_ct[i] : X->slot_base[i]) : -1;
    if (slot >= 0) {
        const double* bp = body_pos + ((size_t)e * N + i) * 3;
        const double bx = bp[0], by = bp[1] - ground_h, bz = bp[2];
        const double f0 = (o00 * bx + oz * by + o02 * bz + ot0) - rr0;
        const double f1 = (oz * bx + o11 * by + oz * bz + ot1) - rr1;
        const double f2 = zsign * (o20 * bx + oz * by + o22 * bz + ot2) - rr2;
        const int b0 = 1 + slot * per;
        opose[perm(b0, D.psize)] = f0;
        opose[perm(b0 + 1, D.psize)] = f1;
        if (D.is3d) {
            double* dst = opose + b0;
            dst[2] = f2;
            double oq[4], bq[4], rq[4];
            roty_to_quat(ch, sh, o11, oq);
            const double* bqp = body_quat + ((size_t)e * N + i) * 4;
            bq[0] = bqp[0]; bq[1] = bqp[1]; bq[2] = bqp[2]; bq[3] = bqp[3];
            quat_mul(oq, bq, rq);
            if (flip) { rq[1] = -rq[1]; rq[2] = -rq[2]; }
            if (rq[0] < 0) { rq[0] = -rq[0]; rq[1] = -rq[1]; rq[2] = -rq[2]; rq[3] = -rq[3]; }
            dst[3] = rq[0]; dst[4] = rq[1]; dst[5] = rq[2]; dst[6] = rq[3];
        }
    }
    // ---- vel block ----
    double gv[3] = {0, 0, 0};
    if (ground_vel) { gv[0] = ground_vel[3 * (size_t)e]; gv[1] = ground_vel[3 * (size_t)e + 1]; gv[2] = ground_vel[3 * (size_t)e + 2]; }
    const int vper = D.pos_dim + (D.is3d ? 3 : 0);
    const bool hopper = cfg.obs_kind == TRL_OBS_TERRAIN_RL_HOPPER;
    {
        const double* bv = body_vel + ((size_t)e * N + i) * 3;
        const double vx = bv[0] - gv[0], vy = bv[1] - gv[1], vz = bv[2] - gv[2];
        const int b0 = i * vper;
        const double w0 = o00 * vx + oz * vy + o02 * vz;
        double w1 = oz * vx + o11 * vy + oz * vz;
        if (hopper && b0 + 1 == D.vsize - 1) w1 = vel[(size_t)e * n + 5];  // root omega_z (MonopedHopperController.cpp:1545-1552)
        ovel[perm(b0, D.vsize)] = w0;
        ovel[perm(b0 + 1, D.vsize)] = w1;
        if (D.is3d) {
            double* dst = ovel + b0;
            dst[2] = zsign * (o20 * vx + oz * vy + o22 * vz);
            const double* av = body_angvel + ((size_t)e * N + i) * 3;
            double a0 = o00 * av[0] + oz * av[1] + o02 * av[2];
            double a1 = oz * av[0] + o11 * av[1] + oz * av[2];
            double a2 = zsign * (o20 * av[0] + oz * av[1] + o22 * av[2]);
            if (flip) { a0 = -a0; a1 = -a1; a2 = -a2; }
            dst[3] = a0; dst[4] = a1; dst[5] = a2;
        }
    }
    // cMonopedHopperController::BuildPoliStatePose: the last pose entry is overwritten by the signed root rotation
    // angle (cSimCharacter::GetRootRotation -> RotMatToAxisAngle, util/MathUtil.cpp:239-260). That entry belongs to
    // body N-1, so the write is issued from that body's lane to keep program order.
    if (hopper && i == N - 1) {
        double R[9];
        quat_to_mat(qr[0], qr[1], qr[2], qr[3], R);
        double c = (R[0] + R[4] + R[8] - 1) * 0.5;
        c = clampd(c, -1.0, 1.0);
        double theta = acos(c);
        double az = 1.0;
        if (!(fabs(theta) < 0.00001)) {
            const double m21 = R[7] - R[5], m02 = R[2] - R[6], m10 = R[3] - R[1];
            az = m10 / sqrt(m21 * m21 + m02 * m02 + m10 * m10);
        }
        if (az < 0) theta = -theta;
        opose[D.psize - 1] = theta;
    }
    if (cfg.num_phase_bins >= 0 && i == 0) {  // cCtPhaseController::BuildPhaseState (CtPhaseController.cpp:120-139)
        double* ph = ovel + D.vsize;
        const double pv = phase ? phase[e] : 0.0;
        for (int k = 0; k < 1 + cfg.num_phase_bins; ++k) ph[k] = 0.0;
        ph[0] = pv;
        if (cfg.num_phase_bins > 0) {
            const int bin = (int)(pv * cfg.num_phase_bins);
            if (bin >= 0 && bin < cfg.num_phase_bins) ph[bin + 1] = 1.0;
        }
    }
}

// k_obs_samples: ParseGround's sample loop -- ONE WARP PER ENV, lane = sample (s = lane, lane + 32, ...).
//
// The cells an env's samples can touch are known before any sample is evaluated: k_obs_env maps the world-space box of
// the sample positions onto the (<= 4) terrain pieces of the env's terrain -- which pieces the box touches, and the cell
// rectangle it covers in each of them -- and this kernel STAGES those rectangles in shared memory with bulk asynchronous
// copies (cp.async.bulk + mbarrier, the TMA path): lane r moves row r of a rectangle, so ONE warp instruction puts up to
// 32 rows in flight, lane 0 moves the env's piece descriptors, and the warp waits once on the mbarrier. (The device copy
// of a heightfield is re-pitched at registration so that row segments starting at a column multiple of 4 are 16-byte
// aligned.) A 16 x 16 window of biped3D covers <= 26 x 26 cells; its buffer is 27 rows x 32 floats = 3.4 KB. Per sample
// there is then no dependent global gather: the 3 (2-D: 2) vertices come from shared memory.
// ncu A/B against the gather version: profiles/README.md.
//
//   * box inside one piece, no earlier piece touching it ("uniform", ~60 % of the envs when pieces meet at the origin):
//     the piece's constants live in registers. If the box also keeps a cell of distance from the grid's edge
//     ("interior") the reference's clamps cannot fire and are skipped.
//   * box touching several pieces: every touched piece's rectangle is staged (one buffer each); per sample the first
//     containing piece (the reference's scan order) selects the constants from the descriptors in shared memory.
//   * no finite box, more touched pieces than buffers, rectangles larger than a buffer, operands outside the guard-free
//     range: per-sample search and global gathers (the general code).
//
// Grid coordinates keep the reference's rounding sequence exactly (sub, IEEE-equal division by the spacing, add);
// (int)c and c - (double)(int)c are taken with one round-down add of 2^52 (exact for 0 <= c < 2^31) instead of two
// conversion instructions.
constexpr int kObsWarps = 2;       // envs per block (small blocks: a finished env's shared memory is reusable at once)
constexpr int kObsUnroll = 4;      // samples in flight per lane
constexpr double kTwo52 = 4503599627370496.0;
constexpr int kObsHead = 16 + 4 * (int)sizeof(DevPiece);  // per warp: mbarrier (16 bytes), 4 piece descriptors

__host__ __device__ inline size_t obs_warp_bytes(const DevGround& g) {
    return (size_t)kObsHead + (size_t)g.patch_nbuf * g.patch_w * g.patch_h * sizeof(float);
}

// floor and fraction of a grid coordinate 0 <= c < 2^31
TRL_DEV void floor_frac(double c, int& i, double& frac) {
    const double t = __dadd_rd(c, kTwo52);  // 2^52 + floor(c), exactly
    i = __double2loint(t);
    frac = __dsub_rn(c, __dsub_rn(t, kTwo52));
}

// cMathUtil::Clamp(c, 0, hi) for a finite c (the staged paths): c < 0 is the sign bit (-0 -> +0 like the comparison)
TRL_DEV double clamp0(double c, double hi) {
    c = (c < hi) ? c : hi;
    return (__double2hiint(c) < 0) ? 0.0 : c;
}

// x / d by Markstein's correction sequence from r = RN(1 / d) (see div_by_const), without its range guard: k_obs_env
// admits an env to the staged paths only if its operands are in the range where the sequence is exact
TRL_DEV double div_by_const_unguarded(double x, double d, double r) {
    const double q0 = __dmul_rn(x, r);
    const double e0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(e0, r, q0);
    const double e1 = __fma_rn(-q1, d, x);
    return __fma_rn(e1, r, q1);
}

// one height sample out of the staged patch: cell (i, j) of the piece sits at patch[shift + j * pw + i]; `c` is the piece
// descriptor in registers (uniform piece) or in shared memory (per-sample piece)
template <int KIND, bool INTERIOR>
TRL_DEV double patch_sample(const DevPiece& c, const float* __restrict__ patch, int shift, int pw, double px, double pz,
                            int& ci, int& cj) {
    const double sy = c.scale[1];
    if (KIND == 1) {
        double c0 = __dadd_rn(div_by_const_unguarded(__dsub_rn(px, c.origin[0]), c.scale[0], c.rscale[0]), c.half[0]);
        if (!INTERIOR) c0 = clamp0(c0, c.res_x - 1.0);
        int i;
        double lerp;
        floor_frac(c0, i, lerp);
        const int j = INTERIOR ? i + 1 : min(c.res_x - 1, i + 1);
        const double a = __dmul_rn((double)patch[shift + i], sy), b = __dmul_rn((double)patch[shift + j], sy);
        ci = i; cj = j;
        return __dadd_rn(__dmul_rn(__dsub_rn(1.0, lerp), a), __dmul_rn(lerp, b));
    } else {
        const double qx = div_by_const_unguarded(__dsub_rn(px, c.origin[0]), c.scale[0], c.rscale[0]);
        const double qz = div_by_const_unguarded(__dsub_rn(pz, c.origin[2]), c.scale[2], c.rscale[1]);
        double c0 = __dadd_rn(qx, c.half[0]), c1 = __dadd_rn(qz, c.half[1]);
        if (!INTERIOR) {
            c0 = clamp0(c0, c.cmax3[0]);
            c1 = clamp0(c1, c.cmax3[1]);
        }
        int i, j;
        double lx, lz;
        floor_frac(c0, i, lx);
        floor_frac(c1, j, lz);
        // cMathUtil::CalcBarycentric on the unit triangles, see sample_prep_3d: with (a, b) = (lx, lz) in the lower and
        // (1 - lx, 1 - lz) in the upper triangle both cases are d20 = a, d21 = a + b (-RN(lx - 1) == RN(1 - lx)).
        const bool lower = (lz <= lx);
        const double a = lower ? lx : __dsub_rn(1.0, lx);
        const double b = lower ? lz : __dsub_rn(1.0, lz);
        const double d21 = __dadd_rn(a, b);
        const double bv = __fma_rn(2.0, a, -d21);  // RN(RN(2 a) - d21): 2 a is exact
        const double bw = __dsub_rn(d21, a);
        const double w0 = __dsub_rn(__dsub_rn(1.0, bv), bw);
        const int idx = shift + j * pw + i;
        const int diag = idx + pw + 1;
        const double v0 = __dmul_rn((double)patch[lower ? idx : diag], sy);
        const double v1 = __dmul_rn((double)patch[lower ? idx + 1 : idx + pw], sy);
        const double v2 = __dmul_rn((double)patch[lower ? diag : idx], sy);
        ci = i; cj = j;
        return __dadd_rn(__dadd_rn(__dmul_rn(w0, v0), __dmul_rn(bv, v1)), __dmul_rn(bw, v2));
    }
}

template <int KIND, bool WANT_CELL>  // ground kind (1 = var2d, 2 = var3d, 0/3 = none/plane) as a compile-time constant
__global__ void __launch_bounds__(kObsWarps * 32, 14)
k_obs_samples(DevGround g, trl_obs_cfg cfg, int F, int E, const ObsEnvRec* __restrict__ rec,
              const double* __restrict__ sample_local, double* __restrict__ out_state, int* __restrict__ out_cell) {
    KTrace trace_(TR_OBS_SAMPLES);
    extern __shared__ __align__(128) unsigned char s_obs_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * kObsWarps + warp;
    if (e >= E) return;  // warp-uniform; the kernel has no block barrier
    const int G = cfg.num_samples;
    ObsEnvRec r;
    {
        const uint4* src = reinterpret_cast<const uint4*>(rec + e);
        uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int k = 0; k < kRecSampleQ; ++k) dst[k] = __ldg(src + k);
    }
    const bool flip_z = r.flip && (cfg.obs_kind == TRL_OBS_CT || cfg.obs_kind == TRL_OBS_CT_3D);
    const double i02 = flip_z ? -r.i02 : r.i02, i22 = flip_z ? -r.i22 : r.i22;  // local z -> -z folded into the map
    double* out = out_state + (size_t)e * F;
    int* oc = (WANT_CELL && out_cell) ? out_cell + (size_t)e * G * 3 : nullptr;
    const double2* sl2 = reinterpret_cast<const double2*>(sample_local);

    if (KIND != 1 && KIND != 2) {  // no heightfield: cGround::SampleHeight = 0 (sim/Ground.cpp:176-186)
        const double hs = r.has_ground ? (0.0 - r.it1) : 0.0;
        for (int s = lane; s < G; s += 32) {
            out[s] = hs;
            if (oc) { oc[3 * s] = -1; oc[3 * s + 1] = 0; oc[3 * s + 2] = 0; }
        }
        return;
    }

    const int gP = g.pieces_per_field;
    const DevPiece* pcs = g.pieces + (long long)ground_field_of_env(g, e) * gP;
    const unsigned cand = (unsigned)r.cand;
    const int up = r.up;
    const int pw = g.patch_w;

    if (r.staged) {
        unsigned char* wbase = s_obs_raw + (size_t)warp * obs_warp_bytes(g);
        const unsigned bar_a = smem_u32(wbase);
        DevPiece* spc = reinterpret_cast<DevPiece*>(wbase + 16);
        float* patch = reinterpret_cast<float*>(wbase + kObsHead);
        // ---- stage: the env's piece descriptors (lane 0) and one row of a rectangle per lane, all on one mbarrier ----
        const unsigned smask = (up >= 0) ? (1u << up) : cand;
        const unsigned row_bytes = (unsigned)pw * 4u, desc_bytes = (unsigned)gP * (unsigned)sizeof(DevPiece);
        if (lane == 0) {
            unsigned total = desc_bytes;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if ((smask >> t) & 1u) total += (unsigned)r.ph[t] * row_bytes;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(spc)), "l"(pcs), "r"(desc_bytes), "r"(bar_a) : "memory");
        }
        __syncwarp();
        {
            int buf = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (!((smask >> t) & 1u)) continue;
                const float* src = g.data + r.src_off[t];
                float* dst = patch + (size_t)buf * pw * g.patch_h;
                for (int row = lane; row < r.ph[t]; row += 32)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(dst + row * pw)), "l"(src + (long long)row * r.pitch[t]), "r"(row_bytes), "r"(bar_a) : "memory");
                ++buf;
            }
        }
        double2 nxt[kObsUnroll];  // the first group's entries of the sample table: loaded while the copies are in flight
#pragma unroll
        for (int k = 0; k < kObsUnroll; ++k) {
            const int s = lane + 32 * k;
            nxt[k] = __ldg(sl2 + (s < G ? s : 0));
        }
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "TRL_OWAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra TRL_ODONE_%=;\n"
            "bra TRL_OWAIT_%=;\n"
            "TRL_ODONE_%=:\n"
            "}\n" ::"r"(bar_a)
            : "memory");

        if (up >= 0) {
            // ---- uniform piece: constants in registers ----
            const DevPiece c = spc[up];
            const int sh = (up == 0) ? r.shift[0] : (up == 1) ? r.shift[1] : (up == 2) ? r.shift[2] : r.shift[3];
            auto run = [&](auto inner_tag) {
                constexpr bool INNER = decltype(inner_tag)::value;
                for (int s0 = lane; s0 < G; s0 += 32 * kObsUnroll) {
                    double2 cur[kObsUnroll];
#pragma unroll
                    for (int k = 0; k < kObsUnroll; ++k) cur[k] = nxt[k];
                    if (s0 + 32 * kObsUnroll < G) {  // the next group's table entries load while this group computes
#pragma unroll
                        for (int k = 0; k < kObsUnroll; ++k) {
                            const int s = s0 + 32 * (kObsUnroll + k);
                            nxt[k] = __ldg(sl2 + (s < G ? s : 0));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kObsUnroll; ++k) {
                        const int s = s0 + 32 * k;
                        if (s >= G) break;
                        const double px = fma(r.i00, cur[k].x, fma(i02, cur[k].y, r.it0));
                        const double pz = fma(r.i20, cur[k].x, fma(i22, cur[k].y, r.it2));
                        int ci, cj;
                        out[s] = patch_sample<KIND, INNER>(c, patch, sh, pw, px, pz, ci, cj) - r.it1;
                        if (WANT_CELL && oc) { oc[3 * s] = up; oc[3 * s + 1] = ci; oc[3 * s + 2] = cj; }
                    }
                }
            };
            if (r.interior & 1) run(std::true_type{});
            else run(std::false_type{});
            return;
        }
        // ---- several pieces staged: first containing candidate per sample selects the descriptor (shared memory) ----
        for (int s0 = lane; s0 < G; s0 += 32 * kObsUnroll) {
            double2 cur[kObsUnroll];
#pragma unroll
            for (int k = 0; k < kObsUnroll; ++k) cur[k] = nxt[k];
            if (s0 + 32 * kObsUnroll < G) {
#pragma unroll
                for (int k = 0; k < kObsUnroll; ++k) {
                    const int s = s0 + 32 * (kObsUnroll + k);
                    nxt[k] = __ldg(sl2 + (s < G ? s : 0));
                }
            }
#pragma unroll
            for (int k = 0; k < kObsUnroll; ++k) {
                const int s = s0 + 32 * k;
                if (s >= G) break;
                const double px = fma(r.i00, cur[k].x, fma(i02, cur[k].y, r.it0));
                const double pz = fma(r.i20, cur[k].x, fma(i22, cur[k].y, r.it2));
                int pi = -1;
#pragma unroll
                for (int t = 3; t >= 0; --t)
                    if (((cand >> t) & 1u) && piece_contains(KIND, spc[t], px, pz)) pi = t;
                int ci = 0, cj = 0;
                double hs = -CUDART_INF;
                if (pi >= 0) {
                    const int sh = (pi == 0) ? r.shift[0] : (pi == 1) ? r.shift[1] : (pi == 2) ? r.shift[2] : r.shift[3];
                    hs = patch_sample<KIND, false>(spc[pi], patch, sh, pw, px, pz, ci, cj);
                }
                out[s] = hs - r.it1;
                if (WANT_CELL && oc) { oc[3 * s] = pi; oc[3 * s + 1] = ci; oc[3 * s + 2] = cj; }
            }
        }
        return;
    }
    // ---- general code: per-sample search over the candidate pieces, vertices gathered from global memory ----
    for (int s = lane; s < G; s += 32) {
        const double2 l2 = __ldg(sl2 + s);
        const double px = fma(r.i00, l2.x, fma(i02, l2.y, r.it0));
        const double pz = fma(r.i20, l2.x, fma(i22, l2.y, r.it2));
        int pi = -1;
#pragma unroll
        for (int t = 3; t >= 0; --t)
            if (t < gP && ((cand >> t) & 1u) && piece_contains(KIND, pcs[t], px, pz)) pi = t;
        int vs, cell[3];
        out[s] = ground_sample_piece(KIND, pcs, pi, g.data, px, pz, &vs, cell) - r.it1;
        if (WANT_CELL && oc) { oc[3 * s] = cell[0]; oc[3 * s + 1] = cell[1]; oc[3 * s + 2] = cell[2]; }
    }
}

// k_obs_gather: the same sample loop with the vertices gathered straight from global memory -- block = one env,
// kObsSPT samples per thread (independent load chains the compiler interleaves), the env's record and piece descriptors
// staged once in shared memory. Many small independent blocks keep ~35 warps per SM resident; on the 3-D heightfields
// (3 scattered vertices per sample, a rectangle of ~26 rows to stage per env) this beats the staged kernel, which wins on
// the 2-D ones (one contiguous row per env): see the ncu A/B in profiles/README.md. Same arithmetic (patch_sample reads
// "patch" = the piece's device array with its row pitch).
#ifndef TRL_GATHER_SPT
#define TRL_GATHER_SPT 4
#endif
constexpr int kObsSPT = TRL_GATHER_SPT;
template <int KIND, bool WANT_CELL>
#ifndef TRL_GATHER_MINB
#define TRL_GATHER_MINB 8  // 32 registers: all 64 warps of an SM resident (5 -> 8: step 0.0830 -> 0.0782 ms; the kernel waits on HBM)
#endif
__global__ void __launch_bounds__(256, TRL_GATHER_MINB)
k_obs_gather(DevGround g, trl_obs_cfg cfg, int F, int E, const ObsEnvRec* __restrict__ rec,
             const double* __restrict__ sample_local, double* __restrict__ out_state, int* __restrict__ out_cell) {
    KTrace trace_(TR_OBS_SAMPLES);
    __shared__ __align__(16) DevPiece s_pieces[4];
    __shared__ __align__(16) ObsEnvRec s_rec;
    const int e = blockIdx.x;
    const int G = cfg.num_samples;
    const int gP = g.pieces_per_field;
    const DevPiece* pcs = g.pieces + (long long)ground_field_of_env(g, e) * gP;
    {
        constexpr int kRecW = kRecSampleQ;
        const int pw = gP * (int)(sizeof(DevPiece) / 16);
        for (int t = threadIdx.x; t < pw + kRecW; t += blockDim.x) {
            if (t < pw) reinterpret_cast<uint4*>(s_pieces)[t] = __ldg(reinterpret_cast<const uint4*>(pcs) + t);
            else reinterpret_cast<uint4*>(&s_rec)[t - pw] = __ldg(reinterpret_cast<const uint4*>(rec + e) + (t - pw));
        }
    }
    __syncthreads();
    const ObsEnvRec r = s_rec;  // registers
    const bool flip_z = r.flip && (cfg.obs_kind == TRL_OBS_CT || cfg.obs_kind == TRL_OBS_CT_3D);
    const double i02 = flip_z ? -r.i02 : r.i02, i22 = flip_z ? -r.i22 : r.i22;
    double* out = out_state + (size_t)e * F;
    int* oc = (WANT_CELL && out_cell) ? out_cell + (size_t)e * G * 3 : nullptr;
    const double2* sl2 = reinterpret_cast<const double2*>(sample_local);
    const int up = r.up;
    const unsigned cand = (unsigned)r.cand;
    const bool safe = (r.interior & 2) != 0;
    if (safe && up >= 0) {
        // every sample in one piece, operands in the guard-free range: the piece's descriptor in registers
        const DevPiece c = s_pieces[up];
        const float* base = g.data + c.data_offset;
        auto run = [&](auto inner_tag) {
            constexpr bool INNER = decltype(inner_tag)::value;
            for (int s0 = threadIdx.x; s0 < G; s0 += blockDim.x * kObsSPT) {
#pragma unroll
                for (int k = 0; k < kObsSPT; ++k) {
                    const int s = s0 + k * blockDim.x;
                    if (s >= G) break;
                    const double2 l2 = __ldg(sl2 + s);
                    const double px = fma(r.i00, l2.x, fma(i02, l2.y, r.it0));
                    const double pz = fma(r.i20, l2.x, fma(i22, l2.y, r.it2));
                    int ci, cj;
                    out[s] = patch_sample<KIND, INNER>(c, base, 0, c.pitch, px, pz, ci, cj) - r.it1;
                    if (WANT_CELL && oc) { oc[3 * s] = up; oc[3 * s + 1] = ci; oc[3 * s + 2] = cj; }
                }
            }
        };
        if (r.interior & 1) run(std::true_type{});
        else run(std::false_type{});
        return;
    }
    // general code: per-sample search over the candidate pieces (lowest index first like the reference's scan)
    for (int s0 = threadIdx.x; s0 < G; s0 += blockDim.x * kObsSPT) {
#pragma unroll
        for (int k = 0; k < kObsSPT; ++k) {
            const int s = s0 + k * blockDim.x;
            if (s >= G) break;
            const double2 l2 = __ldg(sl2 + s);
            const double px = fma(r.i00, l2.x, fma(i02, l2.y, r.it0));
            const double pz = fma(r.i20, l2.x, fma(i22, l2.y, r.it2));
            int pi = -1;
#pragma unroll
            for (int t = 3; t >= 0; --t)
                if (t < gP && ((cand >> t) & 1u) && piece_contains(KIND, s_pieces[t], px, pz)) pi = t;
            int ci = 0, cj = 0;
            double hs = -CUDART_INF;
            if (pi >= 0) {
                if (safe) {
                    const DevPiece& c = s_pieces[pi];
                    hs = patch_sample<KIND, false>(c, g.data + c.data_offset, 0, c.pitch, px, pz, ci, cj);
                } else {
                    int vs, cell[3];
                    hs = ground_sample_piece(KIND, s_pieces, pi, g.data, px, pz, &vs, cell);
                    ci = cell[1]; cj = cell[2];
                }
            }
            out[s] = hs - r.it1;
            if (WANT_CELL && oc) { oc[3 * s] = pi; oc[3 * s + 1] = ci; oc[3 * s + 2] = cj; }
        }
    }
}

// k_f64_to_f32: single-precision copy of the policy state (the reference hands the state to Caffe as floats,
// learning/NeuralNet.cpp); round to nearest like the C++ double -> float conversion. Two elements per thread, grid-stride.
__global__ void __launch_bounds__(256) k_f64_to_f32(const double* __restrict__ src, float* __restrict__ dst, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < count; i += stride) {
        if (i + 1 < count && ((reinterpret_cast<size_t>(src + i) & 15) == 0) && ((reinterpret_cast<size_t>(dst + i) & 7) == 0)) {
            const double2 v = *reinterpret_cast<const double2*>(src + i);
            *reinterpret_cast<float2*>(dst + i) = make_float2((float)v.x, (float)v.y);
        } else {
            dst[i] = (float)src[i];
            if (i + 1 < count) dst[i + 1] = (float)src[i + 1];
        }
    }
}

// k_ee_mask: cRBDUtil::BuildEndEffectorJacobian (sim/RBDUtil.cpp:181-233) from the world Jacobian -- one thread per
// (env, dof): a column survives iff its joint lies on the chain root .. joint_id of that env.
__global__ void __launch_bounds__(256) k_ee_mask(const DevModel* __restrict__ M, int E, int n, const int* __restrict__ joint_ids,
                                                 int joint_id_all, double* __restrict__ J) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)E * n) return;
    const int e = (int)(tid / n), dof = (int)(tid - (long long)e * n);
    const int target = joint_ids ? joint_ids[e] : joint_id_all;
    const int jd = M->ix.dof_joint[dof];
    bool on_chain = false;
    if (target >= 0 && target < M->N)
        for (int j = target; j >= 0; j = M->ix.parent[j])
            if (j == jd) { on_chain = true; break; }
    if (!on_chain) {
        double* col = J + (size_t)e * 6 * n + dof;
#pragma unroll
        for (int r = 0; r < 6; ++r) col[(size_t)r * n] = 0.0;
    }
}

// k_exp_pd: explicit PD (cExpPDController::UpdateControlForce, sim/ExpPDController.cpp:59-66 -> cPDController::
// CalcJointTau*, sim/PDController.cpp:302-428) -- one thread per (env, joint); the joints are independent. A joint with a
// world-coordinate target (UseWorldCoord) composes its world rotation up its ancestor chain first.
__global__ void __launch_bounds__(128) k_exp_pd(const DevModel* __restrict__ M, int E, StepPtrs p) {
    __shared__ ModelIdx s_ix;
    stage_model_idx(&s_ix, M);
    const ModelIdx* X = &s_ix;
    const int N = M->N, n = M->n;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)E * N) return;
    const size_t e = (size_t)(tid / N);
    const int j = (int)(tid - (long long)e * N);
    if (!X->pd_valid[j]) return;
    if (p.joint_active && !p.joint_active[e * N + j]) return;  // PDController.cpp:137
    const int type = X->type[j], o = X->offset[j], sz = X->size[j];
    if (type != JT_REVOLUTE && type != JT_PRISMATIC && type != JT_SPHERICAL) return;  // planar (:358-364), fixed (:391-395)
    const double* q = p.pose + e * n;
    const double* qd = p.vel + e * n;
    const double* tq = p.tar_pose + e * n;
    const double* tqd = p.tar_vel + e * n;
    const double kp = p.kp ? p.kp[e * N + j] : M->kp[j];
    const double kd = p.kd ? p.kd[e * N + j] : M->kd[j];
    double* tau = p.tau + e * n;
    double Wm[9];
    const bool world = M->bx.use_world[j] != 0;
    if (world) {  // joint world rotation = product of the child -> parent rotations from the joint to the root
        double qj[7];
#pragma unroll
        for (int i = 0; i < 4; ++i) qj[i] = (i < sz) ? q[o + i] : 0.0;
        double t3[3];
        joint_local_trans(X, j, qj, M->attR[j], M->attT[j], Wm, t3);
        for (int a = X->parent[j]; a >= 0; a = X->parent[a]) {
            const int oa = X->offset[a], sa = (X->parent[a] == -1) ? 7 : X->size[a];
#pragma unroll
            for (int i = 0; i < 7; ++i) qj[i] = (i < sa) ? q[oa + i] : 0.0;
            double Ra[9], Rn[9];
            joint_local_trans(X, a, qj, M->attR[a], M->attT[a], Ra, t3);
            mat3_mul(Ra, Wm, Rn);
#pragma unroll
            for (int i = 0; i < 9; ++i) Wm[i] = Rn[i];
        }
    }
    if (type == JT_SPHERICAL) {
        double tar[4] = {tq[o], tq[o + 1], tq[o + 2], tq[o + 3]};
        const double qq[4] = {q[o], q[o + 1], q[o + 2], q[o + 3]};
        if (tar[0] * tar[0] + tar[1] * tar[1] + tar[2] * tar[2] + tar[3] * tar[3] == 0.0) tar[0] = 1.0;  // SetTargetTheta, :191-211
        else quat_normalize(tar);
        auto axis_angle = [](const double* v, double* axis, double& theta) {  // cMathUtil::QuaternionToAxisAngle (MathUtil.cpp:399-416)
            const double st = sqrt(v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
            axis[0] = 0; axis[1] = 0; axis[2] = 1;
            theta = 0.0;
            if (st > 0.0001) {
                const double sg = (double)((0.0 < v[0]) - (v[0] < 0.0));
                theta = 2 * sg * asin(st);
                axis[0] = v[1] / st; axis[1] = v[2] / st; axis[2] = v[3] / st;
            }
        };
        if (world) {  // tar = rel_rot * (world_rot^-1 * tar), PDController.cpp:254-267
            double ax3[3], th;
            axis_angle(qq, ax3, th);
            double sh, ch;
            sincos(th / 2, &sh, &ch);
            const double rel[4] = {ch, sh * ax3[0], sh * ax3[1], sh * ax3[2]};
            double wq[4];
            pb_rot_to_quat(Wm, wq);
            const double wc[4] = {wq[0], -wq[1], -wq[2], -wq[3]};
            double diff[4], t2[4];
            quat_mul(wc, tar, diff);
            quat_mul(rel, diff, t2);
#pragma unroll
            for (int i = 0; i < 4; ++i) tar[i] = t2[i];
        }
        const double qc[4] = {qq[0], -qq[1], -qq[2], -qq[3]};
        double qerr[4], axis[3], theta;
        quat_mul(qc, tar, qerr);  // q^-1 * tar, PDController.cpp:404-409
        axis_angle(qerr, axis, theta);
#pragma unroll
        for (int i = 0; i < 3; ++i) tau[o + i] += (kp * theta) * axis[i] + kd * (tqd[o + i] - qd[o + i]);
        tau[o + 3] += (kp * theta) * 0.0;
    } else {
        double tar0 = tq[o];
        if (world && type == JT_REVOLUTE) {  // PDController.cpp:238-252
            double axis_world[3], theta_world;
            pb_rot_to_axis_angle(Wm, axis_world, theta_world);
            const double dw = axis_world[0] * Wm[2] + axis_world[1] * Wm[5] + axis_world[2] * Wm[8];
            tar0 -= dw * theta_world - Wm[8] * q[o];
        }
        tau[o] += kp * (tar0 - q[o]) + kd * (tqd[o] - qd[o]);
    }
}

// ------------------------------------------------------------------------------------------------
// k_fp64_probe: register-resident DFMA chains, the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_probe(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chain live, never true in practice
}

}  // namespace

// Kernel timeline for diagnostics: enable allocates a [16][2] buffer of (first block start, last block end) in
// %globaltimer nanoseconds; read copies it out and re-arms it. Slots: 1 stable PD (k_pd_aba / k_pd_block), 4 k_kin_tile,
// 5 k_obs_env, 6 k_obs_samples / k_obs_gather. The buffer belongs to the device that was current at enable time.
static unsigned long long* g_trace_host_ptr = nullptr;
static void trace_rearm() {
    unsigned long long init[2 * kTraceSlots];
    for (int i = 0; i < kTraceSlots; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
    cudaMemcpy(g_trace_host_ptr, init, sizeof(init), cudaMemcpyHostToDevice);
}
int trl_trace_enable(int on) {
    cudaDeviceSynchronize();
    if (on && !g_trace_host_ptr) {
        if (cudaMalloc((void**)&g_trace_host_ptr, sizeof(unsigned long long) * 2 * kTraceSlots) != cudaSuccess) return -1;
        trace_rearm();
    }
    unsigned long long* v = on ? g_trace_host_ptr : nullptr;
    return (int)cudaMemcpyToSymbol(g_trace, &v, sizeof(v));
}
int trl_trace_read(unsigned long long* out32) {
    if (!g_trace_host_ptr) return -1;
    cudaDeviceSynchronize();
    cudaMemcpy(out32, g_trace_host_ptr, sizeof(unsigned long long) * 2 * kTraceSlots, cudaMemcpyDeviceToHost);
    trace_rearm();
    return 0;
}

// returns achieved FP64 TFLOP/s of a pure DFMA loop on the current device (<0: CUDA error)
double trl_fp64_probe_tflops(void* stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out = nullptr;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaStream_t st = (cudaStream_t)stream;
    k_fp64_probe<<<blocks, threads, 0, st>>>(out, 64, 0.999999, 1e-9);  // warm-up
    double best = -1.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        k_fp64_probe<<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------


static int ensure_smem(const void* func, size_t bytes) {
    // set once per (device, kernel, size): cudaFuncSetAttribute is kept out of graph captures. The attribute is per
    // device, so the cache is keyed on the current device as well (batches on several GPUs in one process).
    constexpr int kSeen = 256;
    static const void* seen_func[kSeen];
    static size_t seen_bytes[kSeen];
    static int seen_dev[kSeen];
    static int nseen = 0;
    static std::mutex mu;  // batches on different host threads share this cache
    if (bytes <= 48 * 1024) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < nseen; ++i)
        if (seen_func[i] == func && seen_dev[i] == dev && seen_bytes[i] >= bytes) return 0;
    cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) return (int)err;
    if (nseen < kSeen) { seen_func[nseen] = func; seen_bytes[nseen] = bytes; seen_dev[nseen] = dev; ++nseen; }
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int trl_launch_apply_tau(const DevModel* dm, const DevModel& hm, int E, const double* pose, const double* tau,
                         double* out_tau, double* out_body_wrench, void* stream) {
    const size_t smem = (size_t)hm.num_levels * 12 * kTile * sizeof(double);
    int rc = ensure_smem((const void*)k_apply_tau, smem);
    if (rc) return rc;
    const int tiles = (E + kTile - 1) / kTile;
    k_apply_tau<<<tiles, 32, smem, (cudaStream_t)stream>>>(dm, E, pose, tau, out_tau, out_body_wrench);
    return (int)cudaGetLastError();
}

int trl_launch_rbd_extras(const DevModel* dm, const DevModel& hm, int E, const double* pose, double* out_J,
                          double* out_Jcom, double* out_gforce, void* stream) {
    const int wpb = 1;
    const size_t smem = ((size_t)(kTreeHead + hm.num_levels * kXSlot) * kTile + (size_t)2 * kTile * ((6 * hm.n) | 1)) * sizeof(double) * wpb;
    if (smem > 227 * 1024 - 2048) return (int)cudaErrorInvalidValue;
    int rc = ensure_smem((const void*)k_rbd_extras, smem);
    if (rc) return rc;
    const int tiles = (E + kTile - 1) / kTile;
    k_rbd_extras<<<(tiles + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>(dm, E, pose, out_J, out_Jcom, out_gforce);
    return (int)cudaGetLastError();
}

// per env; the allocation is rounded up to whole tiles of kTile envs by the caller (trl_pd_scratch_tile())
size_t trl_pd_scratch_doubles(const DevModel& hm) {
    // k_pd_aba's hand-off between its two walks (+ f6 u3 kd1 per joint for its slim layout); k_pd_block keeps everything in
    // shared memory
    return (size_t)std::max(hm.bx.aba_slots, 1) + (size_t)10 * hm.N;
}
int trl_pd_scratch_tile() { return kTile; }

// k_pd_block: one block per tile of TE envs. TE = 32 unless the tile's state does not fit in shared memory; the number
// of warps per block keeps about 16 warps per SM (4 per sub-partition) across however many tiles share an SM.
// Returns 0, a cudaError_t, or -1 when the model does not fit (the caller falls back to the three-kernel chain).
static int launch_pd_block(const DevModel* dm, const DevModel& hm, int E, const StepPtrs& p, void* stream) {
    const PdBlkLayout L = pd_blk_layout(hm.N, hm.n, hm.nl, hm.bx.hcount);
    const size_t cst_bytes = (size_t)pd_cst_doubles(hm.N) * sizeof(double);
    const size_t kStatic = 2048, kBudget = 227 * 1024;
    int te = 0;
    size_t smem = 0;
    for (int cand = 32; cand >= 16 && !te; cand >>= 1) {
        smem = cst_bytes + (size_t)L.total * cand * sizeof(double);
        if (smem + kStatic <= kBudget) te = cand;
    }
    if (!te) return -1;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    const int tiles = (E + te - 1) / te;
    const int per_sm_fit = (int)(kBudget / (smem + kStatic + 1024));
    const int per_sm = std::max(1, std::min(per_sm_fit, (tiles + num_sms - 1) / num_sms));
    int W = 16 / per_sm;
    W = std::max(4, std::min(16, W));
    int rc;
    if (te == 32) {
        rc = ensure_smem((const void*)k_pd_block<32>, smem);
        if (rc) return rc;
        k_pd_block<32><<<tiles, W * 32, smem, (cudaStream_t)stream>>>(dm, E, p);
    } else {
        rc = ensure_smem((const void*)k_pd_block<16>, smem);
        if (rc) return rc;
        k_pd_block<16><<<tiles, W * 32, smem, (cudaStream_t)stream>>>(dm, E, p);
    }
    return (int)cudaGetLastError();
}

// k_pd_aba: lane = env, one warp per branch of the tree, one 32-env tile per block. Returns 0, a cudaError_t, or -1
// when the kernel does not apply (M / C requested, a shape-less body, or a tree too deep for the workspace).
static int launch_pd_aba(const DevModel* dm, const DevModel& hm, int E, const StepPtrs& p, double* scratch, void* stream) {
    (void)dm;
    if (p.out_mass || p.out_bias || !hm.bx.all_valid || !scratch || hm.N < 2) return -1;
    const int nb = hm.bx.num_branches;
    const int W = std::min(nb, 8);  // one warp per branch
    // per-warp area: the branch's level slots (22 doubles per level, 12 in the slim layout) + the blocks of its joints with
    // several children
    int area_full = 0, area_slim = 0;
    for (int b = 0; b < nb; ++b) {
        int extra = 0;
        for (int q = hm.bx.b_pre0[b]; q < hm.bx.b_pre1[b]; ++q) extra += hm.bx.nchild[hm.bx.preorder[q]] >= 2 ? 1 : 0;
        area_full = std::max(area_full, 22 * hm.bx.b_depth[b] + 39 * extra);
        area_slim = std::max(area_slim, 12 * hm.bx.b_depth[b] + 39 * extra);
    }
    const size_t smem_full = (size_t)aba_ws_elems(nb, area_full, W) * kTile * sizeof(double);
    const size_t smem_slim = (size_t)aba_ws_elems(nb, area_slim, W) * kTile * sizeof(double);
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    const int tiles = (E + kTile - 1) / kTile;
    const int by_regs = std::max(1, 65536 / (W * 32 * 256));
    auto per_sm = [&](size_t smem) { return std::min(by_regs, (int)((227 * 1024) / (smem + 1280))); };
    auto waves = [&](size_t smem) { const int ps = per_sm(smem); return ps > 0 ? (tiles + num_sms * ps - 1) / (num_sms * ps) : 1 << 20; };
    static const int slim_env = env_int("TRL_PD_SLIM", -1);  // A/B hook: 0 = never, 1 = always
    // The slim layout pays a trip to L2 per joint. It wins when it saves a whole wave of tiles (raptor x 8192: PD entry 39.3
    // -> 25.0 us), when the full one does not fit, and when the resident tiles of the full layout leave less than ~48 KB of
    // the SM's 256 KB to L1 (biped3D x 16384: four 54 KB tiles per SM; the walk's strided input rows and the hand-off then
    // miss L1: 27.6 -> 24.2 us with 39 KB tiles). Otherwise the full layout is faster (dog x 4096: 42.7 vs 43.4 us).
    const bool fits_full = smem_full <= 227 * 1024 - 3072;
    const int resident_full = std::min(per_sm(smem_full), (tiles + num_sms - 1) / num_sms);
    const bool l1_squeezed = (size_t)resident_full * (smem_full + 1280) > 208 * 1024 &&
                             (size_t)resident_full * (smem_full - smem_slim) >= 32 * 1024;
    const bool slim = slim_env >= 0 ? slim_env != 0 : (!fits_full || waves(smem_slim) < waves(smem_full) || l1_squeezed);
    const size_t smem = slim ? smem_slim : smem_full;
    if (smem > 227 * 1024 - 3072) return -1;
    const int stride = (int)trl_pd_scratch_doubles(hm), fuk_off = std::max(hm.bx.aba_slots, 1);
    // kernel parameter: index tables + per-joint constants of the (host copy of the) model
    auto launch = [&](auto nj_tag, auto slim_tag) -> int {
        constexpr int NJ = decltype(nj_tag)::value;
        constexpr bool SLIM = decltype(slim_tag)::value;
        int rc = ensure_smem((const void*)k_pd_aba<NJ, SLIM>, smem);
        if (rc) return rc;
        AbaTables<NJ> tabs;
        tabs.ix = hm.ix;
        tabs.bx = hm.bx;
        memcpy(tabs.cst, hm.pd_cst, sizeof(tabs.cst));
        k_pd_aba<NJ, SLIM><<<tiles, 32 * W, smem, (cudaStream_t)stream>>>(tabs, hm.gravity[0], hm.gravity[1], hm.gravity[2], E, p, scratch,
                                                                         stride, fuk_off, SLIM ? area_slim : area_full, hm.N, hm.n);
        return (int)cudaGetLastError();
    };
    using NjS = std::integral_constant<int, 12>;
    using NjL = std::integral_constant<int, TRL_MAX_JOINTS>;
    if (hm.N <= 12) return slim ? launch(NjS{}, std::true_type{}) : launch(NjS{}, std::false_type{});
    return slim ? launch(NjL{}, std::true_type{}) : launch(NjL{}, std::false_type{});
}

// Returns the number of kernels launched (>0) or a negative cudaError.
int trl_launch_imp_pd(const DevModel* dm, const DevModel& hm, int E, const StepPtrs& p, double* scratch, void* stream) {
    // k_pd_aba (articulated-body elimination, no M) when every body is valid and neither M nor C is asked for; else
    // k_pd_block (CRBA + RNEA + branch-sparse LTDL, M / C outputs, shape-less bodies). Returns the number of kernels
    // launched (> 0) or a negative cudaError_t.
    static const int aba_env = env_int("TRL_PD_ABA", 1);  // A/B hook: 0 = k_pd_block only (the tests run both)
    if (aba_env) {
        const int rc = launch_pd_aba(dm, hm, E, p, scratch, stream);
        if (rc == 0) return 1;
        if (rc > 0) return -rc;
    }
    const int rc = launch_pd_block(dm, hm, E, p, stream);
    if (rc == 0) return 1;
    return rc > 0 ? -rc : -(int)cudaErrorInvalidConfiguration;  // -1: the model's tile does not fit in shared memory
}

int trl_launch_fk(const DevModel* dm, const DevModel& hm, int E, const double* pose, double* out_joint_world,
                  double* out_body_pos, void* stream) {
    const size_t smem = (size_t)(hm.n + 24 * hm.N) * sizeof(double) * kWarpsPerBlock;
    const int blocks = (E + kWarpsPerBlock - 1) / kWarpsPerBlock;
    k_fk<<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(dm, E, pose, out_joint_world, out_body_pos);
    return (int)cudaGetLastError();
}

int trl_launch_kin_char(const DevModel* dm, const DevModel& hm, const double* frames, int E, const double* time,
                        const double* origin_pos, const double* origin_rot, double* out_pose, double* out_vel,
                        void* stream) {
    const long long total = (long long)E * hm.N;
    const int blocks = (int)((total + 127) / 128);
    k_kin_char<<<blocks, 128, 0, (cudaStream_t)stream>>>(dm, frames, E, time, origin_pos, origin_rot, out_pose, out_vel);
    return (int)cudaGetLastError();
}

int trl_launch_kin(const DevModel* dm, const DevModel& hm, const double* frames, int E, const double* time,
                   const double* origin_pos, const double* origin_rot, double* out_pose, double* out_vel,
                   double* out_body_pos, void* stream) {
    // warp per (tile of 32 envs, joint): local transforms [N][12][32] in shared memory
    const size_t smem = (size_t)hm.N * 12 * kTile * sizeof(double);
    int rc = ensure_smem((const void*)k_kin_tile, smem);
    if (rc) return rc;
    const int tiles = (E + kTile - 1) / kTile;
    static const int kin_warps_env = env_int("TRL_KIN_WARPS", 0);  // A/B hook: warps per tile (joints are looped over)
    // A batch of more than one tile per SM shares the SMs with the stable-PD tiles (two to four of them hold 50 - 100 % of
    // an SM's registers): blocks of half the joints' warps (two joints per warp) still find room beside them (biped3D x
    // 16384: step 0.0923 -> 0.0890 ms; raptor x 8192: 68.7 -> 62.6 us).
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    const int kin_auto = (tiles > num_sms) ? (hm.N + 1) / 2 : kKinWarps;
    const int kin_warps = kin_warps_env > 0 ? std::min(kin_warps_env, kKinWarps) : std::min(kin_auto, kKinWarps);
    k_kin_tile<<<tiles, std::min(hm.N, kin_warps) * 32, smem, (cudaStream_t)stream>>>(dm, hm.ix, frames, E, time, origin_pos, origin_rot, out_pose,
                                                                out_vel, out_body_pos);
    return (int)cudaGetLastError();
}

template <int G>
static int launch_reward_t(const DevModel* dm, const DevModel& hm, const DevGround& g, int E, const double* pose0,
                           const double* vel0, const double* pose1, const double* vel1, const double* body_vel,
                           const unsigned char* fallen, double* out_reward, double* out_terms, void* stream) {
    constexpr int EPW = 32 / G;
    const int wpb = 4;
    const size_t smem = (size_t)42 * hm.N * sizeof(double) * wpb * EPW;
    int rc = ensure_smem((const void*)k_imitate_reward<G>, smem);
    if (rc) return rc;
    const int epb = wpb * EPW;
    k_imitate_reward<G><<<(E + epb - 1) / epb, wpb * 32, smem, (cudaStream_t)stream>>>(
        dm, g, E, pose0, vel0, pose1, vel1, body_vel, fallen, out_reward, out_terms);
    return (int)cudaGetLastError();
}

int trl_launch_imitate_reward(const DevModel* dm, const DevModel& hm, const DevGround& g, int E, const double* pose0,
                              const double* vel0, const double* pose1, const double* vel1, const double* body_vel,
                              const unsigned char* fallen, double* out_reward, double* out_terms, void* stream) {
    if (hm.N <= 8) return launch_reward_t<8>(dm, hm, g, E, pose0, vel0, pose1, vel1, body_vel, fallen, out_reward, out_terms, stream);
    if (hm.N <= 16) return launch_reward_t<16>(dm, hm, g, E, pose0, vel0, pose1, vel1, body_vel, fallen, out_reward, out_terms, stream);
    return launch_reward_t<32>(dm, hm, g, E, pose0, vel0, pose1, vel1, body_vel, fallen, out_reward, out_terms, stream);
}

int trl_launch_sample_height(const DevGround& g, int E, const double* pos, int S, double* out_h,
                             unsigned char* out_valid, int* out_cell, void* stream) {
    const long long total = (long long)E * S;
    const int blocks = (int)((total + 255) / 256);
    k_sample_height<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, E, pos, S, out_h, out_valid, out_cell);
    return (int)cudaGetLastError();
}

int trl_launch_poli_state(const DevModel* dm, const DevModel& hm, const DevGround& g_in, const trl_obs_cfg& cfg, int F,
                          int E, const double* pose, const double* vel, const double* body_pos,
                          const double* body_vel, const double* body_quat, const double* body_angvel,
                          const double* ground_vel, const double* phase, const unsigned char* flip,
                          const double* sample_local, void* env_rec, double* out_state, int* out_cell, void* stream,
                          void* side_stream, void* ev_fork, void* ev_join) {
    // staged rows (one contiguous row per env) win on the 2-D heightfields, direct gathers on the 3-D ones (profiles/README.md)
    static const int staged_env = env_int("TRL_OBS_STAGED", -1);  // A/B hook: 0 = always gather, 1 = always stage
    const bool use_staged = g_in.patch_nbuf > 0 && (staged_env >= 0 ? staged_env != 0 : g_in.kind == 1);
    DevGround g = g_in;
    if (!use_staged) g.patch_nbuf = 0;  // k_obs_env skips the cell-rectangle mapping
    // record (4 lanes per env) -> body features (thread per (env, body)) -> ground samples
    k_obs_rec<<<(int)(((long long)E * kRecG + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dm, g, cfg, E, pose, flip, sample_local,
                                                                                          (ObsEnvRec*)env_rec, out_state);
    // the body features only need the record: with a side stream they run beside the ground samples instead of before them
    const bool side = side_stream && ev_fork && ev_join && cfg.num_samples > 0;
    cudaStream_t fs = side ? (cudaStream_t)side_stream : (cudaStream_t)stream;
    if (side) {
        if (cudaEventRecord((cudaEvent_t)ev_fork, (cudaStream_t)stream) != cudaSuccess) return (int)cudaGetLastError();
        if (cudaStreamWaitEvent(fs, (cudaEvent_t)ev_fork, 0) != cudaSuccess) return (int)cudaGetLastError();
    }
    k_obs_feat<<<(int)(((long long)E * hm.N + 127) / 128), 128, 0, fs>>>(
        dm, cfg, E, pose, vel, body_pos, body_vel, body_quat, body_angvel, ground_vel, phase, (const ObsEnvRec*)env_rec, out_state);
    if (side && cudaEventRecord((cudaEvent_t)ev_join, fs) != cudaSuccess) return (int)cudaGetLastError();
    int rc = (int)cudaGetLastError();
    if (rc || cfg.num_samples <= 0) return rc;
    const ObsEnvRec* er = (const ObsEnvRec*)env_rec;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (g.kind == 1 || g.kind == 2) ? (size_t)kObsWarps * obs_warp_bytes(g) : 0;
    const int grid = (E + kObsWarps - 1) / kObsWarps;
#define TRL_OBS_LAUNCH(K)                                                                                          \
    if (out_cell) {                                                                                                \
        rc = ensure_smem((const void*)k_obs_samples<K, true>, smem);                                               \
        if (rc) return rc;                                                                                         \
        k_obs_samples<K, true><<<grid, kObsWarps * 32, smem, st>>>(g, cfg, F, E, er, sample_local, out_state, out_cell);  \
    } else {                                                                                                       \
        rc = ensure_smem((const void*)k_obs_samples<K, false>, smem);                                              \
        if (rc) return rc;                                                                                         \
        k_obs_samples<K, false><<<grid, kObsWarps * 32, smem, st>>>(g, cfg, F, E, er, sample_local, out_state, out_cell); \
    }
#define TRL_OBS_GATHER(K)                                                                                          \
    if (out_cell) k_obs_gather<K, true><<<E, gtpb, 0, st>>>(g, cfg, F, E, er, sample_local, out_state, out_cell);   \
    else k_obs_gather<K, false><<<E, gtpb, 0, st>>>(g, cfg, F, E, er, sample_local, out_state, out_cell);
    const int gtpb = std::min(256, std::max(32, ((cfg.num_samples + kObsSPT - 1) / kObsSPT + 31) / 32 * 32));
    if (g.kind == 1) { if (use_staged) { TRL_OBS_LAUNCH(1) } else { TRL_OBS_GATHER(1) } }
    else if (g.kind == 2) { if (use_staged) { TRL_OBS_LAUNCH(2) } else { TRL_OBS_GATHER(2) } }
    else if (g.kind == 3) { TRL_OBS_LAUNCH(3) }
    else { TRL_OBS_LAUNCH(0) }
#undef TRL_OBS_GATHER
#undef TRL_OBS_LAUNCH
    rc = (int)cudaGetLastError();
    if (!rc && side && cudaStreamWaitEvent(st, (cudaEvent_t)ev_join, 0) != cudaSuccess) rc = (int)cudaGetLastError();
    return rc;
}

int trl_launch_exp_pd(const DevModel* dm, const DevModel& hm, int E, const StepPtrs& p, void* stream) {
    const long long total = (long long)E * hm.N;
    k_exp_pd<<<(int)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dm, E, p);
    return (int)cudaGetLastError();
}

int trl_launch_ee_mask(const DevModel* dm, int E, int n, const int* joint_ids, int joint_id_all, double* J, void* stream) {
    const long long total = (long long)E * n;
    k_ee_mask<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dm, E, n, joint_ids, joint_id_all, J);
    return (int)cudaGetLastError();
}

size_t trl_obs_env_rec_bytes() { return sizeof(ObsEnvRec); }

int trl_launch_f64_to_f32(const double* src, float* dst, size_t count, void* stream) {
    if (count == 0) return 0;
    const int blocks = (int)std::min<size_t>((count / 2 + 255) / 256, (size_t)148 * 16);
    k_f64_to_f32<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, count);
    return (int)cudaGetLastError();
}

int trl_obs_state_size(int N, const trl_obs_cfg& cfg) { return obs_dims(N, cfg).F; }

#ifdef TRL_PD_CLK
// diagnostic build only (tools/pd_phases.py): clock64 stamps of one block of k_pd_aba, [warp][stamp]
extern "C" int trl_debug_pd_clk(long long* out256) {
    return (int)cudaMemcpyFromSymbol(out256, g_pd_clk, sizeof(long long) * 8 * 32);
}
#endif
