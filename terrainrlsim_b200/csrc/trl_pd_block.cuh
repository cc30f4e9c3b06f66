// k_pd_block: stable-PD torque (cRBDModel::Update + cImpPDController::CalcControlForces, sim/RBDModel.cpp:39-55,
// sim/RBDUtil.cpp:4-179,931-935, sim/ImpPDController.cpp:137-196) as ONE kernel in which a block of W warps works on
// one tile of TE environments. Included by trl_kernels.cu (inside its anonymous namespace).
//
// Why: with lane = env (k_pd_tree / k_pd_h / k_pd_solve) 16384 envs are 512 warps on 592 SM sub-partitions and every
// warp is a 20k-instruction dependent chain -- the chain is bound by single-warp latency, not by a pipe (ncu:
// 0.16 eligible warps per cycle). Here the lanes still are the envs (uniform control flow, conflict-free [element][lane]
// shared memory, no shuffles), but the tile's work is cut into tasks that different warps run concurrently:
//   A  per joint   : PD error terms, joint-local transform                                   (N tasks)
//   B  per level   : root-frame transform, spatial velocity / bias acceleration (parent -> child), overlapped with
//                    the body inertia / body force tasks of the previous level and the world-coordinate PD targets
//   C  per level   : subtree sums of inertia and force (children -> parent)
//   D  per column  : S_c, F_c = I_c S_c, bias force, H row against the column's ancestor chain    (nl tasks)
//   E  per column k, last to first: one step of the branch-sparse L^T D L factorisation of H = M + dt Kd (Featherstone,
//                    "Efficient factorization of the joint-space inertia matrix for branched kinematic trees"), its
//                    rank-1 update pairs spread over the warps; the forward substitution of the rhs rides along
//   F  per column-depth: back substitution
//   G  coalesced write of tau
// The whole state of the tile (transforms, subtree inertias, the sparse H) stays in shared memory: no scratch hand-off
// through L2 (the three-kernel chain moved 4.1 KB per env through it).
#pragma once

// element `idx` of region `base` for this lane's env
#define PB(base, idx) ws[((base) + (idx)) * TE + ln]

struct PdBlkLayout {
    int T;      // [N][12] root-frame transform of the joint [R(9) | t(3)]  (task A leaves the joint-local one here)
    int IF;     // [N][16] inertia (m, h(3), Ixx Ixy Ixz Iyy Iyz Izz) + force (6); subtree sums after phase C
    int H;      // [hcount] sparse H rows; phases A-B: [N][12] spatial velocity / acceleration (first 6 of a slot: qd stash)
    int rhs;    // [nl]
    int t0;     // [n]  Kp e_p + Kd e_v per dof (tau = t0 - dt Kd_masked qdd)
    int kdm;    // [N]  masked Kd
    int kd;     // [N]  unmasked Kd
    int total;  // elements (x TE doubles)
};

__host__ __device__ inline PdBlkLayout pd_blk_layout(int N, int n, int nl, int hcount) {
    PdBlkLayout L;
    int o = 0;
    L.T = o; o += 12 * N;
    L.IF = o; o += 16 * N;
    L.H = o; o += (hcount > 12 * N ? hcount : 12 * N);
    L.rhs = o; o += nl;
    L.t0 = o; o += n;
    L.kdm = o; o += N;
    L.kd = o; o += N;
    L.total = o;
    return L;
}

template <int TE>
TRL_DEV void pb_subspace(int kind, int a, const double* Tj /* &PB(T + 12 j, 0) */, const double* Rq9, double* s) {
    if (kind == CK_ANG) {
        s[0] = Tj[a * TE]; s[1] = Tj[(3 + a) * TE]; s[2] = Tj[(6 + a) * TE];
        const double t0 = Tj[9 * TE], t1 = Tj[10 * TE], t2 = Tj[11 * TE];
        s[3] = t1 * s[2] - t2 * s[1];
        s[4] = t2 * s[0] - t0 * s[2];
        s[5] = t0 * s[1] - t1 * s[0];
    } else if (kind == CK_LIN) {
        s[0] = s[1] = s[2] = 0.0;
        s[3] = Tj[a * TE]; s[4] = Tj[(3 + a) * TE]; s[5] = Tj[(6 + a) * TE];
    } else {  // CK_ROOT_LIN: row a of R(q_root) (cRBDUtil::BuildJointSubspaceRoot, sim/RBDUtil.cpp:776-783)
        s[0] = s[1] = s[2] = 0.0;
        s[3] = Rq9[3 * a]; s[4] = Rq9[3 * a + 1]; s[5] = Rq9[3 * a + 2];
    }
}

// cMathUtil::RotMatToAxisAngle (util/MathUtil.cpp:239-260) of a row-major 3x3
TRL_DEV void pb_rot_to_axis_angle(const double* R, double* axis, double& theta) {
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = clampd(c, -1.0, 1.0);
    theta = acos(c);
    if (fabs(theta) < 0.00001) {
        axis[0] = 0; axis[1] = 0; axis[2] = 1;
    } else {
        const double m21 = R[7] - R[5], m02 = R[2] - R[6], m10 = R[3] - R[1];
        const double denom = sqrt(m21 * m21 + m02 * m02 + m10 * m10);
        axis[0] = m21 / denom; axis[1] = m02 / denom; axis[2] = m10 / denom;
    }
}

// cMathUtil::RotMatToQuaternion (util/MathUtil.cpp:272-307), (w,x,y,z)
TRL_DEV void pb_rot_to_quat(const double* R, double* q) {
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        const double S = sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * S; q[1] = (R[7] - R[5]) / S; q[2] = (R[2] - R[6]) / S; q[3] = (R[3] - R[1]) / S;
    } else if ((R[0] > R[4]) && (R[0] > R[8])) {
        const double S = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[0] = (R[7] - R[5]) / S; q[1] = 0.25 * S; q[2] = (R[1] + R[3]) / S; q[3] = (R[2] + R[6]) / S;
    } else if (R[4] > R[8]) {
        const double S = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[0] = (R[2] - R[6]) / S; q[1] = (R[1] + R[3]) / S; q[2] = 0.25 * S; q[3] = (R[5] + R[7]) / S;
    } else {
        const double S = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[0] = (R[3] - R[1]) / S; q[1] = (R[2] + R[6]) / S; q[2] = (R[5] + R[7]) / S; q[3] = 0.25 * S;
    }
}

// PD error terms of one joint (cImpPDController::CalcControlForces, sim/ImpPDController.cpp:137-196, with the target
// post-processing of cPDController::SetTargetTheta / SetTargetVel, sim/PDController.cpp:191-211,430-461). `Wm` is the
// joint's world rotation when its target is a world-coordinate one (cPDController::GetTargetTheta, :223-273), else null.
template <int TE>
TRL_DEV void pb_pd_terms(const ModelIdx* X, const double* cj_, int j, int N, int n, size_t e, int ln, const StepPtrs& p,
                         double* ws, const PdBlkLayout& L, const double* Wm, bool env_ok) {
    (void)env_ok;
    const int o = X->offset[j], sz = X->size[j], type = X->type[j];
    const bool is_root = X->parent[j] == -1;
    const double dt = p.dt;
    const double* q = p.pose + e * n;
    const double* qd = p.vel + e * n;
    const double* tq = p.tar_pose + e * n;
    const double* tqd = p.tar_vel + e * n;
    const bool valid = X->pd_valid[j] != 0;
    const bool active = valid && (p.joint_active ? (p.joint_active[e * N + j] != 0) : true);
    double kp = 0.0, kd = 0.0;
    if (valid) {
        kp = p.kp ? p.kp[e * N + j] : cj_[19];
        kd = p.kd ? p.kd[e * N + j] : cj_[20];
    }
    const double kpm = active ? kp : 0.0, kdm = active ? kd : 0.0;
    PB(L.kd, j) = kd;  // unmasked gains go on the diagonal
    PB(L.kdm, j) = kdm;
    double ep[4] = {0, 0, 0, 0};
    const bool sph = valid && !is_root && type == JT_SPHERICAL;
    double tar0 = 0.0;  // revolute / prismatic target after the world-coordinate conversion
    if (valid && !sph && sz > 0) tar0 = tq[o];
    if (sph) {
        double qj[4], wj[4], dq[4], qi[4], tar[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qj[i] = q[o + i]; wj[i] = qd[o + i]; tar[i] = tq[o + i]; }
        quat_diff_mul(qj, wj, dq);
#pragma unroll
        for (int i = 0; i < 4; ++i) qi[i] = qj[i] + dt * dq[i];
        quat_normalize(qi);
        if (tar[0] * tar[0] + tar[1] * tar[1] + tar[2] * tar[2] + tar[3] * tar[3] == 0.0) tar[0] = 1.0;
        else quat_normalize(tar);
        if (Wm) {  // tar = rel_rot * (world_rot^-1 * tar), PDController.cpp:254-267
            double st = sqrt(qj[1] * qj[1] + qj[2] * qj[2] + qj[3] * qj[3]);
            double ax[3] = {0, 0, 1}, th = 0.0;
            if (st > 0.0001) {  // cMathUtil::QuaternionToAxisAngle, util/MathUtil.cpp:399-416
                const double sg = (double)((0.0 < qj[0]) - (qj[0] < 0.0));
                th = 2 * sg * asin(st);
                ax[0] = qj[1] / st; ax[1] = qj[2] / st; ax[2] = qj[3] / st;
            }
            double sh, ch;
            sincos(th / 2, &sh, &ch);
            const double rel[4] = {ch, sh * ax[0], sh * ax[1], sh * ax[2]};  // AxisAngleToQuaternion
            double wq[4];
            pb_rot_to_quat(Wm, wq);
            const double wc[4] = {wq[0], -wq[1], -wq[2], -wq[3]};
            double diff[4], t2[4];
            quat_mul(wc, tar, diff);
            quat_mul(rel, diff, t2);
#pragma unroll
            for (int i = 0; i < 4; ++i) tar[i] = t2[i];
        }
        quat_vel_rel(qi, tar, 1.0, ep);
    } else if (Wm && valid && type == JT_REVOLUTE) {  // PDController.cpp:238-252
        double axis_world[3], theta_world;
        pb_rot_to_axis_angle(Wm, axis_world, theta_world);
        const double ar0 = Wm[2], ar1 = Wm[5], ar2 = Wm[8];  // cJoint::CalcAxisWorld: world rotation * z
        const double dw = axis_world[0] * ar0 + axis_world[1] * ar1 + axis_world[2] * ar2;
        const double dr = ar2;  // axis_rel = z
        tar0 -= dw * theta_world - dr * q[o];
    }
    for (int i = 0; i < sz; ++i) {
        const double qdi = qd[o + i];
        double epi = 0.0, evi_ = 0.0;
        if (valid) {
            if (sph) {
                epi = (i == 0) ? ep[0] : (i == 1) ? ep[1] : (i == 2) ? ep[2] : ep[3];
                evi_ = ((i < 3) ? tqd[o + i] : 0.0) - qdi;
            } else {
                const double ti = (i == 0) ? tar0 : tq[o + i];
                epi = (ti - (q[o + i] + dt * qdi)) / 1.0;
                evi_ = tqd[o + i] - qdi;
            }
        }
        const double u = kpm * epi + kdm * evi_;
        double t0 = u;
        if (X->dof_col[o + i] < 0) {  // dead slot: decoupled 1x1 system (0 + dt*Kd) qdd = rhs, zero pivot => 0
            const double dg = dt * kd;
            const double acc = (fabs(dg) > DBL_MIN) ? u / dg : 0.0;
            t0 = kpm * epi + kdm * (evi_ - dt * acc);
        }
        PB(L.t0, o + i) = t0;
    }
}

template <int TE>
__global__ void __launch_bounds__(512)
k_pd_block(const DevModel* __restrict__ M, int E, StepPtrs p) {
    KTrace trace_(TR_PD_TREE);
    extern __shared__ double smem[];
    __shared__ ModelIdx s_ix;
    __shared__ PdBlkIdx s_bx;
    __shared__ int s_status[32];
    {
        const int cnt = M->N * kCst;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) smem[i] = __ldg(M->pd_cst + i);
        const uint4* src = reinterpret_cast<const uint4*>(&M->bx);
        uint4* d = reinterpret_cast<uint4*>(&s_bx);
        for (int i = threadIdx.x; i < (int)(sizeof(PdBlkIdx) / 16); i += blockDim.x) d[i] = __ldg(src + i);
        if (threadIdx.x < 32) s_status[threadIdx.x] = 0;
    }
    stage_model_idx(&s_ix, M);  // contains the barrier
    const ModelIdx* X = &s_ix;
    const PdBlkIdx* B = &s_bx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int ln = lane % TE;  // TE < 32: the upper lanes mirror the lower ones (same values, same addresses)
    const int e0 = blockIdx.x * TE;
    const int envs = min(TE, E - e0);
    const bool env_ok = ln < envs;
    const size_t e = (size_t)(e0 + (env_ok ? ln : envs - 1));  // tail lanes redo the last env
    const int N = M->N, n = M->n, nl = M->nl;
    const PdBlkLayout L = pd_blk_layout(N, n, nl, B->hcount);
    const double* cst = smem;
    double* ws = smem + pd_cst_doubles(N);
    const double dt = p.dt;
    const double* q = p.pose + e * n;
    const double* qd = p.vel + e * n;
    const int num_levels = M->num_levels;
    const int VA = L.H;  // spatial velocity / acceleration live in the H region until phase D
    if (p.out_mass) {  // rare path (GetMassMat): zero the tile's [envs][n][n] block, coalesced
        double* om = p.out_mass + (size_t)e0 * n * n;
        for (int idx = threadIdx.x; idx < envs * n * n; idx += blockDim.x) om[idx] = 0.0;
        __syncthreads();
    }

    // R(q_root) and the world <- root rotation, needed by a few tasks only: recomputed where used
    auto root_rot = [&](double* rq9, double* Rw) {
        quat_to_mat(q[3], q[4], q[5], q[6], rq9);
        if (Rw) mat3_mul(cst, rq9, Rw);
    };

    // ---------------- phase A: per joint -- PD terms, joint-local transform; the root's velocity / acceleration ----------------
    for (int j = warp; j < N; j += W) {
        const double* cj_ = cst + kCst * j;
        const int o = X->offset[j], sz = X->size[j];
        const bool is_root = X->parent[j] == -1;
        if (!B->use_world[j]) pb_pd_terms<TE>(X, cj_, j, N, n, e, ln, p, ws, L, nullptr, env_ok);
        if (is_root) {
            double wr[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) wr[i] = qd[o + i];
            double rq9[9], Rw[9];
            root_rot(rq9, Rw);
            // a_0 = X_{world->root}(0, -g): E = Rw^T  (RBDUtil.cpp:17-18,59-60)
            const double ng[3] = {-M->gravity[0], -M->gravity[1], -M->gravity[2]};
            double g0[3];
            mat3T_mulv(Rw, ng, g0);
            // c_root (cRBDUtil::BuildCjRoot, RBDUtil.cpp:867-904)
            double dq[4];
            const double qr4[4] = {q[3], q[4], q[5], q[6]};
            quat_diff_mul(qr4, wr + 3, dq);
            const double qw = qr4[0], qx = qr4[1], qy = qr4[2], qz = qr4[3];
            const double dw = dq[0], dx = dq[1], dy = dq[2], dz = dq[3];
            const double m00 = 4 * (qw * dw + qx * dx), m11 = 4 * (qw * dw + qy * dy), m22 = 4 * (qw * dw + qz * dz);
            const double m10 = 2 * (dx * qy + qx * dy - dw * qz - qw * dz);
            const double m01 = 2 * (dx * qy + qx * dy + dw * qz + qw * dz);
            const double m20 = 2 * (dx * qz + qx * dz + dw * qy + qw * dy);
            const double m02 = 2 * (dx * qz + qx * dz - dw * qy - qw * dy);
            const double m21 = 2 * (dy * qz + qy * dz - dw * qx - qw * dx);
            const double m12 = 2 * (dy * qz + qy * dz + dw * qx + qw * dx);
#pragma unroll
            for (int i = 0; i < 12; ++i) PB(L.T + 12 * j, i) = (i < 9 && i % 4 == 0) ? 1.0 : 0.0;
            // V_root = S qd: omega = qd[3..5] (root frame), v = R(q)^T-rows . qd[0..2]  (S = [0 | I ; Rq rows | 0])
            double V[6], A[6];
            V[0] = wr[3]; V[1] = wr[4]; V[2] = wr[5];
            V[3] = rq9[0] * wr[0] + rq9[3] * wr[1] + rq9[6] * wr[2];
            V[4] = rq9[1] * wr[0] + rq9[4] * wr[1] + rq9[7] * wr[2];
            V[5] = rq9[2] * wr[0] + rq9[5] * wr[1] + rq9[8] * wr[2];
            A[0] = A[1] = A[2] = 0.0;
            A[3] = g0[0] + (m00 * wr[0] + m01 * wr[1] + m02 * wr[2]);
            A[4] = g0[1] + (m10 * wr[0] + m11 * wr[1] + m12 * wr[2]);
            A[5] = g0[2] + (m20 * wr[0] + m21 * wr[1] + m22 * wr[2]);
#pragma unroll
            for (int k = 0; k < 6; ++k) { PB(VA + 12 * j, k) = V[k]; PB(VA + 12 * j, 6 + k) = A[k]; }
        } else {
            double qj[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qj[i] = (i < sz) ? q[o + i] : 0.0;
            double R[9], t[3];
            joint_local_trans(X, j, qj, cj_, cj_ + 9, R, t);
#pragma unroll
            for (int i = 0; i < 9; ++i) PB(L.T + 12 * j, i) = R[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) PB(L.T + 12 * j, 9 + i) = t[i];
            // the joint's live velocities, stashed where its V / A go (read back by its phase-B task)
            const int c0 = X->first_col[j], nc = X->ncol[j];
            for (int c = 0; c < nc; ++c) PB(VA + 12 * j, c) = qd[X->col_dof[c0 + c]];
        }
    }
    __syncthreads();

    // body inertia in the root frame (m, h = m c, I about the frame origin) and f = I a + V x* (I V) of joint j
    auto body_task = [&](int j) {
        const double* cj_ = cst + kCst * j;
        double T[12], Vj[6], Aj[6];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = PB(L.T + 12 * j, i);
#pragma unroll
        for (int k = 0; k < 6; ++k) { Vj[k] = PB(VA + 12 * j, k); Aj[k] = PB(VA + 12 * j, 6 + k); }
        double I[10], f[6] = {0, 0, 0, 0, 0, 0};
        if (X->valid_body[j]) {
            const double m = cj_[12];
            double c[3];
            mat3_mulv(T, cj_ + 13, c);
            c[0] += T[9]; c[1] += T[10]; c[2] += T[11];
            const double d0 = cj_[16], d1 = cj_[17], d2 = cj_[18];
            double Ixx = T[0] * T[0] * d0 + T[1] * T[1] * d1 + T[2] * T[2] * d2;
            double Ixy = T[0] * T[3] * d0 + T[1] * T[4] * d1 + T[2] * T[5] * d2;
            double Ixz = T[0] * T[6] * d0 + T[1] * T[7] * d1 + T[2] * T[8] * d2;
            double Iyy = T[3] * T[3] * d0 + T[4] * T[4] * d1 + T[5] * T[5] * d2;
            double Iyz = T[3] * T[6] * d0 + T[4] * T[7] * d1 + T[5] * T[8] * d2;
            double Izz = T[6] * T[6] * d0 + T[7] * T[7] * d1 + T[8] * T[8] * d2;
            Ixx += m * (c[1] * c[1] + c[2] * c[2]);
            Iyy += m * (c[0] * c[0] + c[2] * c[2]);
            Izz += m * (c[0] * c[0] + c[1] * c[1]);
            Ixy -= m * c[0] * c[1];
            Ixz -= m * c[0] * c[2];
            Iyz -= m * c[1] * c[2];
            I[0] = m; I[1] = m * c[0]; I[2] = m * c[1]; I[3] = m * c[2];
            I[4] = Ixx; I[5] = Ixy; I[6] = Ixz; I[7] = Iyy; I[8] = Iyz; I[9] = Izz;
            const double* h = I + 1;
            double hxv[3], hxw[3];
            cross3(h, Aj + 3, hxv);
            cross3(h, Aj, hxw);
            const double na[3] = {Ixx * Aj[0] + Ixy * Aj[1] + Ixz * Aj[2] + hxv[0],
                                  Ixy * Aj[0] + Iyy * Aj[1] + Iyz * Aj[2] + hxv[1],
                                  Ixz * Aj[0] + Iyz * Aj[1] + Izz * Aj[2] + hxv[2]};
            const double fa[3] = {m * Aj[3] - hxw[0], m * Aj[4] - hxw[1], m * Aj[5] - hxw[2]};
            cross3(h, Vj + 3, hxv);
            cross3(h, Vj, hxw);
            const double nv[3] = {Ixx * Vj[0] + Ixy * Vj[1] + Ixz * Vj[2] + hxv[0],
                                  Ixy * Vj[0] + Iyy * Vj[1] + Iyz * Vj[2] + hxv[1],
                                  Ixz * Vj[0] + Iyz * Vj[1] + Izz * Vj[2] + hxv[2]};
            const double fv[3] = {m * Vj[3] - hxw[0], m * Vj[4] - hxw[1], m * Vj[5] - hxw[2]};
            double wxn[3], vxf[3], wxf[3];
            cross3(Vj, nv, wxn);
            cross3(Vj + 3, fv, vxf);
            cross3(Vj, fv, wxf);
            f[0] = na[0] + wxn[0] + vxf[0]; f[1] = na[1] + wxn[1] + vxf[1]; f[2] = na[2] + wxn[2] + vxf[2];
            f[3] = fa[0] + wxf[0]; f[4] = fa[1] + wxf[1]; f[5] = fa[2] + wxf[2];
        } else {
#pragma unroll
            for (int i = 0; i < 10; ++i) I[i] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) PB(L.IF + 16 * j, i) = I[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) PB(L.IF + 16 * j, 10 + i) = f[i];
        if (B->use_world[j]) {  // world-coordinate PD target: needs the joint's world rotation Rw * R_j
            double rq9[9], Rw[9], Wm[9];
            root_rot(rq9, Rw);
            mat3_mul(Rw, T, Wm);
            pb_pd_terms<TE>(X, cj_, j, N, n, e, ln, p, ws, L, Wm, env_ok);
        }
    };

    // ---------------- phase B: level by level -- compose transforms / velocities (level l) + body tasks (level l-1) ----------------
    for (int lv = 1; lv <= num_levels; ++lv) {
        const int a0 = (lv < num_levels) ? X->level_start[lv] : 0, a1 = (lv < num_levels) ? X->level_start[lv + 1] : 0;
        const int b0 = X->level_start[lv - 1], b1 = X->level_start[lv];
        const int ntask = (a1 - a0) + (b1 - b0);
        for (int t = warp; t < ntask; t += W) {
            if (t < a1 - a0) {
                const int j = X->level_joint[a0 + t];
                const int pj = X->parent[j];
                double Tp[12], Lj[12], T[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) { Tp[i] = PB(L.T + 12 * pj, i); Lj[i] = PB(L.T + 12 * j, i); }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        T[r * 3 + c] = Tp[r * 3] * Lj[c] + Tp[r * 3 + 1] * Lj[3 + c] + Tp[r * 3 + 2] * Lj[6 + c];
                    T[9 + r] = Tp[9 + r] + (Tp[r * 3] * Lj[9] + Tp[r * 3 + 1] * Lj[10] + Tp[r * 3 + 2] * Lj[11]);
                }
#pragma unroll
                for (int i = 0; i < 12; ++i) PB(L.T + 12 * j, i) = T[i];
                // s_j = S_j qd_j ; V_j = V_p + s_j ; a_j = a_p + V_p x s_j (motion cross product; V_j x s_j = V_p x s_j)
                double s[6] = {0, 0, 0, 0, 0, 0};
                const int c0 = X->first_col[j], nc = X->ncol[j];
                for (int c = 0; c < nc; ++c) {
                    const double w = PB(VA + 12 * j, c);
                    const int kind = X->col_kind[c0 + c], ax = X->col_axis[c0 + c];
                    double s6[6];
                    if (kind == CK_ANG) {
                        s6[0] = T[ax]; s6[1] = T[3 + ax]; s6[2] = T[6 + ax];
                        s6[3] = T[10] * s6[2] - T[11] * s6[1];
                        s6[4] = T[11] * s6[0] - T[9] * s6[2];
                        s6[5] = T[9] * s6[1] - T[10] * s6[0];
                    } else {  // CK_LIN
                        s6[0] = s6[1] = s6[2] = 0.0;
                        s6[3] = T[ax]; s6[4] = T[3 + ax]; s6[5] = T[6 + ax];
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) s[k] += s6[k] * w;
                }
                double Vp[6], Ap[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) { Vp[k] = PB(VA + 12 * pj, k); Ap[k] = PB(VA + 12 * pj, 6 + k); }
                double x0[3], x1[3], x2[3];
                cross3(Vp, s, x0);
                cross3(Vp + 3, s, x1);
                cross3(Vp, s + 3, x2);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    PB(VA + 12 * j, k) = Vp[k] + s[k];
                    PB(VA + 12 * j, 3 + k) = Vp[3 + k] + s[3 + k];
                    PB(VA + 12 * j, 6 + k) = Ap[k] + x0[k];
                    PB(VA + 12 * j, 9 + k) = Ap[3 + k] + (x1[k] + x2[k]);
                }
            } else {
                body_task(X->level_joint[b0 + (t - (a1 - a0))]);
            }
        }
        __syncthreads();
    }

    // ---------------- phase C: subtree sums, deepest level first; a task gathers the children of one joint ----------------
    for (int lv = num_levels - 2; lv >= 0; --lv) {
        const int a0 = X->level_start[lv], a1 = X->level_start[lv + 1];
        const int c0 = X->level_start[lv + 1], c1 = X->level_start[lv + 2];
        for (int t = warp; t < a1 - a0; t += W) {
            const int j = X->level_joint[a0 + t];
            for (int ci = c0; ci < c1; ++ci) {
                const int c = X->level_joint[ci];
                if (X->parent[c] != j) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) PB(L.IF + 16 * j, i) += PB(L.IF + 16 * c, i);
            }
        }
        __syncthreads();
    }

    // ---------------- phase D: per column -- S_c, F_c = I_c S_c, bias force, rhs, H row against the ancestor chain ----------------
    // (the H region held V / A until here: every read of them is behind the barriers above)
    for (int t = warp; t < nl; t += W) {
        const int c = nl - 1 - t;  // deepest (longest) rows first
        const int j = X->col_joint[c];
        double rq9[9];
        root_rot(rq9, nullptr);  // the root translation columns (rows of R(q_root)) sit on every ancestor chain
        double sv[6], Fc[6];
        pb_subspace<TE>(X->col_kind[c], X->col_axis[c], &PB(L.T + 12 * j, 0), rq9, sv);
        double I[10], f[6];
#pragma unroll
        for (int i = 0; i < 10; ++i) I[i] = PB(L.IF + 16 * j, i);
#pragma unroll
        for (int i = 0; i < 6; ++i) f[i] = PB(L.IF + 16 * j, 10 + i);
        const double m = I[0];
        const double* h = I + 1;
        double hxv[3], hxw[3];
        cross3(h, sv + 3, hxv);
        cross3(h, sv, hxw);
        Fc[0] = I[4] * sv[0] + I[5] * sv[1] + I[6] * sv[2] + hxv[0];
        Fc[1] = I[5] * sv[0] + I[7] * sv[1] + I[8] * sv[2] + hxv[1];
        Fc[2] = I[6] * sv[0] + I[8] * sv[1] + I[9] * sv[2] + hxv[2];
        Fc[3] = m * sv[3] - hxw[0];
        Fc[4] = m * sv[4] - hxw[1];
        Fc[5] = m * sv[5] - hxw[2];
        const double Cc = (sv[0] * f[0] + sv[1] * f[1] + sv[2] * f[2]) + (sv[3] * f[3] + sv[4] * f[4] + sv[5] * f[5]);
        const double hd = (Fc[0] * sv[0] + Fc[1] * sv[1] + Fc[2] * sv[2]) + (Fc[3] * sv[3] + Fc[4] * sv[4] + Fc[5] * sv[5]);
        const int di = X->col_dof[c];
        // the H region held V / A of the joints: their last readers (the body tasks of phase B) are behind barriers
        const int rb = B->rowbase[c];
        PB(L.H, rb) = hd + dt * PB(L.kd, j);
        PB(L.rhs, c) = PB(L.t0, di) - Cc;
        if (p.out_bias && env_ok) p.out_bias[e * n + di] = Cc;
        if (p.out_mass && env_ok) p.out_mass[(e * n + di) * n + di] = hd;
        int a = X->col_parent[c], k = 1;
        while (a != -1) {
            double sa[6];
            const int ja = X->col_joint[a];
            pb_subspace<TE>(X->col_kind[a], X->col_axis[a], &PB(L.T + 12 * ja, 0), rq9, sa);
            const double hv = (Fc[0] * sa[0] + Fc[1] * sa[1] + Fc[2] * sa[2]) + (Fc[3] * sa[3] + Fc[4] * sa[4] + Fc[5] * sa[5]);
            PB(L.H, rb + k) = hv;
            if (p.out_mass && env_ok) {
                const int da = X->col_dof[a];
                p.out_mass[(e * n + di) * n + da] = hv;
                p.out_mass[(e * n + da) * n + di] = hv;
            }
            a = X->col_parent[a];
            ++k;
        }
    }
    __syncthreads();

    // ---------------- phase E: branch-sparse L^T D L, column by column from the leaves; forward substitution rides along ----------------
    // Step k: for ancestors a_i (i = 1..d, a_1 = lambda(k)): H(a_i, a_j) -= H(k,a_i) H(k,a_j) / H(k,k) for j >= i, and
    // rhs(a_i) -= H(k,a_i) rhs(k) / H(k,k). Row k itself is kept unscaled; its diagonal slot receives 1 / H(k,k).
    // Warp w takes ancestor rows i = w+1 and d-w (d+1 updates together: the triangle is balanced across the warps).
    for (int k = nl - 1; k >= 0; --k) {
        const int d = B->cdepth[k];
        const int rb = B->rowbase[k];
        const double piv = PB(L.H, rb);
        const double inv = 1.0 / piv;
        if (d > 0) {
            const double bk = PB(L.rhs, k);
            for (int w0 = warp; 2 * w0 < d; w0 += W) {
#pragma unroll 1
                for (int side = 0; side < 2; ++side) {
                    const int i = side ? (d - w0) : (w0 + 1);
                    if (side && i == w0 + 1) break;  // middle row of an odd triangle: once
                    // a_i = i-th ancestor of k
                    int ai = X->col_parent[k];
                    for (int s = 1; s < i; ++s) ai = X->col_parent[ai];
                    const double li = PB(L.H, rb + i) * inv;
                    const int rbi = B->rowbase[ai];
                    for (int jj = i; jj <= d; ++jj) PB(L.H, rbi + (jj - i)) -= li * PB(L.H, rb + jj);
                    PB(L.rhs, ai) -= li * bk;
                }
            }
        }
        if (warp == 0) {
            if (!(piv > 0.0)) s_status[ln] |= 2;  // not positive definite: flagged instead of pivoting
        }
        __syncthreads();
        if (warp == (k % W)) PB(L.H, rb) = inv;  // after the barrier: every warp has read the pivot
    }
    __syncthreads();

    // ---------------- phase F: back substitution, column-depth by column-depth: x_k = (rhs_k - sum_i H(k,a_i) x_{a_i}) / H(k,k) ----------------
    for (int dep = 0; dep <= B->max_cdepth; ++dep) {
        const int s0 = B->cd_start[dep], s1 = B->cd_start[dep + 1];
        for (int t = warp; t < s1 - s0; t += W) {
            const int k = B->cd_order[s0 + t];
            const int rb = B->rowbase[k];
            double acc = PB(L.rhs, k);
            int a = X->col_parent[k];
            for (int i = 1; i <= dep; ++i) {
                acc -= PB(L.H, rb + i) * PB(L.rhs, a);
                a = X->col_parent[a];
            }
            PB(L.rhs, k) = acc * PB(L.H, rb);
        }
        __syncthreads();
    }

    // ---------------- phase G: tau = t0 - dt Kd_masked qdd, staged row-major ([env][n | 1]: conflict-free) and written coalesced ----------------
    const int np = n | 1;
    double* stage = ws + (size_t)L.T * TE;  // the transform region is dead
    for (int i = warp; i < n; i += W) {
        const int c = X->dof_col[i];
        double v = PB(L.t0, i);
        if (c >= 0) v -= (PB(L.kdm, X->dof_joint[i]) * dt) * PB(L.rhs, c);
        if (!isfinite(v) && lane < TE) atomicOr(&s_status[ln], 1);
        stage[ln * np + i] = v;
    }
    __syncthreads();
    {
        double* out = p.tau + (size_t)e0 * n;
        const int total = envs * n;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int r = idx / n, cidx = idx - r * n;
            const double v = stage[r * np + cidx];
            out[idx] = p.tau_accumulate ? out[idx] + v : v;
        }
        if (p.status && threadIdx.x < envs) p.status[e0 + threadIdx.x] = s_status[threadIdx.x];
    }
    if (p.out_bias && env_ok && warp == 0) {  // dead dofs are exact zeros
        for (int i = 0; i < n; ++i)
            if (X->dof_col[i] < 0) p.out_bias[e * n + i] = 0.0;
    }
}

#undef PB
