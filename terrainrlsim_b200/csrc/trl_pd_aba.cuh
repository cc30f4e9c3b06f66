// k_pd_aba: stable-PD torque WITHOUT materialising M -- thread per env, two depth-first walks. Included by
// trl_kernels.cu (inside its anonymous namespace).
//
// The reference builds M (CRBA) and C (RNEA), adds dt*Kd to the diagonal and solves (M + dt Kd) qdd = u - C with a
// dense LDLT (sim/ImpPDController.cpp:137-196). The same qdd follows from eliminating the system joint by joint from
// the leaves of the kinematic tree, which is what Featherstone's branch-sparse L^T D L factorisation does to M -- and
// carried out on 6x6 articulated quantities instead of on columns of M it is the articulated-body recursion with the
// diagonal dt*Kd acting as a joint-space ("rotor") inertia. In the root-joint frame used by all kernels here (every
// transform, velocity, inertia and force expressed in one frame, so parents accumulate children by plain addition):
//
//   down   : T_j, V_j, A_j (bias acceleration with qdd = 0, gravity included), body inertia I_j, f_j = I_j A_j + V_j x* I_j V_j
//            (identical to the RNEA pass of k_pd_block); u_j = Kp e_p + Kd e_v
//   up     : I^A_j = I_j + sum_c I^a_c ,  p^A_j = f_j + sum_c p^a_c                       (children c of j)
//            U = I^A S_j ,  D = S_j^T U + dt Kd_j(unmasked) ,  uh = u_j - S_j^T p^A
//            I^a_j = I^A - U D^-1 U^T ,  p^a_j = p^A + U D^-1 uh                           (handed to the parent)
//   root   : 6x6 solve  (S^T I^A S) qdd_root = -S^T p^A ,  delta_root = S qdd_root
//   down 2 : qdd_j = D^-1 (uh - U^T delta_parent) ,  delta_j = delta_parent + S_j qdd_j ,  tau_j = u_j - dt Kd_j(masked) qdd_j
//
// where delta_j is the part of the body's spatial acceleration caused by the joint accelerations. This is exactly
// (M + dt Kd) qdd = u - C: S^T(I^A delta + p^A) is row j of M qdd + C restricted to the subtree. The result agrees with the
// LDLT path (k_pd_block, and the reference through the oracle) to ~1e-13; 1e-9 is the bar.
//
// Cost: O(N) 6x6 work -- about 3.3 kFLOP per biped3D env instead of 8.2 kFLOP (16.3 k by SURVEY.md's count) for
// CRBA + RNEA + LDLT, and no n x n matrix.
//
// Mapping: lane = env (32-env tile per block), ONE WARP PER BRANCH (a branch = the subtree of one child of the root):
// the branches are independent between the root's "down" step and its solve, so a biped's two legs, a raptor's four
// limbs, ... are walked concurrently by different warps -- this is what lifts the chain off the single-warp latency
// floor (16384 envs are only 512 tiles on 592 SM sub-partitions). Warp 0 also does the root. Three block barriers.
//
// State: a walker keeps one slot per tree LEVEL of its branch in shared memory -- T12, f6, u3, Kd1 (22 doubles), or T12
// alone in the SLIM layout, where f / u / Kd of a joint go through the global scratch between its enter and its leave
// (chosen by the launcher when it saves a wave of tiles or when the full tiles squeeze L1: see launch_pd_aba). The
// articulated inertia / bias force travel up a chain in REGISTERS (a joint's "leave" is followed by its parent's when it
// is an only child) and velocities / accelerations travel down in registers (a joint's "enter" is followed by its first
// child's); only joints with two or more children own a [V6 A6 IA21 P6] block. The depth-first event list is walked as
// alternating runs of enters and leaves so that each loop carries only its own hand-off. biped3D, slim: 27 + 2 x 27 +
// 2 x 36 = 153 doubles per env = 39 KB per tile (full: 213 = 54 KB); all 512 tiles of 16384 envs are resident at once
// either way. What the second walk needs (U D^-1, D^-1 uh, S, u, Kd: 14k+1 doubles per joint with k dofs, plus the
// finished tau of dead slots) goes through a tile-major global scratch (1.6 KB per env, written and read back by the same
// warp: L2-resident). The model's index tables and per-joint constants arrive as a __grid_constant__ kernel parameter
// (constant bank; every lookup is warp-uniform), the root is solved as I^A delta = -p^A.
//
// Restrictions (callers fall back to k_pd_block): every body valid (the reference's handling of shape-less bodies
// leaves M structurally inconsistent: SURVEY.md, k_rbd_extras note), no M / C output requested.
#pragma once

// symmetric 6x6 in 21 doubles, upper triangle row by row: (r, c), r <= c
__host__ __device__ constexpr int sym6(int r, int c) { return r <= c ? r * 6 - r * (r - 1) / 2 + (c - r) : c * 6 - c * (c - 1) / 2 + (r - c); }

// y = A x for symmetric A (21)
TRL_DEV void sym6_mulv(const double* A, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += A[sym6(r, c)] * x[c];
        y[r] = s;
    }
}

TRL_DEV double dot6(const double* a, const double* b) {
    return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) + (a[3] * b[3] + a[4] * b[4] + a[5] * b[5]);
}


// body inertia of joint j in the root frame from its transform T: I = (m, h = m c, Ixx Ixy Ixz Iyy Iyz Izz about the origin)
TRL_DEV void aba_inertia(const double* T, const double* cj_, double* I) {
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
}

// body inertia and f = I a + V x* (I V)
TRL_DEV void aba_body(const double* T, const double* cj_, const double* Vj, const double* Aj, double* I, double* f) {
    aba_inertia(T, cj_, I);
    const double m = I[0];
    const double* h = I + 1;
    const double Ixx = I[4], Ixy = I[5], Ixz = I[6], Iyy = I[7], Iyz = I[8], Izz = I[9];
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
}

// joint-subspace column in the root frame from the joint's transform T (registers): angular about / linear along axis `ax`
TRL_DEV void aba_column(int kind, int ax, const double* T, double* S) {
    const double a0 = (ax == 0) ? T[0] : (ax == 1) ? T[1] : T[2];
    const double a1 = (ax == 0) ? T[3] : (ax == 1) ? T[4] : T[5];
    const double a2 = (ax == 0) ? T[6] : (ax == 1) ? T[7] : T[8];
    if (kind == CK_ANG) {
        S[0] = a0; S[1] = a1; S[2] = a2;
        S[3] = T[10] * a2 - T[11] * a1; S[4] = T[11] * a0 - T[9] * a2; S[5] = T[9] * a1 - T[10] * a0;
    } else {
        S[0] = S[1] = S[2] = 0.0;
        S[3] = a0; S[4] = a1; S[5] = a2;
    }
}

// workspace layout, in elements of 32 doubles ([element][lane]):
//   [0, 27)                       root V, A, f, world <- root rotation (second walk: delta of the root in [0, 6))
//   [27, 27 + 27 nb)              per branch: I^a (21), p^a (6) handed to the root; afterwards the tau staging rows
//   [.., + W x branch_area)       one area per WARP: 22 per level below the root (T12 f6 u3 Kd1; second walk: delta in
//                                 [0, 6)), then the [V6 A6 IA21 P6] blocks of joints with several children
__host__ __device__ inline int aba_ws_elems(int nb, int branch_area, int warps) { return 27 + 27 * nb + warps * branch_area; }

#ifdef TRL_PD_CLK
__device__ long long g_pd_clk[8][32];
#define PD_CLK(i) do { if (blockIdx.x == 100 && (threadIdx.x & 31) == 0) g_pd_clk[threadIdx.x >> 5][(i)] = clock64(); } while (0)
#else
#define PD_CLK(i) do { } while (0)
#endif

// The model's index tables and per-joint constants travel as a KERNEL PARAMETER (7.3 KB, __grid_constant__: the constant
// bank, read with warp-uniform indices -- lane = env, so every table lookup of the walk is uniform): no staging pass at the
// start of every tile (512 blocks used to pull the same 3 KB through L2: 2.7 k cycles before the first useful instruction),
// no barrier, and 3 KB of shared memory back per tile.
// Two sizes: up to 12 joints everything fits the classic 4 KB parameter space (3.9 KB); larger characters use the large
// parameter space (7.3 KB for 32 joints), whose launches cost about 2 us more outside CUDA graphs.
template <int NJ>
struct AbaTables {
    ModelIdx ix;
    PdBlkIdx bx;
    double cst[NJ * kCst];
};

// SLIM: a level slot keeps only T (12 instead of 22 doubles); the body force, the PD forces and Kd of a joint go through the
// global scratch between its enter and its leave. For characters whose tile would otherwise allow only one tile per SM
// while the batch needs more (raptor x 8192: 147 KB per tile, 256 tiles on 148 SMs = two waves; slim 96 KB = one wave).
template <int NJ, bool SLIM>
__global__ void __launch_bounds__(256)
k_pd_aba(const __grid_constant__ AbaTables<NJ> tabs, double gx, double gy, double gz, int E, StepPtrs p, double* __restrict__ scratch,
         int scratch_stride, int fuk_off, int area_elems, int N, int n) {
    constexpr int kSlot = SLIM ? 12 : 22;
    KTrace trace_(TR_PD_TREE);
    extern __shared__ double smem[];
    __shared__ int s_status[32];
    PD_CLK(0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int tile = blockIdx.x;
    const int e0 = tile * kTile;
    const int envs = min(kTile, E - e0);
    const bool env_ok = lane < envs;
    const size_t e = (size_t)(e0 + (env_ok ? lane : envs - 1));  // tail lanes redo the last env and store nothing
    const double* q = p.pose + e * n;
    const double* qd = p.vel + e * n;
    const double* tq = p.tar_pose + e * n;
    const double* tqd = p.tar_vel + e * n;
    {
        // The tile's rows of the four input arrays are contiguous ([32][n] doubles each) but every lane walks its own row,
        // joint by joint: pull them into L2 now, in whole lines, so that the walk's strided loads do not each wait on HBM.
        const size_t row0 = (size_t)e0 * n * sizeof(double);
        const int lines = (envs * n * (int)sizeof(double) + 127) >> 7;
        for (int l = threadIdx.x; l < lines; l += blockDim.x) {
            const size_t off = row0 + ((size_t)l << 7);
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)p.pose + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)p.vel + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)p.tar_pose + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)p.tar_vel + off));
        }
    }
    // the root's coordinates (warp 0 needs them first): in flight while the model tables are staged
    double q_root[4], qd_root[7], grav[3];
    if (warp == 0) {
        grav[0] = gx; grav[1] = gy; grav[2] = gz;
#pragma unroll
        for (int i = 0; i < 4; ++i) q_root[i] = q[3 + i];
#pragma unroll
        for (int i = 0; i < 7; ++i) qd_root[i] = qd[i];
    }
    if (threadIdx.x < 32) s_status[threadIdx.x] = 0;  // (ordered before its first use by the barriers of the walk)
    PD_CLK(1);
    const ModelIdx* X = &tabs.ix;
    const PdBlkIdx* B = &tabs.bx;
    const int nb = B->num_branches;
    const double* cst = tabs.cst;
    double* ws0 = smem;  // workspace base (no lane offset)
    double* ws = ws0 + lane;                 // element k of this env: ws[k * 32]
    double* sc = scratch + (size_t)tile * scratch_stride * kTile + lane;
    const double dt = p.dt;
    const int kOut = 27;
    double* area = ws + (size_t)(kOut + 27 * nb + warp * area_elems) * kTile;  // this warp's branch area
    int status = 0;

    // ======================= root, down: V_0 = S qd, A_0 = gravity + c_root (warp 0) =======================
    if (warp == 0) {
        double rq9[9], Rw[9];  // R(q_root) and the world <- root rotation
        quat_to_mat(q_root[0], q_root[1], q_root[2], q_root[3], rq9);
        mat3_mul(cst, rq9, Rw);
        double wr[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) wr[i] = qd_root[i];
        const double ng[3] = {-grav[0], -grav[1], -grav[2]};
        double g0[3];
        mat3T_mulv(Rw, ng, g0);  // a_0 = X_{world->root}(0, -g)  (RBDUtil.cpp:17-18,59-60)
        double dq[4];
        const double qr4[4] = {q_root[0], q_root[1], q_root[2], q_root[3]};
        quat_diff_mul(qr4, wr + 3, dq);  // c_root (cRBDUtil::BuildCjRoot, RBDUtil.cpp:867-904)
        const double qw = qr4[0], qx = qr4[1], qy = qr4[2], qz = qr4[3];
        const double dw = dq[0], dx = dq[1], dy = dq[2], dz = dq[3];
        const double m00 = 4 * (qw * dw + qx * dx), m11 = 4 * (qw * dw + qy * dy), m22 = 4 * (qw * dw + qz * dz);
        const double m10 = 2 * (dx * qy + qx * dy - dw * qz - qw * dz);
        const double m01 = 2 * (dx * qy + qx * dy + dw * qz + qw * dz);
        const double m20 = 2 * (dx * qz + qx * dz + dw * qy + qw * dy);
        const double m02 = 2 * (dx * qz + qx * dz - dw * qy - qw * dy);
        const double m21 = 2 * (dy * qz + qy * dz - dw * qx - qw * dx);
        const double m12 = 2 * (dy * qz + qy * dz + dw * qx + qw * dx);
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
        for (int k = 0; k < 6; ++k) { ws[k * kTile] = V[k]; ws[(6 + k) * kTile] = A[k]; }
        const double Tid[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
        double Ir[10], fr[6];  // f of the root body: parked in shared memory until the root solve
        aba_body(Tid, cst, V, A, Ir, fr);
#pragma unroll
        for (int k = 0; k < 6; ++k) ws[(12 + k) * kTile] = fr[k];
#pragma unroll
        for (int k = 0; k < 9; ++k) ws[(18 + k) * kTile] = Rw[k];  // for the joints with world-coordinate targets
    }
    PD_CLK(2);
    __syncthreads();
    PD_CLK(3);
#ifdef TRL_PD_CLK
    int clk_i = 4;
#endif

    // ======================= branches: first walk (down: T V A f u ; up: articulated elimination) =======================
    const int nev_all = 2 * N;
    (void)nev_all;
    for (int b = warp; b < nb; b += W) {
        // The depth-first event list is walked as alternating RUNS: a run of "enter" events goes down a chain (V / A of
        // the joint entered last reach its first child in registers), a run of "leave" events comes back up (I^a / p^a
        // of the joint left last reach its parent in registers when it is an only child). Nothing is carried from one
        // run into the next -- an enter after a leave is a later child, which reads its parent's block; a leave after an
        // enter is a leaf -- so each loop keeps only its own hand-off live.
        const int ev0 = B->b_ev0[b], ev1 = B->b_ev1[b];
        int evi = ev0;
        while (evi < ev1) {
            double Vc[6], Ac[6];      // V, A of the joint entered last
#pragma unroll
            for (int i = 0; i < 6; ++i) { Vc[i] = 0.0; Ac[i] = 0.0; }
            int last_entered = -1;
            for (; evi < ev1; ++evi) {
                const int code = X->dfs_event[evi];
                if (code < 0) break;
                // ---------------- enter joint j ----------------
                const int j = code;
                const int lv = X->level[j];
                const int pj = X->parent[j];
                double* Wl = area + (size_t)(kSlot * (lv - 1)) * kTile;  // T12 (f6 u3 kd1)
                double* Fx = SLIM ? sc + (size_t)(fuk_off + 10 * j) * kTile : Wl + 12 * kTile;  // f6 u3 kd1 of the joint
                const int o = X->offset[j], sz = X->size[j], type = X->type[j];
                const double* cj_ = cst + kCst * j;
                double qj[4], wj[4], tar[4], tarv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool in = i < sz;
                    qj[i] = in ? q[o + i] : 0.0;
                    wj[i] = in ? qd[o + i] : 0.0;
                    tar[i] = in ? tq[o + i] : 0.0;
                    tarv[i] = in ? tqd[o + i] : 0.0;
                }
                // the PD terms' own inputs are requested here as well, with the joint's coordinates: asked for where they
                // are used (after the kinematics) their trip to L2 / HBM would sit on the chain
                const bool valid = X->pd_valid[j] != 0;
                const unsigned char ja = (valid && p.joint_active) ? p.joint_active[e * N + j] : (unsigned char)1;
                const double kp_in = (valid && p.kp) ? p.kp[e * N + j] : cj_[19];
                const double kd_in = (valid && p.kd) ? p.kd[e * N + j] : cj_[20];
                double R[9], t[3], T[12];
                joint_local_trans(X, j, qj, cj_, cj_ + 9, R, t);
                double Vp[6], Ap[6];
                if (lv == 1) {  // parent = root: identity transform, V / A from the root block
#pragma unroll
                    for (int i = 0; i < 9; ++i) T[i] = R[i];
                    T[9] = t[0]; T[10] = t[1]; T[11] = t[2];
#pragma unroll
                    for (int k = 0; k < 6; ++k) { Vp[k] = ws[k * kTile]; Ap[k] = ws[(6 + k) * kTile]; }
                } else {
                    const double* Wp = area + (size_t)(kSlot * (lv - 2)) * kTile;
                    double Tp[12];
#pragma unroll
                    for (int i = 0; i < 12; ++i) Tp[i] = Wp[i * kTile];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            T[r * 3 + c] = Tp[r * 3] * R[c] + Tp[r * 3 + 1] * R[3 + c] + Tp[r * 3 + 2] * R[6 + c];
                        T[9 + r] = Tp[9 + r] + (Tp[r * 3] * t[0] + Tp[r * 3 + 1] * t[1] + Tp[r * 3 + 2] * t[2]);
                    }
                    if (last_entered == pj) {  // first child: the parent's V / A are still in registers
#pragma unroll
                        for (int k = 0; k < 6; ++k) { Vp[k] = Vc[k]; Ap[k] = Ac[k]; }
                    } else {                   // later child: the parent owns a block
                        const double* Xp = area + (size_t)(B->xoff[pj] - (SLIM ? 10 * B->b_depth[b] : 0)) * kTile;
#pragma unroll
                        for (int k = 0; k < 6; ++k) { Vp[k] = Xp[k * kTile]; Ap[k] = Xp[(6 + k) * kTile]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < 12; ++i) Wl[i * kTile] = T[i];
                // s_j = S_j qd_j ; V_j = V_p + s_j ; A_j = A_p + V_p x s_j
                double s[6] = {0, 0, 0, 0, 0, 0};
                const int c0 = X->first_col[j], nc = X->ncol[j];
                for (int c = c0; c < c0 + nc; ++c) {
                    const int di = X->col_dof[c] - o;
                    const double w = (di == 0) ? wj[0] : (di == 1) ? wj[1] : (di == 2) ? wj[2] : wj[3];
                    const int kind = X->col_kind[c], ax = X->col_axis[c];
                    const double a0 = (ax == 0) ? T[0] : (ax == 1) ? T[1] : T[2];
                    const double a1 = (ax == 0) ? T[3] : (ax == 1) ? T[4] : T[5];
                    const double a2 = (ax == 0) ? T[6] : (ax == 1) ? T[7] : T[8];
                    if (kind == CK_ANG) {
                        s[0] += a0 * w; s[1] += a1 * w; s[2] += a2 * w;
                        s[3] += (T[10] * a2 - T[11] * a1) * w;
                        s[4] += (T[11] * a0 - T[9] * a2) * w;
                        s[5] += (T[9] * a1 - T[10] * a0) * w;
                    } else {
                        s[3] += a0 * w; s[4] += a1 * w; s[5] += a2 * w;
                    }
                }
                double x0[3], x1[3], x2[3];
                cross3(Vp, s, x0);
                cross3(Vp + 3, s, x1);
                cross3(Vp, s + 3, x2);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    Vc[k] = Vp[k] + s[k];
                    Vc[3 + k] = Vp[3 + k] + s[3 + k];
                    Ac[k] = Ap[k] + x0[k];
                    Ac[3 + k] = Ap[3 + k] + (x1[k] + x2[k]);
                }
                last_entered = j;
                if (B->xoff[j] >= 0) {  // several children: V / A for the later ones, zeroed accumulators
                    double* Xj = area + (size_t)(B->xoff[j] - (SLIM ? 10 * B->b_depth[b] : 0)) * kTile;
#pragma unroll
                    for (int k = 0; k < 6; ++k) { Xj[k * kTile] = Vc[k]; Xj[(6 + k) * kTile] = Ac[k]; }
#pragma unroll
                    for (int k = 12; k < 39; ++k) Xj[k * kTile] = 0.0;
                }
                // body force f_j = I_j A_j + V_j x* I_j V_j (the inertia itself is rebuilt from T_j on the way up)
                {
                    double Ib[10], f[6];
                    aba_body(T, cj_, Vc, Ac, Ib, f);
#pragma unroll
                    for (int k = 0; k < 6; ++k) Fx[k * kTile] = f[k];
                }
                // PD terms (cImpPDController::CalcControlForces, ImpPDController.cpp:137-196; targets through
                // cPDController::SetTargetTheta / GetTargetTheta, PDController.cpp:191-273)
                const bool active = valid && ja != 0;
                const double kp = valid ? kp_in : 0.0, kd = valid ? kd_in : 0.0;
                const double kpm = active ? kp : 0.0, kdm = active ? kd : 0.0;
                Fx[9 * kTile] = kd;  // unmasked gains go on the diagonal
                const bool world = B->use_world[j] != 0;
                double Wm[9];
                if (world) {  // joint world rotation = (attach_root R(q_root)) R_j; the first factor is warp 0's, from the root step
                    double Rw[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) Rw[k] = ws[(18 + k) * kTile];
                    mat3_mul(Rw, T, Wm);
                }
                double ep[4] = {0, 0, 0, 0};
                const bool sph = valid && type == JT_SPHERICAL;
                double tar0 = 0.0;
                if (valid && !sph && sz > 0) tar0 = tar[0];
                if (sph) {
                    double dq[4], qi[4];
                    quat_diff_mul(qj, wj, dq);
#pragma unroll
                    for (int i = 0; i < 4; ++i) qi[i] = qj[i] + dt * dq[i];
                    quat_normalize(qi);
                    if (tar[0] * tar[0] + tar[1] * tar[1] + tar[2] * tar[2] + tar[3] * tar[3] == 0.0) tar[0] = 1.0;
                    else quat_normalize(tar);
                    if (world) {  // tar = rel_rot * (world_rot^-1 * tar), PDController.cpp:254-267
                        // rel_rot = AxisAngleToQuaternion(QuaternionToAxisAngle(q_j)) (util/MathUtil.cpp:399-416): with theta =
                        // 2 sg asin(st) the round trip is [cos(asin st), sin(sg asin st) q_xyz / st] = [sqrt((1 - st)(1 + st)),
                        // sg q_xyz] -- no asin, sincos or divisions; the identity below the 1e-4 threshold and for q_w == 0
                        const double st = sqrt(qj[1] * qj[1] + qj[2] * qj[2] + qj[3] * qj[3]);
                        const double sg = (double)((0.0 < qj[0]) - (qj[0] < 0.0));
                        const bool turn = (st > 0.0001) && (sg != 0.0);
                        const double ch = turn ? sqrt((1.0 - st) * (1.0 + st)) : 1.0;
                        const double rel[4] = {ch, turn ? sg * qj[1] : 0.0, turn ? sg * qj[2] : 0.0, turn ? sg * qj[3] : 0.0};
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
                } else if (world && valid && type == JT_REVOLUTE) {  // PDController.cpp:238-252
                    double axis_world[3], theta_world;
                    pb_rot_to_axis_angle(Wm, axis_world, theta_world);
                    const double dwv = axis_world[0] * Wm[2] + axis_world[1] * Wm[5] + axis_world[2] * Wm[8];
                    tar0 -= dwv * theta_world - Wm[8] * qj[0];
                }
                // per dof: u = Kp e_p + Kd e_v. Live dofs keep u (<= 3 per joint) for the way up; dead slots
                // (spherical slot 3) are decoupled 1x1 systems and are finished here.
                int lc = 0, nd = 0;
                if (sph && X->dof_col[o + 3] < 0 && X->dof_col[o] >= 0 && X->dof_col[o + 1] >= 0 && X->dof_col[o + 2] >= 0) {
                    // the usual spherical joint (three live dofs, dead slot 3) without the per-dof bookkeeping
#pragma unroll
                    for (int i = 0; i < 3; ++i) Fx[(6 + i) * kTile] = kpm * ep[i] + kdm * (tarv[i] - wj[i]);
                    const double ev3 = 0.0 - wj[3];
                    const double u3 = kpm * ep[3] + kdm * ev3;
                    const double dg = dt * kd;
                    const double acc = (fabs(dg) > DBL_MIN) ? u3 / dg : 0.0;
                    const double t0 = kpm * ep[3] + kdm * (ev3 - dt * acc);
                    if (!isfinite(t0)) status |= 1;
                    sc[(size_t)(B->aba_off[j] + 1 + 14 * nc) * kTile] = t0;
                } else
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i >= sz) break;
                    const double qdi = wj[i];
                    double epi = 0.0, evi_ = 0.0;
                    if (valid) {
                        if (sph) {
                            epi = ep[i];
                            evi_ = ((i < 3) ? tarv[i] : 0.0) - qdi;
                        } else {
                            const double ti = (i == 0) ? tar0 : tar[i];
                            epi = (ti - (qj[i] + dt * qdi)) / 1.0;
                            evi_ = tarv[i] - qdi;
                        }
                    }
                    const double u = kpm * epi + kdm * evi_;
                    if (X->dof_col[o + i] < 0) {
                        const double dg = dt * kd;
                        const double acc = (fabs(dg) > DBL_MIN) ? u / dg : 0.0;
                        const double t0 = kpm * epi + kdm * (evi_ - dt * acc);
                        if (!isfinite(t0)) status |= 1;
                        sc[(size_t)(B->aba_off[j] + 1 + 14 * nc + nd) * kTile] = t0;  // after the joint's live data
                        ++nd;
                    } else {
                        if (lc < 3) Fx[(6 + lc) * kTile] = u;
                        ++lc;
                    }
                }
                sc[(size_t)B->aba_off[j] * kTile] = kdm * dt;  // masked Kd dt, first scratch slot of the joint
#ifdef TRL_PD_CLK
                PD_CLK(clk_i); ++clk_i;
#endif
            }
            double IAc[21], Pc[6];    // articulated inertia / bias force of the joint left last
#pragma unroll
            for (int i = 0; i < 21; ++i) IAc[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) Pc[i] = 0.0;
            for (; evi < ev1; ++evi) {
                const int code = X->dfs_event[evi];
                if (code >= 0) break;
                // ---------------- leave joint j: its subtree is complete ----------------
                const int j = -1 - code;
                const int lv = X->level[j];
                const int pj = X->parent[j];
                double* Wl = area + (size_t)(kSlot * (lv - 1)) * kTile;
                const double* Fx = SLIM ? sc + (size_t)(fuk_off + 10 * j) * kTile : Wl + 12 * kTile;
                const double* cj_ = cst + kCst * j;
                double P[6], uk[4];  // requested first (SLIM: a trip to L2 that overlaps the inertia rebuild)
#pragma unroll
                for (int i = 0; i < 6; ++i) P[i] = Fx[i * kTile];
#pragma unroll
                for (int i = 0; i < 4; ++i) uk[i] = Fx[(6 + i) * kTile];
                double T[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) T[i] = Wl[i * kTile];
                // I^A = I_j + children, p^A = f_j + children
                double IA[21];
                {
                    double Ib[10];
                    aba_inertia(T, cj_, Ib);
                    const double m = Ib[0], h0 = Ib[1], h1 = Ib[2], h2 = Ib[3];
                    IA[sym6(0, 0)] = Ib[4]; IA[sym6(0, 1)] = Ib[5]; IA[sym6(0, 2)] = Ib[6];
                    IA[sym6(1, 1)] = Ib[7]; IA[sym6(1, 2)] = Ib[8]; IA[sym6(2, 2)] = Ib[9];
                    IA[sym6(0, 3)] = 0.0; IA[sym6(0, 4)] = -h2; IA[sym6(0, 5)] = h1;
                    IA[sym6(1, 3)] = h2;  IA[sym6(1, 4)] = 0.0; IA[sym6(1, 5)] = -h0;
                    IA[sym6(2, 3)] = -h1; IA[sym6(2, 4)] = h0;  IA[sym6(2, 5)] = 0.0;
                    IA[sym6(3, 3)] = m; IA[sym6(3, 4)] = 0.0; IA[sym6(3, 5)] = 0.0;
                    IA[sym6(4, 4)] = m; IA[sym6(4, 5)] = 0.0;
                    IA[sym6(5, 5)] = m;
                }
                const int nch = B->nchild[j];
                if (nch == 1) {  // only child: its result is still in registers
#pragma unroll
                    for (int i = 0; i < 21; ++i) IA[i] += IAc[i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) P[i] += Pc[i];
                } else if (nch > 1) {
                    const double* Xj = area + (size_t)(B->xoff[j] - (SLIM ? 10 * B->b_depth[b] : 0)) * kTile;
#pragma unroll
                    for (int i = 0; i < 21; ++i) IA[i] += Xj[(12 + i) * kTile];
#pragma unroll
                    for (int i = 0; i < 6; ++i) P[i] += Xj[(33 + i) * kTile];
                }
                const int c0 = X->first_col[j], nc = X->ncol[j];
                double* sj = sc + (size_t)B->aba_off[j] * kTile;  // [kdm dt][u k][uh' k][S 6k][U' 6k][dead tau]
                const double kdd = dt * uk[3];
                // The joint's columns are eliminated one after the other (a k-dof joint = k stacked 1-dof joints: the same
                // arithmetic as an LDL^T of its k x k block, with a third of the registers). dt*Kd sits on every diagonal.
                for (int a = 0; a < nc; ++a) {
                    double S[6], U[6];
                    aba_column(X->col_kind[c0 + a], X->col_axis[c0 + a], T, S);
                    sym6_mulv(IA, S, U);
                    const double D = dot6(S, U) + kdd;
                    if (!(D > 0.0)) status |= 2;
                    const double u = (a == 0) ? uk[0] : (a == 1) ? uk[1] : uk[2];
                    const double inv = fast_rcp(D);
                    const double uhs = (u - dot6(S, P)) * inv;
                    double Us[6];
#pragma unroll
                    for (int r = 0; r < 6; ++r) Us[r] = U[r] * inv;
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
#pragma unroll
                        for (int c = r; c < 6; ++c) IA[sym6(r, c)] -= Us[r] * U[c];
                        P[r] += U[r] * uhs;
                    }
                    sj[(1 + a) * kTile] = u;
                    sj[(1 + nc + a) * kTile] = uhs;
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
                        sj[(1 + 2 * nc + 6 * a + r) * kTile] = S[r];
                        sj[(1 + 8 * nc + 6 * a + r) * kTile] = Us[r];
                    }
                }
                // hand I^a, p^a to the parent: registers (only child), its block (several children) or the root's slot
                if (lv == 1) {
                    double* O = ws + (size_t)(kOut + 27 * b) * kTile;
#pragma unroll
                    for (int i = 0; i < 21; ++i) O[i * kTile] = IA[i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) O[(21 + i) * kTile] = P[i];
                } else if (B->nchild[pj] == 1) {
#pragma unroll
                    for (int i = 0; i < 21; ++i) IAc[i] = IA[i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) Pc[i] = P[i];
                } else {
                    double* Xp = area + (size_t)(B->xoff[pj] - (SLIM ? 10 * B->b_depth[b] : 0)) * kTile;
#pragma unroll
                    for (int i = 0; i < 21; ++i) Xp[(12 + i) * kTile] += IA[i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) Xp[(33 + i) * kTile] += P[i];
                }
#ifdef TRL_PD_CLK
                PD_CLK(clk_i); ++clk_i;
#endif
            }
        }
    }
    PD_CLK(16);
    __syncthreads();
    PD_CLK(17);

    // ======================= root: I^A delta = -p^A (warp 0) =======================
    const int np = n | 1;
    const bool staged = np <= 27 * nb;
    if (warp == 0) {
        double IA[21], P[6];
        {
            const double Tid[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
            double Ir[10];
            aba_inertia(Tid, cst, Ir);  // constant: the root body in its own frame
            const double m = Ir[0], h0 = Ir[1], h1 = Ir[2], h2 = Ir[3];
            IA[sym6(0, 0)] = Ir[4]; IA[sym6(0, 1)] = Ir[5]; IA[sym6(0, 2)] = Ir[6];
            IA[sym6(1, 1)] = Ir[7]; IA[sym6(1, 2)] = Ir[8]; IA[sym6(2, 2)] = Ir[9];
            IA[sym6(0, 3)] = 0.0; IA[sym6(0, 4)] = -h2; IA[sym6(0, 5)] = h1;
            IA[sym6(1, 3)] = h2;  IA[sym6(1, 4)] = 0.0; IA[sym6(1, 5)] = -h0;
            IA[sym6(2, 3)] = -h1; IA[sym6(2, 4)] = h0;  IA[sym6(2, 5)] = 0.0;
            IA[sym6(3, 3)] = m; IA[sym6(3, 4)] = 0.0; IA[sym6(3, 5)] = 0.0;
            IA[sym6(4, 4)] = m; IA[sym6(4, 5)] = 0.0;
            IA[sym6(5, 5)] = m;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) P[i] = ws[(12 + i) * kTile];
        for (int b = 0; b < nb; ++b) {
            const double* O = ws + (size_t)(kOut + 27 * b) * kTile;
#pragma unroll
            for (int i = 0; i < 21; ++i) IA[i] += O[i * kTile];
#pragma unroll
            for (int i = 0; i < 6; ++i) P[i] += O[(21 + i) * kTile];
        }
        // The root's six columns S = [0 I; Q^T 0] (three translations = the rows of Q = R(q_root), three unit rotations) span
        // all of R^6 and S is orthogonal, so (S^T I^A S) qdd = -S^T p^A is I^A delta = -p^A for delta = S qdd -- the root's
        // spatial acceleration, which is all the second walk needs (the root has no controller: its tau is 0). The 6 x 6
        // system is solved on I^A itself: no products with Q, no R(q_root).
        double A[6][6], bb[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int c = 0; c <= r; ++c) A[r][c] = IA[sym6(r, c)];  // lower triangle
            bb[r] = -P[r];
        }
        // unpivoted LDL^T (I^A of the whole character is symmetric positive definite), forward / back substitution
        double Lm[6][6], dd[6], dinv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double d = A[k][k];
#pragma unroll
            for (int s = 0; s < k; ++s) d -= Lm[k][s] * Lm[k][s] * dd[s];
            dd[k] = d;
            if (!(d > 0.0)) status |= 2;
            dinv[k] = fast_rcp(d);
#pragma unroll
            for (int i = k + 1; i < 6; ++i) {
                double v = A[i][k];
#pragma unroll
                for (int s = 0; s < k; ++s) v -= Lm[i][s] * Lm[k][s] * dd[s];
                Lm[i][k] = v * dinv[k];
            }
        }
        double y[6], x[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = bb[k];
#pragma unroll
            for (int s = 0; s < k; ++s) v -= Lm[k][s] * y[s];
            y[k] = v;
        }
#pragma unroll
        for (int k = 5; k >= 0; --k) {
            double v = y[k] * dinv[k];
#pragma unroll
            for (int i = k + 1; i < 6; ++i) v -= Lm[i][k] * x[i];
            x[k] = v;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) ws[r * kTile] = x[r];  // delta of the root (V is dead)
    }
    PD_CLK(18);
    __syncthreads();
    PD_CLK(19);
#ifdef TRL_PD_CLK
    clk_i = 20;
#endif

    // ======================= second walk: qdd_j = D^-1 (uh - U^T delta_p), delta_j = delta_p + S_j qdd_j, tau =======================
    // tau is staged as [32][n | 1] rows (conflict-free for the per-env writes here and for the coalesced write below) in the
    // region that carried the branch results; delta of a level reuses the first 6 elements of its slot (T is dead).
    double* stage = ws0 + (size_t)kOut * kTile + (size_t)lane * np;
    auto put_tau = [&](int dof, double v) {
        if (staged) stage[dof] = v;
        else if (env_ok) {
            double* dst = p.tau + e * n + dof;
            *dst = p.tau_accumulate ? *dst + v : v;
        }
    };
    if (warp == 0)
        for (int i = 0; i < 7; ++i) put_tau(i, 0.0);  // root dofs: no controller
    for (int b = warp; b < nb; b += W) {
        const int p0 = B->b_pre0[b], p1 = B->b_pre1[b];
        for (int pi = p0; pi < p1; ++pi) {
            const int j = B->preorder[pi];
            const int lv = X->level[j];
            double* Wl = area + (size_t)(kSlot * (lv - 1)) * kTile;
            const double* Dp = (lv == 1) ? ws : area + (size_t)(kSlot * (lv - 2)) * kTile;
            if (pi + 1 < p1) {  // the next joint's scratch rows (contiguous, 256 B each): one cooperative L1 prefetch
                const int jn = B->preorder[pi + 1];
                const char* base = (const char*)(scratch + ((size_t)tile * scratch_stride + B->aba_off[jn]) * kTile);
                const int lines = (1 + 14 * X->ncol[jn]) * 2;
                for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)l * 128));
            }
            double dl[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) dl[r] = Dp[r * kTile];
            const int c0 = X->first_col[j], nc = X->ncol[j];
            const double* sj = sc + (size_t)B->aba_off[j] * kTile;
            const double kdmdt = sj[0];
            // every column's hand-off (u, D^-1 uh, S, U D^-1) is requested before the first one is used: one trip to L2 per
            // joint instead of one per column
            double cu[3], cq[3], cS[3][6], cU[3][6];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (a < nc) {
                    cu[a] = sj[(1 + a) * kTile];
                    cq[a] = sj[(1 + nc + a) * kTile];
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
                        cS[a][r] = sj[(1 + 2 * nc + 6 * a + r) * kTile];
                        cU[a][r] = sj[(1 + 8 * nc + 6 * a + r) * kTile];
                    }
                }
            }
#pragma unroll
            for (int a = 2; a >= 0; --a) {  // reverse order of the elimination
                if (a < nc) {
                    double qdd = cq[a] - dot6(cU[a], dl);
#pragma unroll
                    for (int r = 0; r < 6; ++r) dl[r] += cS[a][r] * qdd;
                    const double tv = cu[a] - kdmdt * qdd;
                    if (!isfinite(tv)) status |= 1;
                    put_tau(X->col_dof[c0 + a], tv);
                }
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) Wl[r * kTile] = dl[r];
            const int o = X->offset[j], sz = X->size[j];
            for (int i = 0, nd = 0; i < sz; ++i)
                if (X->dof_col[o + i] < 0) {
                    put_tau(o + i, sj[(1 + 14 * nc + nd) * kTile]);
                    ++nd;
                }
#ifdef TRL_PD_CLK
            PD_CLK(clk_i); ++clk_i;
#endif
        }
    }
    if (status) atomicOr(&s_status[lane], status);
    PD_CLK(28);
    __syncthreads();
    PD_CLK(29);
    if (staged) {
        // [32][np] staging rows -> the tile's contiguous [envs][n] block of tau, coalesced. Row / column of a flat index
        // by a multiply-shift (exact for idx < 2^20 / n), four elements in flight per thread.
        const double* st = ws0 + (size_t)kOut * kTile;
        double* out = p.tau + (size_t)e0 * n;
        const int total = envs * n;
        const bool dense = (np == n) && ((reinterpret_cast<size_t>(out) & 15) == 0);
        if (dense) {
            // odd n, 16-byte aligned block: the staging rows ARE the output block -- a flat 16-byte copy with
            // eight elements in flight per thread and no index arithmetic
            const double2* st2 = reinterpret_cast<const double2*>(st);
            double2* out2 = reinterpret_cast<double2*>(out);
            const int pairs = total >> 1;
            for (int base = threadIdx.x; base < pairs; base += 8 * blockDim.x) {
                double2 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (base + k * (int)blockDim.x < pairs) {
                        v[k] = st2[base + k * blockDim.x];
                        if (p.tau_accumulate) {
                            const double2 o = out2[base + k * blockDim.x];
                            v[k].x += o.x; v[k].y += o.y;
                        }
                    }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (base + k * (int)blockDim.x < pairs) out2[base + k * blockDim.x] = v[k];
            }
            if ((total & 1) && threadIdx.x == 0) out[total - 1] = p.tau_accumulate ? out[total - 1] + st[total - 1] : st[total - 1];
        }
        const unsigned magic = (np == n) ? 0u : ((1u << 20) + (unsigned)n - 1u) / (unsigned)n;
        for (int base = threadIdx.x; base < (dense ? 0 : total); base += 4 * blockDim.x) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = base + k * blockDim.x;
                if (idx < total) {
                    int si = idx;  // odd n: the staging rows are dense
                    if (np != n) {
                        const int r = (int)(((unsigned)idx * magic) >> 20);
                        si = idx + r;  // r * np + (idx - r * n) with np = n + 1
                    }
                    v[k] = st[si];
                    if (p.tau_accumulate) v[k] += out[idx];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = base + k * blockDim.x;
                if (idx < total) out[idx] = v[k];
            }
        }
    }
    if (p.status && threadIdx.x < envs) p.status[e0 + threadIdx.x] = s_status[threadIdx.x];
    PD_CLK(30);
}
