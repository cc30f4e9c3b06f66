// Device-side math shared by the kernels: quaternion / rotation helpers (cMathUtil restated for the GPU),
// joint-local transforms (cKinTree::ChildParentTrans*), and heightfield sampling (cGroundVar2D/3D).
#pragma once

#include <cfloat>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "trl_internal.h"

#define TRL_DEV __device__ __forceinline__

// Cooperative copy of the model's integer tables into shared memory (call from every thread of the block, before
// any early return; contains the __syncthreads()).
TRL_DEV void stage_model_idx(ModelIdx* dst, const DevModel* __restrict__ M) {
    const uint4* src = reinterpret_cast<const uint4*>(&M->ix);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(ModelIdx) / 16); i += blockDim.x) d[i] = __ldg(src + i);
    __syncthreads();
}

// ---- 3-vectors / 3x3 (row-major) ----
TRL_DEV void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

TRL_DEV void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

TRL_DEV void mat3_mulv(const double* A, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = A[r * 3] * x[0] + A[r * 3 + 1] * x[1] + A[r * 3 + 2] * x[2];
}

TRL_DEV void mat3T_mulv(const double* A, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = A[r] * x[0] + A[3 + r] * x[1] + A[6 + r] * x[2];
}

// cMathUtil::RotateMat(quaternion): divides by |q|^2 (util/MathUtil.cpp:176-206). q = (w,x,y,z)
TRL_DEV void quat_to_mat(double w, double x, double y, double z, double* R) {
    double sqw = w * w, sqx = x * x, sqy = y * y, sqz = z * z;
    double invs = 1.0 / (sqx + sqy + sqz + sqw);
    R[0] = (sqx - sqy - sqz + sqw) * invs;
    R[4] = (-sqx + sqy - sqz + sqw) * invs;
    R[8] = (-sqx - sqy + sqz + sqw) * invs;
    double t1 = x * y, t2 = z * w;
    R[3] = 2.0 * (t1 + t2) * invs;
    R[1] = 2.0 * (t1 - t2) * invs;
    t1 = x * z; t2 = y * w;
    R[6] = 2.0 * (t1 - t2) * invs;
    R[2] = 2.0 * (t1 + t2) * invs;
    t1 = y * z; t2 = x * w;
    R[7] = 2.0 * (t1 + t2) * invs;
    R[5] = 2.0 * (t1 - t2) * invs;
}

// Hamilton product, (w,x,y,z)
TRL_DEV void quat_mul(const double* a, const double* b, double* r) {
    r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    r[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    r[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    r[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

TRL_DEV void quat_normalize(double* q) {
    double n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
    if (n2 > 0) {  // Eigen divides each component by the norm; rsqrt + 4 products differs by an ulp or two (bar: 1e-9)
        const double inv = rsqrt(n2);  // <= 2 ulp; cheaper than sqrt + division
        q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
    }
}

// Eigen's q * v  (v + w*t + u x t, t = 2 u x v)
TRL_DEV void quat_rot_vec(const double* q, const double* v, double* o) {
    double uv[3], uuv[3];
    cross3(q + 1, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(q + 1, uv, uuv);
    o[0] = v[0] + q[0] * uv[0] + uuv[0];
    o[1] = v[1] + q[0] * uv[1] + uuv[1];
    o[2] = v[2] + q[0] * uv[2] + uuv[2];
}

// cMathUtil::BuildQuaternionDiffMat(q) * w4 (util/MathUtil.cpp:418-426)
TRL_DEV void quat_diff_mul(const double* q, const double* w, double* o) {
    o[0] = -0.5 * q[1] * w[0] + -0.5 * q[2] * w[1] + -0.5 * q[3] * w[2];
    o[1] = 0.5 * q[0] * w[0] + -0.5 * q[3] * w[1] + 0.5 * q[2] * w[2];
    o[2] = 0.5 * q[3] * w[0] + 0.5 * q[0] * w[1] + -0.5 * q[1] * w[2];
    o[3] = -0.5 * q[2] * w[0] + 0.5 * q[1] * w[1] + 0.5 * q[0] * w[2] + w[3];
}

// cMathUtil::CalcQuaternionVelRel (util/MathUtil.cpp:437-445) with QuaternionToAxisAngle's asin form (:399-416, Q3)
TRL_DEV void quat_vel_rel(const double* q0, const double* q1, double dt, double* o) {
    double c[4] = {q0[0], -q0[1], -q0[2], -q0[3]};
    double d[4];
    quat_mul(c, q1, d);
    double st = sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
    // reference: axis = d.xyz / st, out = (theta / dt) * axis (axis = z, theta = 0 below the 1e-4 threshold). One
    // division instead of four: out = (theta / (dt * st)) * d.xyz -- the same value to an ulp or two (bar: 1e-9).
    double k = 0.0;
    if (st > 0.0001) {
        double sg = (double)((0.0 < d[0]) - (d[0] < 0.0));
        k = (2 * sg * asin(st)) / (dt * st);
    }
    o[0] = k * d[1]; o[1] = k * d[2]; o[2] = k * d[3]; o[3] = 0.0;
}

// Eigen 3.3 slerp followed by normalize (anim/KinTree.cpp:1544-1545)
TRL_DEV void quat_slerp_norm(const double* a, double t, const double* b, double* r) {
    const double one = 1.0 - DBL_EPSILON;
    double d = a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[0] * b[0];
    double ad = fabs(d);
    double s0, s1;
    if (ad >= one) {
        s0 = 1.0 - t;
        s1 = t;
    } else {
        // Eigen: theta = acos(|d|), s0 = sin((1-t) theta) / sin(theta), s1 = sin(t theta) / sin(theta). Same values from
        // one sincos instead of three sines: sin(theta) = sqrt((1-|d|)(1+|d|)) (1-|d| is exact for |d| >= 0.5, so this
        // is accurate to an ulp even next to 1) and sin((1-t) theta) = sin(theta) cos(t theta) - |d| sin(t theta).
        // Agreement with the three-sine form is at the 1e-16 level (parity bar: 1e-9).
        const double th = acos(ad);
        const double sn = sqrt((1.0 - ad) * (1.0 + ad));
        double st, ct;
        sincos(t * th, &st, &ct);
        s1 = st / sn;
        s0 = ct - ad * s1;
    }
    if (d < 0) s1 = -s1;
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = s0 * a[i] + s1 * b[i];
    quat_normalize(r);
}

// ---- joint-local child->parent transform [R(9) | t(3)], cKinTree::ChildParentTrans* (anim/KinTree.cpp:1794-1866) ----
// For the root pass world=true to get A*T(pos)*R(quat); otherwise callers treat the root frame as the origin.
TRL_DEV void joint_local_trans(const ModelIdx* X, int j, const double* q /* joint params */,
                               const double* A /* attach rotation, 9 */, const double* at /* attach point, 3 */,
                               double* R, double* t) {
    const int type = (X->parent[j] == -1) ? -1 : X->type[j];
    if (type == -1) {  // root: A * T * R
        double Rq[9];
        quat_to_mat(q[3], q[4], q[5], q[6], Rq);
        mat3_mul(A, Rq, R);
        mat3_mulv(A, q, t);
        t[0] += at[0]; t[1] += at[1]; t[2] += at[2];
    } else if (type == JT_REVOLUTE || type == JT_PLANAR) {
        double th = (type == JT_REVOLUTE) ? q[0] : q[2];
        double s, c;
        sincos(th, &s, &c);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            R[r * 3 + 0] = A[r * 3 + 0] * c + A[r * 3 + 1] * s;
            R[r * 3 + 1] = A[r * 3 + 1] * c - A[r * 3 + 0] * s;
            R[r * 3 + 2] = A[r * 3 + 2];
        }
        if (type == JT_PLANAR) {  // A * R * T
#pragma unroll
            for (int r = 0; r < 3; ++r) t[r] = at[r] + R[r * 3] * q[0] + R[r * 3 + 1] * q[1];
        } else {
            t[0] = at[0]; t[1] = at[1]; t[2] = at[2];
        }
    } else if (type == JT_PRISMATIC) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = A[i];
#pragma unroll
        for (int r = 0; r < 3; ++r) t[r] = at[r] + A[r * 3] * q[0];
    } else if (type == JT_SPHERICAL) {
        double Rq[9];
        quat_to_mat(q[0], q[1], q[2], q[3], Rq);
        mat3_mul(A, Rq, R);
        t[0] = at[0]; t[1] = at[1]; t[2] = at[2];
    } else {  // fixed
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = A[i];
        t[0] = at[0]; t[1] = at[1]; t[2] = at[2];
    }
}

// ---- heightfield sampling. Index arithmetic uses explicit round-to-nearest intrinsics so that nvcc never
// contracts it into FMAs: cell indices must be bit-exact with the reference's double expression
// (sim/GroundVar2D.cpp:652-681, GroundVar3D.cpp:602-621; SURVEY.md "Bit-exact terrain indices"). ----
// cMathUtil::Clamp = max(lo, min(v, hi)) as two compare + select pairs (3 instructions each). FP64 fmin / fmax cost
// ~10 instructions apiece here (NaN canonicalisation sequences): measured 40 instructions per sample for two clamps.
TRL_DEV double clampd(double v, double lo, double hi) {
    const double t = (v < hi) ? v : hi;
    return (lo < t) ? t : lo;
}

// Correctly-rounded x / d from the correctly-rounded reciprocal r = RN(1/d) (computed once per piece on the host):
// Markstein's FMA sequence -- q0 = RN(x r) is within 2 ulp, one residual correction makes it faithful, and a second
// one with the exact residual yields RN(x / d) (Markstein 1990, "Computation of elementary functions on the IBM RISC
// System/6000 processor", Thm. on division by correction with a correctly rounded reciprocal). 5 instructions instead
// of the ~35 of a full IEEE division; bit-identical to `x / d` for normal-range operands, which is asserted over 1e7
// random operands in tests/test_gpu_parity.py. Outside the safe range (inf, NaN, tiny, huge) it falls back to __ddiv_rn.
TRL_DEV double div_by_const(double x, double d, double r) {
    // safe range test on the exponent field (integer pipe, not FP64): 2^-600 < |x| < 2^600; zero, subnormal, inf, NaN fail
    const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
    if (ex - 423u > 1200u) return __ddiv_rn(x, d);
    const double q0 = __dmul_rn(x, r);
    const double e0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(e0, r, q0);
    const double e1 = __fma_rn(-q1, d, x);
    return __fma_rn(e1, r, q1);
}

TRL_DEV double grid_coord(double pos, double origin, double scale, double rscale, double half /* (res - 1) * 0.5 */) {
    double c = __dsub_rn(pos, origin);
    c = div_by_const(c, scale, rscale);
    return __dadd_rn(c, half);
}

// both grid coordinates of a sample with ONE range guard (the common case is two in-range operands)
TRL_DEV void grid_coord2(double px, double pz, const DevPiece& pc, double& c0, double& c1) {
    const double x = __dsub_rn(px, pc.origin[0]), z = __dsub_rn(pz, pc.origin[2]);
    const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu, ez = ((unsigned)__double2hiint(z) >> 20) & 0x7ffu;
    double qx, qz;
    if ((ex - 423u) > 1200u || (ez - 423u) > 1200u) {  // zero, subnormal, huge, inf, NaN: IEEE division
        qx = __ddiv_rn(x, pc.scale[0]);
        qz = __ddiv_rn(z, pc.scale[2]);
    } else {  // Markstein division by the per-piece constant spacing, see div_by_const
        const double rx = pc.rscale[0], rz = pc.rscale[1], dx = pc.scale[0], dz = pc.scale[2];
        const double ax = __dmul_rn(x, rx), az = __dmul_rn(z, rz);
        const double e0x = __fma_rn(-ax, dx, x), e0z = __fma_rn(-az, dz, z);
        const double bx = __fma_rn(e0x, rx, ax), bz = __fma_rn(e0z, rz, az);
        const double e1x = __fma_rn(-bx, dx, x), e1z = __fma_rn(-bz, dz, z);
        qx = __fma_rn(e1x, rx, bx);
        qz = __fma_rn(e1z, rz, bz);
    }
    c0 = __dadd_rn(qx, pc.half[0]);
    c1 = __dadd_rn(qz, pc.half[1]);
}

// A height sample split in two halves so callers can batch the (dependent-address) heightfield loads of several
// samples before combining them: prep computes cell indices / weights, finish applies the reference's interpolation.
struct SamplePrep {
    unsigned int o0, o1, o2;  // element offsets into the float heightfield array (< 2^32 floats; o2 unused in 2-D)
    double w0, w1, w2;     // 2-D: (1-lerp, lerp, -) ; 3-D: barycentric (u, v, w)
    double sy;             // scale.y
    int valid;
    int cell[3];           // piece, i, j
};

TRL_DEV void sample_prep_2d(const DevPiece& pc, int s, double px, double pz, SamplePrep& sp) {
    const double tol = 0.0001;
    const int w = pc.res_x, l = pc.res_z;
    double c0 = grid_coord(px, pc.origin[0], pc.scale[0], pc.rscale[0], pc.half[0]);
    double c1 = grid_coord(pz, pc.origin[2], pc.scale[2], pc.rscale[1], pc.half[1]);
    if (c0 > -tol && c0 < w - 1 + tol && c1 > -tol && c1 < l - 1 + tol) {
        c0 = clampd(c0, 0.0, w - 1.0);
        c1 = clampd(c1, 0.0, l - 1.0);
    }
    const bool outside = (c0 < 0 || c0 > w - 1 || c1 < 0 || c1 > l - 1);
    sp.valid = !outside;
    c0 = clampd(c0, 0.0, w - 1.0);
    const int i = (int)c0;
    const int j = (w - 1 < i + 1) ? (w - 1) : (i + 1);
    const double lerp = __dsub_rn(c0, (double)i);
    sp.o0 = (unsigned int)(pc.data_offset + i);
    sp.o1 = (unsigned int)(pc.data_offset + j);
    sp.o2 = sp.o0;
    sp.w0 = __dsub_rn(1.0, lerp);
    sp.w1 = lerp;
    sp.w2 = 0.0;
    sp.sy = pc.scale[1];
    sp.cell[0] = s; sp.cell[1] = i; sp.cell[2] = j;
}

TRL_DEV double sample_finish_2d(const SamplePrep& sp, float f0, float f1) {
    const double a = __dmul_rn((double)f0, sp.sy);  // GetVertex(i,0)[1] (row 0 only)
    const double b = __dmul_rn((double)f1, sp.sy);
    return __dadd_rn(__dmul_rn(sp.w0, a), __dmul_rn(sp.w1, b));
}

TRL_DEV void sample_prep_3d(const DevPiece& pc, int s, double px, double pz, SamplePrep& sp) {
    const int rx = pc.pitch;  // device row pitch (>= res_x)
    double c0, c1;
    grid_coord2(px, pz, pc, c0, c1);
    c0 = clampd(c0, 0.0, pc.cmax3[0]);  // rx - 1.0 - tol, evaluated once per piece on the host
    c1 = clampd(c1, 0.0, pc.cmax3[1]);
    const int i = (int)c0, j = (int)c1;
    const double lx = __dsub_rn(c0, (double)i), lz = __dsub_rn(c1, (double)j);
    const bool lower = (lz <= lx);
    // lower triangle: (i, j), (i+1, j), (i+1, j+1); upper: (i+1, j+1), (i, j+1), (i, j)
    const unsigned int base = (unsigned int)pc.data_offset + (unsigned int)(j * rx + i);  // < 2^32 floats (checked at registration)
    const unsigned int diag = base + (unsigned int)rx + 1u;
    sp.o0 = lower ? base : diag;
    sp.o1 = lower ? base + 1u : base + (unsigned int)rx;
    sp.o2 = lower ? diag : base;
    // cMathUtil::CalcBarycentric (util/MathUtil.cpp:630-647) on the two unit triangles. With a = (0,0), b = (1,0),
    // c = (1,1) (lower) or a = (1,1), b = (0,1), c = (0,0) (upper) the general expression has d00 = d01 = 1,
    // d11 = 2, denom = 1 exactly, so every product by them and the division are exact and the reference's
    // rounding sequence reduces to the few roundings below (bit-identical, checked by the parity tests).
    double d20, d21;
    if (lower) {
        d20 = lx;                 // v2 = (lx, lz), v0 = (1, 0)
        d21 = __dadd_rn(lx, lz);  // v1 = (1, 1)
    } else {
        const double ax = __dsub_rn(lx, 1.0), az = __dsub_rn(lz, 1.0);  // v2 = p - (1,1)
        d20 = -ax;                                                      // v0 = (-1, 0)
        d21 = __dadd_rn(-ax, -az);                                      // v1 = (-1, -1)
    }
    const double bv = __dsub_rn(__dmul_rn(2.0, d20), d21);
    const double bw = __dsub_rn(d21, d20);
    sp.w0 = __dsub_rn(__dsub_rn(1.0, bv), bw);
    sp.w1 = bv;
    sp.w2 = bw;
    sp.sy = pc.scale[1];
    sp.valid = 1;
    sp.cell[0] = s; sp.cell[1] = i; sp.cell[2] = j;
}

TRL_DEV double sample_finish_3d(const SamplePrep& sp, float f0, float f1, float f2) {
    const double v0 = __dmul_rn((double)f0, sp.sy);
    const double v1 = __dmul_rn((double)f1, sp.sy);
    const double v2 = __dmul_rn((double)f2, sp.sy);
    return __dadd_rn(__dadd_rn(__dmul_rn(sp.w0, v0), __dmul_rn(sp.w1, v1)), __dmul_rn(sp.w2, v2));
}

TRL_DEV double sample_segment_2d(const DevPiece& pc, const float* __restrict__ data, int s, double px, double pz,
                                 int* valid, int* cell) {
    SamplePrep sp;
    sample_prep_2d(pc, s, px, pz, sp);
    *valid = sp.valid;
    cell[0] = sp.cell[0]; cell[1] = sp.cell[1]; cell[2] = sp.cell[2];
    return sample_finish_2d(sp, __ldg(data + sp.o0), __ldg(data + sp.o1));
}

TRL_DEV double sample_slab_3d(const DevPiece& pc, const float* __restrict__ data, int s, double px, double pz,
                              int* valid, int* cell) {
    SamplePrep sp;
    sample_prep_3d(pc, s, px, pz, sp);
    *valid = sp.valid;
    cell[0] = sp.cell[0]; cell[1] = sp.cell[1]; cell[2] = sp.cell[2];
    return sample_finish_3d(sp, __ldg(data + sp.o0), __ldg(data + sp.o1), __ldg(data + sp.o2));
}

TRL_DEV bool piece_contains(int kind, const DevPiece& pc, double px, double pz) {
    if (kind == 1) return px >= pc.bounds[0] && px <= pc.bounds[1];
    return px >= pc.bounds[0] && px <= pc.bounds[1] && pz >= pc.bounds[2] && pz <= pc.bounds[3];
}

// index of the first piece containing (px, pz), or -1 (the reference scans the pieces in order and stops at the
// first). Branch-free: all (<= 4) containment tests are evaluated and the lowest set bit wins.
TRL_DEV int ground_find_piece(int kind, int num_pieces, const DevPiece* pcs, double px, double pz) {
    unsigned mask = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (s < num_pieces && piece_contains(kind, pcs[s], px, pz)) mask |= 1u << s;
    return mask ? (__ffs(mask) - 1) : -1;
}

TRL_DEV double ground_sample_piece(int kind, const DevPiece* pcs, int s, const float* __restrict__ data, double px,
                                   double pz, int* valid, int* cell) {
    if (s < 0) {
        *valid = 0;
        cell[0] = -1; cell[1] = 0; cell[2] = 0;
        return -CUDART_INF;
    }
    return (kind == 1) ? sample_segment_2d(pcs[s], data, s, px, pz, valid, cell)
                       : sample_slab_3d(pcs[s], data, s, px, pz, valid, cell);
}

TRL_DEV double ground_sample_pieces(int kind, int num_pieces, const DevPiece* pcs, const float* __restrict__ data,
                                    double px, double pz, int* valid, int* cell) {
    if (kind == 0 || kind == 3) {  // cGround::SampleHeight (sim/Ground.cpp:176-186)
        *valid = 1;
        cell[0] = -1; cell[1] = 0; cell[2] = 0;
        return 0.0;
    }
    return ground_sample_piece(kind, pcs, ground_find_piece(kind, num_pieces, pcs, px, pz), data, px, pz, valid, cell);
}

TRL_DEV int ground_field_of_env(const DevGround& g, int env) {
    return __ldg(g.env_to_field + env);  // always materialised at registration (env % num_fields by default)
}

TRL_DEV double ground_sample_height(const DevGround& g, int env, double px, double pz, int* valid, int* cell) {
    const DevPiece* pcs = nullptr;
    if (g.kind == 1 || g.kind == 2) pcs = g.pieces + (long long)ground_field_of_env(g, env) * g.pieces_per_field;
    return ground_sample_pieces(g.kind, g.pieces_per_field, pcs, g.data, px, pz, valid, cell);
}
