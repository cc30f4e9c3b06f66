/*
 * trl_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE). See trl_oracle.h.
 *
 * Plain C99, double precision, compile with -O2 -ffp-contract=off so it is a stable
 * reference. Deliberately literal: it keeps the reference's formulation (4x4 homogeneous
 * matrices, dense 6x6 spatial matrices, pivoted LDLT, per-joint chain walks) and order of
 * operations, including the quirks listed in SURVEY.md section 8a-Q.
 *
 * PARITY PINNED TO THE REFERENCE ITSELF: the reference's own translation units (MathUtil, SpAlg, RBDUtil, RBDModel,
 * KinTree, Motion, Character, KinCharacter, the PD controllers, Ground / GroundVar2D / 3D, TerrainGen2D / 3D, Rand ...)
 * are compiled unmodified into oracle/_ref/libtrl_ref.so against stand-in Eigen / Bullet headers (oracle/Makefile target
 * `ref`), and every function here is compared with them on the same inputs at 1e-13 relative (terrain heights, cell
 * indices and the random stream bit for bit): tests/test_ref_pin.py. The same comparisons are frozen as fixtures
 * generated from the reference (tools/make_golden.py -> tests/golden/, tests/test_golden.py) and run everywhere.
 * Independent identities (tests/test_oracle_identities.py, incl. a 50-digit mpmath solve) bound it a second way.
 */
#include "trl_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small fixed-size linear algebra (Eigen Matrix4d / Vector4d / 6x6 semantics)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double m[4][4]; } mat4;
typedef struct { double v[4]; } vec4;
typedef struct { double v[6]; } spvec;
typedef struct { double m[6][6]; } spmat;
typedef struct { double E[3][3]; double r[3]; } sptrans; /* cSpAlg::tSpTrans [E | r], sim/SpAlg.h:9-16 */
typedef struct { double w, x, y, z; } quat;

static mat4 mat4_identity(void)
{
    mat4 a;
    memset(&a, 0, sizeof(a));
    a.m[0][0] = a.m[1][1] = a.m[2][2] = a.m[3][3] = 1;
    return a;
}

static mat4 mat4_mul(const mat4* a, const mat4* b)
{
    mat4 c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = a->m[i][0] * b->m[0][j];
            s += a->m[i][1] * b->m[1][j];
            s += a->m[i][2] * b->m[2][j];
            s += a->m[i][3] * b->m[3][j];
            c.m[i][j] = s;
        }
    return c;
}

static vec4 mat4_mulv(const mat4* a, const vec4* x)
{
    vec4 y;
    for (int i = 0; i < 4; ++i) {
        double s = a->m[i][0] * x->v[0];
        s += a->m[i][1] * x->v[1];
        s += a->m[i][2] * x->v[2];
        s += a->m[i][3] * x->v[3];
        y.v[i] = s;
    }
    return y;
}

static spmat spmat_mul(const spmat* a, const spmat* b)
{
    spmat c;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = a->m[i][0] * b->m[0][j];
            for (int k = 1; k < 6; ++k) s += a->m[i][k] * b->m[k][j];
            c.m[i][j] = s;
        }
    return c;
}

static spvec spmat_mulv(const spmat* a, const spvec* x)
{
    spvec y;
    for (int i = 0; i < 6; ++i) {
        double s = a->m[i][0] * x->v[0];
        for (int k = 1; k < 6; ++k) s += a->m[i][k] * x->v[k];
        y.v[i] = s;
    }
    return y;
}

static void cross3(const double a[3], const double b[3], double out[3])
{
    /* Eigen cross3 */
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

static void mat3_mulv(const double E[3][3], const double x[3], double y[3])
{
    for (int i = 0; i < 3; ++i) {
        double s = E[i][0] * x[0];
        s += E[i][1] * x[1];
        s += E[i][2] * x[2];
        y[i] = s;
    }
}

static void mat3T_mulv(const double E[3][3], const double x[3], double y[3])
{
    for (int i = 0; i < 3; ++i) {
        double s = E[0][i] * x[0];
        s += E[1][i] * x[1];
        s += E[2][i] * x[2];
        y[i] = s;
    }
}

/* ------------------------------------------------------------------------------------------
 * cMathUtil (util/MathUtil.cpp)
 * ---------------------------------------------------------------------------------------- */
static double clampd(double v, double lo, double hi) /* MathUtil.cpp:16-19 */
{
    double t = (v < hi) ? v : hi; /* std::min(val,max) */
    return (lo < t) ? t : lo;     /* std::max(min, .) */
}

static double signd(double v) { return (double)((0.0 < v) - (v < 0.0)); } /* MathUtil.h:123-127 */

/* RotateMat(euler): Rz*Ry*Rx written element-wise, MathUtil.cpp:128-155 (Q6) */
static mat4 rotate_mat_euler(const double e[3])
{
    double x_s = sin(e[0]), x_c = cos(e[0]);
    double y_s = sin(e[1]), y_c = cos(e[1]);
    double z_s = sin(e[2]), z_c = cos(e[2]);
    mat4 a = mat4_identity();
    a.m[0][0] = y_c * z_c;
    a.m[1][0] = y_c * z_s;
    a.m[2][0] = -y_s;
    a.m[0][1] = x_s * y_s * z_c - x_c * z_s;
    a.m[1][1] = x_s * y_s * z_s + x_c * z_c;
    a.m[2][1] = x_s * y_c;
    a.m[0][2] = x_c * y_s * z_c + x_s * z_s;
    a.m[1][2] = x_c * y_s * z_s - x_s * z_c;
    a.m[2][2] = x_c * y_c;
    return a;
}

/* RotateMat(axis, theta), MathUtil.cpp:157-174 */
static mat4 rotate_mat_axis(const double ax[3], double theta)
{
    double c = cos(theta), s = sin(theta);
    double x = ax[0], y = ax[1], z = ax[2];
    mat4 a = mat4_identity();
    a.m[0][0] = c + x * x * (1 - c);     a.m[0][1] = x * y * (1 - c) - z * s; a.m[0][2] = x * z * (1 - c) + y * s;
    a.m[1][0] = y * x * (1 - c) + z * s; a.m[1][1] = c + y * y * (1 - c);     a.m[1][2] = y * z * (1 - c) - x * s;
    a.m[2][0] = z * x * (1 - c) - y * s; a.m[2][1] = z * y * (1 - c) + x * s; a.m[2][2] = c + z * z * (1 - c);
    return a;
}

/* RotateMat(quaternion) -- divides by |q|^2, MathUtil.cpp:176-206 (Q4) */
static mat4 rotate_mat_quat(quat q)
{
    mat4 a = mat4_identity();
    double sqw = q.w * q.w, sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z;
    double invs = 1 / (sqx + sqy + sqz + sqw);
    a.m[0][0] = (sqx - sqy - sqz + sqw) * invs;
    a.m[1][1] = (-sqx + sqy - sqz + sqw) * invs;
    a.m[2][2] = (-sqx - sqy + sqz + sqw) * invs;
    double t1 = q.x * q.y, t2 = q.z * q.w;
    a.m[1][0] = 2.0 * (t1 + t2) * invs;
    a.m[0][1] = 2.0 * (t1 - t2) * invs;
    t1 = q.x * q.z; t2 = q.y * q.w;
    a.m[2][0] = 2.0 * (t1 - t2) * invs;
    a.m[0][2] = 2.0 * (t1 + t2) * invs;
    t1 = q.y * q.z; t2 = q.x * q.w;
    a.m[2][1] = 2.0 * (t1 + t2) * invs;
    a.m[1][2] = 2.0 * (t1 - t2) * invs;
    return a;
}

static mat4 translate_mat(const double t[3]) /* MathUtil.cpp:105-112 */
{
    mat4 a = mat4_identity();
    a.m[0][3] = t[0]; a.m[1][3] = t[1]; a.m[2][3] = t[2];
    return a;
}

static mat4 inv_rigid_mat(const mat4* mat) /* MathUtil.cpp:218-225 */
{
    mat4 inv;
    memset(&inv, 0, sizeof(inv));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) inv.m[i][j] = mat->m[j][i];
    vec4 c = { { mat->m[0][3], mat->m[1][3], mat->m[2][3], mat->m[3][3] } };
    vec4 t = mat4_mulv(&inv, &c);
    for (int i = 0; i < 4; ++i) inv.m[i][3] = -t.v[i];
    inv.m[3][3] = 1;
    return inv;
}

/* RotMatToAxisAngle, MathUtil.cpp:239-260 (threshold 1e-5, Q4) */
static void rot_mat_to_axis_angle(const mat4* mat, double axis[3], double* theta)
{
    double c = (mat->m[0][0] + mat->m[1][1] + mat->m[2][2] - 1) * 0.5;
    c = clampd(c, -1.0, 1.0);
    *theta = acos(c);
    if (fabs(*theta) < 0.00001) {
        axis[0] = 0; axis[1] = 0; axis[2] = 1;
    } else {
        double m21 = mat->m[2][1] - mat->m[1][2];
        double m02 = mat->m[0][2] - mat->m[2][0];
        double m10 = mat->m[1][0] - mat->m[0][1];
        double denom = sqrt(m21 * m21 + m02 * m02 + m10 * m10);
        axis[0] = m21 / denom; axis[1] = m02 / denom; axis[2] = m10 / denom;
    }
}

/* RotMatToQuaternion, MathUtil.cpp:272-307 */
static quat rot_mat_to_quat(const mat4* mat)
{
    double tr = mat->m[0][0] + mat->m[1][1] + mat->m[2][2];
    quat q;
    if (tr > 0) {
        double S = sqrt(tr + 1.0) * 2;
        q.w = 0.25 * S;
        q.x = (mat->m[2][1] - mat->m[1][2]) / S;
        q.y = (mat->m[0][2] - mat->m[2][0]) / S;
        q.z = (mat->m[1][0] - mat->m[0][1]) / S;
    } else if ((mat->m[0][0] > mat->m[1][1]) && (mat->m[0][0] > mat->m[2][2])) {
        double S = sqrt(1.0 + mat->m[0][0] - mat->m[1][1] - mat->m[2][2]) * 2;
        q.w = (mat->m[2][1] - mat->m[1][2]) / S;
        q.x = 0.25 * S;
        q.y = (mat->m[0][1] + mat->m[1][0]) / S;
        q.z = (mat->m[0][2] + mat->m[2][0]) / S;
    } else if (mat->m[1][1] > mat->m[2][2]) {
        double S = sqrt(1.0 + mat->m[1][1] - mat->m[0][0] - mat->m[2][2]) * 2;
        q.w = (mat->m[0][2] - mat->m[2][0]) / S;
        q.x = (mat->m[0][1] + mat->m[1][0]) / S;
        q.y = 0.25 * S;
        q.z = (mat->m[1][2] + mat->m[2][1]) / S;
    } else {
        double S = sqrt(1.0 + mat->m[2][2] - mat->m[0][0] - mat->m[1][1]) * 2;
        q.w = (mat->m[1][0] - mat->m[0][1]) / S;
        q.x = (mat->m[0][2] + mat->m[2][0]) / S;
        q.y = (mat->m[1][2] + mat->m[2][1]) / S;
        q.z = 0.25 * S;
    }
    return q;
}

/* Eigen quaternion product (Hamilton) */
static quat quat_mul(quat a, quat b)
{
    quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}

static quat quat_conj(quat a) { quat r = { a.w, -a.x, -a.y, -a.z }; return r; }

/* Eigen 3.3 Quaternion::normalize(): coeffs /= sqrt(squaredNorm) */
static quat quat_normalize(quat a)
{
    double n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; /* Eigen storage order x,y,z,w */
    if (n2 > 0) {
        double n = sqrt(n2);
        a.x /= n; a.y /= n; a.z /= n; a.w /= n;
    }
    return a;
}

/* Eigen Quaternion * Vector3: v + 2w(u x v) + 2 u x (u x v), as Eigen's _transformVector:
 * uv = u.cross(v); uv += uv; return v + w*uv + u.cross(uv) */
static void quat_rot_vec(quat q, const double v[3], double out[3]) /* MathUtil.cpp:484-489 */
{
    double u[3] = { q.x, q.y, q.z };
    double uv[3];
    cross3(u, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    double uuv[3];
    cross3(u, uv, uuv);
    out[0] = v[0] + q.w * uv[0] + uuv[0];
    out[1] = v[1] + q.w * uv[1] + uuv[1];
    out[2] = v[2] + q.w * uv[2] + uuv[2];
}

/* QuaternionToAxisAngle, MathUtil.cpp:399-416 (Q3: asin of un-normalised vec norm, sign from w) */
static void quat_to_axis_angle(quat q, double axis[3], double* theta)
{
    *theta = 0;
    axis[0] = 0; axis[1] = 0; axis[2] = 1;
    double sin_theta = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (sin_theta > 0.0001) {
        *theta = 2 * signd(q.w) * asin(sin_theta);
        axis[0] = q.x / sin_theta; axis[1] = q.y / sin_theta; axis[2] = q.z / sin_theta;
    }
}

/* AxisAngleToQuaternion, MathUtil.cpp:387-397 */
static quat axis_angle_to_quat(const double axis[3], double theta)
{
    double c = cos(theta / 2), s = sin(theta / 2);
    quat q = { c, s * axis[0], s * axis[1], s * axis[2] };
    return q;
}

/* BuildQuaternionDiffMat(q) * w, MathUtil.cpp:418-426; w is a 4-vector, result ordered (w,x,y,z) */
static void quat_diff_mat_mul(quat q, const double w[4], double out[4])
{
    double m[4][4] = {
        { -0.5 * q.x, -0.5 * q.y, -0.5 * q.z, 0 },
        { 0.5 * q.w, -0.5 * q.z, 0.5 * q.y, 0 },
        { 0.5 * q.z, 0.5 * q.w, -0.5 * q.x, 0 },
        { -0.5 * q.y, 0.5 * q.x, 0.5 * q.w, 1 } };
    for (int i = 0; i < 4; ++i) {
        double s = m[i][0] * w[0];
        s += m[i][1] * w[1];
        s += m[i][2] * w[2];
        s += m[i][3] * w[3];
        out[i] = s;
    }
}

/* CalcQuaternionVelRel, MathUtil.cpp:437-445 */
static void quat_vel_rel(quat q0, quat q1, double dt, double out[4])
{
    quat d = quat_mul(quat_conj(q0), q1);
    double axis[3], theta;
    quat_to_axis_angle(d, axis, &theta);
    double s = theta / dt;
    out[0] = s * axis[0]; out[1] = s * axis[1]; out[2] = s * axis[2]; out[3] = s * 0.0;
}

/* Eigen 3.3 QuaternionBase::slerp (Geometry/Quaternion.h) */
static quat quat_slerp(quat a, double t, quat b)
{
    const double one = 1.0 - DBL_EPSILON;
    double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    double absD = fabs(d);
    double scale0, scale1;
    if (absD >= one) {
        scale0 = 1.0 - t;
        scale1 = t;
    } else {
        double theta = acos(absD);
        double sinTheta = sin(theta);
        scale0 = sin((1.0 - t) * theta) / sinTheta;
        scale1 = sin((t * theta)) / sinTheta;
    }
    if (d < 0) scale1 = -scale1;
    quat r;
    r.x = scale0 * a.x + scale1 * b.x;
    r.y = scale0 * a.y + scale1 * b.y;
    r.z = scale0 * a.z + scale1 * b.z;
    r.w = scale0 * a.w + scale1 * b.w;
    return r;
}

/* CalcBarycentric, MathUtil.cpp:630-647 (inputs have z = w = 0) */
static void calc_barycentric(const double p[2], const double a[2], const double b[2], const double c[2], double out[3])
{
    double v0[2] = { b[0] - a[0], b[1] - a[1] };
    double v1[2] = { c[0] - a[0], c[1] - a[1] };
    double v2[2] = { p[0] - a[0], p[1] - a[1] };
    double d00 = v0[0] * v0[0] + v0[1] * v0[1];
    double d01 = v0[0] * v1[0] + v0[1] * v1[1];
    double d11 = v1[0] * v1[0] + v1[1] * v1[1];
    double d20 = v2[0] * v0[0] + v2[1] * v0[1];
    double d21 = v2[0] * v1[0] + v2[1] * v1[1];
    double denom = d00 * d11 - d01 * d01;
    double v = (d11 * d20 - d01 * d21) / denom;
    double w = (d00 * d21 - d01 * d20) / denom;
    double u = 1.0f - v - w;
    out[0] = u; out[1] = v; out[2] = w;
}

/* ------------------------------------------------------------------------------------------
 * model accessors (anim/KinTree.cpp:783-870)
 * ---------------------------------------------------------------------------------------- */
enum { JD_TYPE = 0, JD_PARENT = 1, JD_ATTACH_X = 2, JD_ATTACH_THETA_X = 5, JD_PARAM_OFFSET = 18 };
enum { BD_SHAPE = 0, BD_MASS = 1, BD_ATTACH_X = 4, BD_ATTACH_THETA_X = 7, BD_PARAM0 = 10 };
enum { PD_KP = 1, PD_KD = 2, PD_USE_WORLD = 17 }; /* sim/PDController.h:12-32 */

static const double* jrow(const trlo_model* m, int j) { return m->joint_mat + (size_t)j * TRLO_JOINT_COLS; }
static const double* brow(const trlo_model* m, int j) { return m->body_defs + (size_t)j * TRLO_BODY_COLS; }

int trlo_joint_type(const trlo_model* m, int j) { return (int)jrow(m, j)[JD_TYPE]; }
int trlo_parent(const trlo_model* m, int j) { return (int)jrow(m, j)[JD_PARENT]; }
int trlo_param_offset(const trlo_model* m, int j) { return (int)jrow(m, j)[JD_PARAM_OFFSET]; }

int trlo_param_size(const trlo_model* m, int j) /* KinTree.cpp:789-823 */
{
    if (trlo_parent(m, j) == -1) return 7;
    switch (trlo_joint_type(m, j)) {
    case TRLO_JOINT_REVOLUTE: return 1;
    case TRLO_JOINT_PRISMATIC: return 1;
    case TRLO_JOINT_PLANAR: return 3;
    case TRLO_JOINT_FIXED: return 0;
    case TRLO_JOINT_SPHERICAL: return 4;
    default: return 0;
    }
}

int trlo_valid_body(const trlo_model* m, int j) /* KinTree.cpp:390-398: shape < eBodyShapeMax(2) */
{
    return ((int)brow(m, j)[BD_SHAPE]) < 2;
}

/* a PD controller is valid iff its cJoint was initialised: non-root with a valid body part
 * (sim/ExpPDController.cpp:23-31, sim/SimCharacter.cpp:779-832, sim/Joint.cpp:41-44) */
int trlo_pd_valid(const trlo_model* m, int j)
{
    return trlo_parent(m, j) != -1 && trlo_valid_body(m, j);
}

static quat pose_quat(const double* p) { quat q = { p[0], p[1], p[2], p[3] }; return q; } /* VecToQuat, MathUtil.cpp:447-450 */

/* ------------------------------------------------------------------------------------------
 * cKinTree transforms (anim/KinTree.cpp:1043-1139,1794-1866)
 * ---------------------------------------------------------------------------------------- */
static mat4 build_attach_trans(const trlo_model* m, int j) /* KinTree.cpp:1043-1053 */
{
    const double* r = jrow(m, j);
    mat4 a = rotate_mat_euler(r + JD_ATTACH_THETA_X);
    a.m[0][3] = r[JD_ATTACH_X]; a.m[1][3] = r[JD_ATTACH_X + 1]; a.m[2][3] = r[JD_ATTACH_X + 2];
    return a;
}

static mat4 child_parent_trans(const trlo_model* m, const double* pose, int j) /* KinTree.cpp:1055-1090 */
{
    mat4 A = build_attach_trans(m, j);
    int off = trlo_param_offset(m, j);
    const double zaxis[3] = { 0, 0, 1 };
    if (trlo_parent(m, j) == -1) { /* ChildParentTransRoot :1794-1805, A*T*R */
        mat4 R = rotate_mat_quat(pose_quat(pose + off + 3));
        mat4 T = translate_mat(pose + off);
        mat4 AT = mat4_mul(&A, &T);
        return mat4_mul(&AT, &R);
    }
    switch (trlo_joint_type(m, j)) {
    case TRLO_JOINT_REVOLUTE: { /* :1807-1817 */
        mat4 R = rotate_mat_axis(zaxis, pose[off]);
        return mat4_mul(&A, &R);
    }
    case TRLO_JOINT_PLANAR: { /* :1819-1833, A*R*T */
        double t[3] = { pose[off], pose[off + 1], 0 };
        mat4 R = rotate_mat_axis(zaxis, pose[off + 2]);
        mat4 T = translate_mat(t);
        mat4 AR = mat4_mul(&A, &R);
        return mat4_mul(&AR, &T);
    }
    case TRLO_JOINT_PRISMATIC: { /* :1842-1853 */
        double t[3] = { pose[off], 0, 0 };
        mat4 T = translate_mat(t);
        return mat4_mul(&A, &T);
    }
    case TRLO_JOINT_SPHERICAL: { /* :1855-1866 */
        mat4 R = rotate_mat_quat(pose_quat(pose + off));
        return mat4_mul(&A, &R);
    }
    case TRLO_JOINT_FIXED:
    default:
        return A; /* :1835-1840 */
    }
}

void trlo_child_parent_trans(const trlo_model* m, const double* pose, int j, double out[16])
{
    mat4 a = child_parent_trans(m, pose, j);
    memcpy(out, a.m, sizeof(a.m));
}

static mat4 joint_world_trans(const trlo_model* m, const double* pose, int j) /* KinTree.cpp:1099-1112 */
{
    mat4 w = mat4_identity();
    int cur = j;
    while (cur != -1) {
        mat4 cp = child_parent_trans(m, pose, cur);
        w = mat4_mul(&cp, &w);
        cur = trlo_parent(m, cur);
    }
    return w;
}

void trlo_joint_world_trans(const trlo_model* m, const double* pose, int j, double out[16])
{
    mat4 a = joint_world_trans(m, pose, j);
    memcpy(out, a.m, sizeof(a.m));
}

static mat4 body_joint_trans(const trlo_model* m, int j) /* KinTree.cpp:1129-1139 (local CoM is zero, :400-416) */
{
    const double* b = brow(m, j);
    mat4 rot = rotate_mat_euler(b + BD_ATTACH_THETA_X);
    double t[3] = { b[BD_ATTACH_X] + 0.0, b[BD_ATTACH_X + 1] + 0.0, b[BD_ATTACH_X + 2] + 0.0 };
    mat4 tr = translate_mat(t);
    return mat4_mul(&tr, &rot);
}

void trlo_body_part_pos(const trlo_model* m, const double* pose, int j, double out[3]) /* KinTree.cpp:299-308 */
{
    mat4 bj = body_joint_trans(m, j);
    mat4 jw = joint_world_trans(m, pose, j);
    vec4 c = { { bj.m[0][3], bj.m[1][3], bj.m[2][3], bj.m[3][3] } };
    vec4 p = mat4_mulv(&jw, &c);
    out[0] = p.v[0]; out[1] = p.v[1]; out[2] = p.v[2];
}

/* VelToPoseDiff, KinTree.cpp:1571-1602 */
void trlo_vel_to_pose_diff(const trlo_model* m, const double* pose, const double* vel, double* out)
{
    memcpy(out, vel, sizeof(double) * (size_t)m->num_dof);
    quat_diff_mat_mul(pose_quat(pose + 3), vel + 3, out + 3);
    for (int j = 1; j < m->num_joints; ++j) {
        if (trlo_joint_type(m, j) == TRLO_JOINT_SPHERICAL) {
            int off = trlo_param_offset(m, j);
            quat_diff_mat_mul(pose_quat(pose + off), vel + off, out + off);
        }
    }
}

static void normalize4(double* v) /* Eigen 3.3 VectorXd segment normalize(): v /= sqrt(squaredNorm) if >0 */
{
    double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
    if (n2 > 0) {
        double n = sqrt(n2);
        v[0] /= n; v[1] /= n; v[2] /= n; v[3] /= n;
    }
}

void trlo_post_process_pose(const trlo_model* m, double* pose) /* PoseProcessPose, KinTree.cpp:1510-1527 */
{
    normalize4(pose + 3);
    for (int j = 1; j < m->num_joints; ++j)
        if (trlo_joint_type(m, j) == TRLO_JOINT_SPHERICAL) normalize4(pose + trlo_param_offset(m, j));
}

void trlo_calc_vel(const trlo_model* m, const double* pose0, const double* pose1, double dt, double* out_vel)
{ /* KinTree.cpp:1470-1508 */
    for (int i = 0; i < 3; ++i) out_vel[i] = (pose1[i] - pose0[i]) / dt;
    quat_vel_rel(pose_quat(pose0 + 3), pose_quat(pose1 + 3), dt, out_vel + 3);
    for (int j = 1; j < m->num_joints; ++j) {
        int off = trlo_param_offset(m, j), size = trlo_param_size(m, j);
        if (trlo_joint_type(m, j) == TRLO_JOINT_SPHERICAL) {
            quat_vel_rel(pose_quat(pose0 + off), pose_quat(pose1 + off), dt, out_vel + off);
        } else {
            for (int i = 0; i < size; ++i) out_vel[off + i] = (pose1[off + i] - pose0[off + i]) / dt;
        }
    }
}

void trlo_lerp_poses(const trlo_model* m, const double* pose0, const double* pose1, double lerp, double* out)
{ /* KinTree.cpp:1529-1569 */
    for (int i = 0; i < 3; ++i) out[i] = (1 - lerp) * pose0[i] + lerp * pose1[i];
    quat q = quat_normalize(quat_slerp(pose_quat(pose0 + 3), lerp, pose_quat(pose1 + 3)));
    out[3] = q.w; out[4] = q.x; out[5] = q.y; out[6] = q.z;
    for (int j = 1; j < m->num_joints; ++j) {
        int off = trlo_param_offset(m, j), size = trlo_param_size(m, j);
        if (trlo_joint_type(m, j) == TRLO_JOINT_SPHERICAL) {
            quat r = quat_normalize(quat_slerp(pose_quat(pose0 + off), lerp, pose_quat(pose1 + off)));
            out[off] = r.w; out[off + 1] = r.x; out[off + 2] = r.y; out[off + 3] = r.z;
        } else {
            for (int i = 0; i < size; ++i) out[off + i] = (1 - lerp) * pose0[off + i] + lerp * pose1[off + i];
        }
    }
}

double trlo_calc_heading(const double* pose) /* KinTree.cpp:1604-1612 */
{
    const double ref[3] = { 1, 0, 0 };
    double rd[3];
    quat_rot_vec(pose_quat(pose + 3), ref, rd);
    return atan2(-rd[2], rd[0]);
}

static mat4 build_origin_trans(const double* pose) /* KinTree.cpp:1620-1639 */
{
    double heading = trlo_calc_heading(pose);
    const double yaxis[3] = { 0, 1, 0 };
    mat4 rot = rotate_mat_axis(yaxis, -heading);
    double neg_origin[3] = { -pose[0], -0.0, -pose[2] };
    mat4 tr = translate_mat(neg_origin);
    return mat4_mul(&rot, &tr);
}

void trlo_build_origin_trans(const double* pose, double out[16])
{
    mat4 a = build_origin_trans(pose);
    memcpy(out, a.m, sizeof(a.m));
}

/* ------------------------------------------------------------------------------------------
 * cSpAlg (sim/SpAlg.cpp)
 * ---------------------------------------------------------------------------------------- */
static sptrans mat_to_trans(const mat4* mat) /* SpAlg.cpp:155-162: E = R, r = -E^T t */
{
    sptrans X;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) X.E[i][j] = mat->m[i][j];
    for (int i = 0; i < 3; ++i) {
        /* (-E^T) * r with r[3] = 0; row 3 of a rigid matrix is (0,0,0,1) */
        double s = (-mat->m[0][i]) * mat->m[0][3];
        s += (-mat->m[1][i]) * mat->m[1][3];
        s += (-mat->m[2][i]) * mat->m[2][3];
        s += (-mat->m[3][i]) * 0.0;
        X.r[i] = s;
    }
    return X;
}

static sptrans inv_trans(const sptrans* X) /* SpAlg.cpp:201-207: (E^T, -E r) */
{
    sptrans Y;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Y.E[i][j] = X->E[j][i];
    for (int i = 0; i < 3; ++i) {
        double s = (-X->E[i][0]) * X->r[0];
        s += (-X->E[i][1]) * X->r[1];
        s += (-X->E[i][2]) * X->r[2];
        Y.r[i] = s;
    }
    return Y;
}

static sptrans comp_trans(const sptrans* X0, const sptrans* X1) /* SpAlg.cpp:337-346: (E0 E1, r1 + E1^T r0) */
{
    sptrans X;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = X0->E[i][0] * X1->E[0][j];
            s += X0->E[i][1] * X1->E[1][j];
            s += X0->E[i][2] * X1->E[2][j];
            X.E[i][j] = s;
        }
    double t[3];
    mat3T_mulv(X1->E, X0->r, t);
    for (int i = 0; i < 3; ++i) X.r[i] = X1->r[i] + t[i];
    return X;
}

static sptrans identity_trans(void)
{
    sptrans X;
    memset(&X, 0, sizeof(X));
    X.E[0][0] = X.E[1][1] = X.E[2][2] = 1;
    return X;
}

static spvec apply_trans_m(const sptrans* X, const spvec* sv) /* SpAlg.cpp:233-245 */
{
    double rxo[3], t[3];
    spvec out;
    cross3(X->r, sv->v, rxo);
    for (int i = 0; i < 3; ++i) t[i] = sv->v[3 + i] - rxo[i];
    mat3_mulv(X->E, sv->v, out.v);
    mat3_mulv(X->E, t, out.v + 3);
    return out;
}

static spvec apply_trans_f(const sptrans* X, const spvec* sv) /* SpAlg.cpp:247-259 */
{
    double rxv[3], t[3];
    spvec out;
    cross3(X->r, sv->v + 3, rxv);
    for (int i = 0; i < 3; ++i) t[i] = sv->v[i] - rxv[i];
    mat3_mulv(X->E, t, out.v);
    mat3_mulv(X->E, sv->v + 3, out.v + 3);
    return out;
}

static spvec cross_m(const spvec* sv, const spvec* mv) /* SpAlg.cpp:50-61 */
{
    spvec out;
    double a[3], b[3];
    cross3(sv->v, mv->v, out.v);
    cross3(sv->v + 3, mv->v, a);
    cross3(sv->v, mv->v + 3, b);
    for (int i = 0; i < 3; ++i) out.v[3 + i] = a[i] + b[i];
    return out;
}

static spvec cross_f(const spvec* sv, const spvec* f) /* SpAlg.cpp:75-86 */
{
    spvec out;
    double a[3], b[3];
    cross3(sv->v, f->v, a);
    cross3(sv->v + 3, f->v + 3, b);
    for (int i = 0; i < 3; ++i) out.v[i] = a[i] + b[i];
    cross3(sv->v, f->v + 3, out.v + 3);
    return out;
}

/* E * CrossMat(r), 3x3 block (SpAlg.cpp:174-198) */
static void e_cross_r(const sptrans* X, double Er[3][3])
{
    double C[3][3] = { { 0, -X->r[2], X->r[1] }, { X->r[2], 0, -X->r[0] }, { -X->r[1], X->r[0], 0 } };
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = X->E[i][0] * C[0][j];
            s += X->E[i][1] * C[1][j];
            s += X->E[i][2] * C[2][j];
            Er[i][j] = s;
        }
}

static spmat build_spatial_mat_m(const sptrans* X) /* SpAlg.cpp:174-185 */
{
    spmat M;
    double Er[3][3];
    memset(&M, 0, sizeof(M));
    e_cross_r(X, Er);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            M.m[i][j] = X->E[i][j];
            M.m[3 + i][3 + j] = X->E[i][j];
            M.m[3 + i][j] = -Er[i][j];
        }
    return M;
}

static spmat build_spatial_mat_f(const sptrans* X) /* SpAlg.cpp:187-198 */
{
    spmat M;
    double Er[3][3];
    memset(&M, 0, sizeof(M));
    e_cross_r(X, Er);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            M.m[i][j] = X->E[i][j];
            M.m[3 + i][3 + j] = X->E[i][j];
            M.m[i][3 + j] = -Er[i][j];
        }
    return M;
}

/* ------------------------------------------------------------------------------------------
 * cRBDUtil inertia (sim/RBDUtil.cpp:600-684)
 * ---------------------------------------------------------------------------------------- */
static spmat build_moment_inertia(const trlo_model* m, int j)
{
    const double* b = brow(m, j);
    double mass = b[BD_MASS];
    spmat I;
    memset(&I, 0, sizeof(I));
    double x, y, z;
    if ((int)b[BD_SHAPE] == 0) { /* box :623-644 */
        double sx = b[BD_PARAM0], sy = b[BD_PARAM0 + 1], sz = b[BD_PARAM0 + 2];
        x = mass / 12.0 * (sy * sy + sz * sz);
        y = mass / 12.0 * (sx * sx + sz * sz);
        z = mass / 12.0 * (sx * sx + sy * sy);
    } else { /* capsule :646-673, including the cm*cm term (Q5) */
        double r = b[BD_PARAM0], h = b[BD_PARAM0 + 1];
        double c_vol = M_PI * r * r * h;
        double hs_vol = M_PI * 2.0 / 3.0 * r * r * r;
        double density = mass / (c_vol + 2 * hs_vol);
        double cm = c_vol * density;
        double hsm = hs_vol * density;
        x = cm * (0.25 * r * r + (1.0 / 12.0) * cm * h * h) +
            2 * hsm * (0.4 * r * r + 0.375 * r * h + 0.25 * h * h);
        y = (0.5 * cm + 0.8 * hsm) * r * r;
        z = x;
    }
    I.m[0][0] = x; I.m[1][1] = y; I.m[2][2] = z;
    I.m[3][3] = mass; I.m[4][4] = mass; I.m[5][5] = mass;
    return I;
}

static spmat build_inertia_spatial_mat(const trlo_model* m, int j) /* RBDUtil.cpp:675-684 */
{
    spmat Ic = build_moment_inertia(m, j);
    const double* b = brow(m, j);
    sptrans X = identity_trans();
    X.r[0] = -b[BD_ATTACH_X]; X.r[1] = -b[BD_ATTACH_X + 1]; X.r[2] = -b[BD_ATTACH_X + 2];
    sptrans Xi = inv_trans(&X);
    spmat F = build_spatial_mat_f(&X);
    spmat Mm = build_spatial_mat_m(&Xi);
    spmat FI = spmat_mul(&F, &Ic);
    return spmat_mul(&FI, &Mm);
}

void trlo_inertia_spatial(const trlo_model* m, int j, double out[36])
{
    spmat I = build_inertia_spatial_mat(m, j);
    memcpy(out, I.m, sizeof(I.m));
}

/* ------------------------------------------------------------------------------------------
 * cRBDModel cache (sim/RBDModel.cpp:39-55,192-283)
 * ---------------------------------------------------------------------------------------- */
typedef struct rbd_cache {
    double S[6][TRLO_MAX_DOF];            /* mJointSubspaceArr 6 x n */
    mat4 child_parent[TRLO_MAX_JOINTS];   /* mChildParentMatArr */
    sptrans world_joint[TRLO_MAX_JOINTS]; /* mSpWorldJointTransArr */
} rbd_cache;

static sptrans sp_child_parent(const rbd_cache* c, int j) { return mat_to_trans(&c->child_parent[j]); } /* RBDModel.cpp:125-129 */
static sptrans sp_parent_child(const rbd_cache* c, int j) /* RBDModel.cpp:118-123,131-135 */
{
    mat4 inv = inv_rigid_mat(&c->child_parent[j]);
    return mat_to_trans(&inv);
}

/* BuildJointSubspace*, RBDUtil.cpp:733-828 */
static void build_joint_subspace(const trlo_model* m, const double* pose, int j, rbd_cache* c)
{
    int off = trlo_param_offset(m, j), dim = trlo_param_size(m, j);
    for (int r = 0; r < 6; ++r)
        for (int k = 0; k < dim; ++k) c->S[r][off + k] = 0;
    if (trlo_parent(m, j) == -1) { /* root :770-784 */
        mat4 E = rotate_mat_quat(pose_quat(pose + off + 3));
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) c->S[3 + r][off + k] = E.m[k][r];
        /* S.block(0, 3, 3, 4).setIdentity(): 3x4 identity -> column 6 stays zero (Q1) */
        for (int r = 0; r < 3; ++r) c->S[r][off + 3 + r] = 1;
        return;
    }
    switch (trlo_joint_type(m, j)) {
    case TRLO_JOINT_REVOLUTE: c->S[2][off] = 1; break;
    case TRLO_JOINT_PRISMATIC: c->S[3][off] = 1; break;
    case TRLO_JOINT_PLANAR: c->S[3][off] = 1; c->S[4][off + 1] = 1; c->S[2][off + 2] = 1; break;
    case TRLO_JOINT_SPHERICAL: c->S[0][off] = 1; c->S[1][off + 1] = 1; c->S[2][off + 2] = 1; break;
    default: break;
    }
}

static void rbd_build_cache(const trlo_model* m, const double* pose, rbd_cache* c)
{
    int N = m->num_joints;
    for (int j = 0; j < N; ++j) build_joint_subspace(m, pose, j, c);              /* UpdateJointSubspaceArr */
    for (int j = 0; j < N; ++j) c->child_parent[j] = child_parent_trans(m, pose, j); /* UpdateChildParentMatArr */
    for (int j = 0; j < N; ++j) {                                                  /* CalcWorldJointTransforms, RBDUtil.cpp:686-710 */
        sptrans pc = sp_parent_child(c, j);
        sptrans wp = identity_trans();
        int p = trlo_parent(m, j);
        if (p != -1) wp = c->world_joint[p];
        c->world_joint[j] = comp_trans(&pc, &wp);
    }
}

/* S_j * x  (6 x dim times dim) */
static spvec subspace_mul(const trlo_model* m, const rbd_cache* c, int j, const double* x)
{
    int off = trlo_param_offset(m, j), dim = trlo_param_size(m, j);
    spvec out;
    for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int k = 0; k < dim; ++k) {
            if (k == 0) s = c->S[r][off] * x[0];
            else s += c->S[r][off + k] * x[k];
        }
        out.v[r] = s;
    }
    return out;
}

/* BuildCjRoot, RBDUtil.cpp:867-904 */
static spvec build_cj_root(const double* q, const double* qd)
{
    quat quatv = pose_quat(q + 3);
    double dqv[4];
    quat_diff_mat_mul(quatv, qd + 3, dqv);
    quat dq = { dqv[0], dqv[1], dqv[2], dqv[3] };
    mat4 mt = mat4_identity();
    mt.m[0][0] = 4 * (quatv.w * dq.w + quatv.x * dq.x);
    mt.m[1][1] = 4 * (quatv.w * dq.w + quatv.y * dq.y);
    mt.m[2][2] = 4 * (quatv.w * dq.w + quatv.z * dq.z);
    mt.m[1][0] = 2 * (dq.x * quatv.y + quatv.x * dq.y - dq.w * quatv.z - quatv.w * dq.z);
    mt.m[0][1] = 2 * (dq.x * quatv.y + quatv.x * dq.y + dq.w * quatv.z + quatv.w * dq.z);
    mt.m[2][0] = 2 * (dq.x * quatv.z + quatv.x * dq.z + dq.w * quatv.y + quatv.w * dq.y);
    mt.m[0][2] = 2 * (dq.x * quatv.z + quatv.x * dq.z - dq.w * quatv.y - quatv.w * dq.y);
    mt.m[2][1] = 2 * (dq.y * quatv.z + quatv.y * dq.z - dq.w * quatv.x - quatv.w * dq.x);
    mt.m[1][2] = 2 * (dq.y * quatv.z + quatv.y * dq.z + dq.w * quatv.x + quatv.w * dq.x);
    vec4 vl = { { qd[0], qd[1], qd[2], 0 } };
    vec4 r = mat4_mulv(&mt, &vl);
    spvec cj;
    memset(&cj, 0, sizeof(cj));
    cj.v[3] = r.v[0]; cj.v[4] = r.v[1]; cj.v[5] = r.v[2];
    return cj;
}

/* SolveInvDyna, RBDUtil.cpp:4-87 */
static void solve_inv_dyna(const trlo_model* m, const rbd_cache* c, const double* pose, const double* vel,
                           const double* acc, double* out_tau)
{
    int N = m->num_joints, n = m->num_dof;
    spvec vels[TRLO_MAX_JOINTS], accs[TRLO_MAX_JOINTS], fs[TRLO_MAX_JOINTS];
    spvec vel0, acc0;
    memset(&vel0, 0, sizeof(vel0));
    memset(&acc0, 0, sizeof(acc0));
    acc0.v[3] = -m->gravity[0]; acc0.v[4] = -m->gravity[1]; acc0.v[5] = -m->gravity[2];
    memset(fs, 0, sizeof(spvec) * (size_t)N);
    memset(vels, 0, sizeof(spvec) * (size_t)N);
    memset(accs, 0, sizeof(spvec) * (size_t)N);

    for (int j = 0; j < N; ++j) {
        if (!trlo_valid_body(m, j)) continue;
        int off = trlo_param_offset(m, j);
        sptrans pc = sp_parent_child(c, j);
        double zero1[1] = { 0 };
        const double* dq = trlo_param_size(m, j) > 0 ? vel + off : zero1;
        const double* ddq = trlo_param_size(m, j) > 0 ? acc + off : zero1;
        spvec cj;
        if (trlo_parent(m, j) == -1) cj = build_cj_root(pose + off, vel + off);
        else memset(&cj, 0, sizeof(cj));
        spvec vj = subspace_mul(m, c, j, dq);
        spmat I = build_inertia_spatial_mat(m, j);
        int p = trlo_parent(m, j);
        const spvec* vp = (p != -1) ? &vels[p] : &vel0;
        const spvec* ap = (p != -1) ? &accs[p] : &acc0;
        spvec cv = apply_trans_m(&pc, vp);
        for (int i = 0; i < 6; ++i) cv.v[i] = cv.v[i] + vj.v[i];
        spvec ca = apply_trans_m(&pc, ap);
        spvec sa = subspace_mul(m, c, j, ddq);
        spvec cx = cross_m(&cv, &vj);
        for (int i = 0; i < 6; ++i) ca.v[i] = ((ca.v[i] + sa.v[i]) + cj.v[i]) + cx.v[i];
        spvec Ia = spmat_mulv(&I, &ca);
        spvec Iv = spmat_mulv(&I, &cv);
        spvec cf = cross_f(&cv, &Iv);
        for (int i = 0; i < 6; ++i) fs[j].v[i] = Ia.v[i] + cf.v[i];
        vels[j] = cv;
        accs[j] = ca;
    }

    for (int i = 0; i < n; ++i) out_tau[i] = 0;
    for (int j = N - 1; j >= 0; --j) {
        if (!trlo_valid_body(m, j)) continue;
        int off = trlo_param_offset(m, j), dim = trlo_param_size(m, j);
        for (int k = 0; k < dim; ++k) { /* S^T f */
            double s = c->S[0][off + k] * fs[j].v[0];
            for (int r = 1; r < 6; ++r) s += c->S[r][off + k] * fs[j].v[r];
            out_tau[off + k] = s;
        }
        int p = trlo_parent(m, j);
        if (p != -1) {
            sptrans cp = sp_child_parent(c, j);
            spvec t = apply_trans_f(&cp, &fs[j]);
            for (int i = 0; i < 6; ++i) fs[p].v[i] += t.v[i];
        }
    }
}

/* BuildMassMat (CRBA), RBDUtil.cpp:113-179 */
static void build_mass_mat(const trlo_model* m, const rbd_cache* c, double* H)
{
    int N = m->num_joints, n = m->num_dof;
    static const spmat zero_spmat; /* mInertiaBuffer is zero-initialised at Init (RBDModel.cpp:36) */
    spmat Is[TRLO_MAX_JOINTS], cpF[TRLO_MAX_JOINTS], pcM[TRLO_MAX_JOINTS];
    for (int i = 0; i < n * n; ++i) H[i] = 0;
    for (int j = 0; j < N; ++j) {
        Is[j] = zero_spmat;
        if (trlo_valid_body(m, j)) Is[j] = build_inertia_spatial_mat(m, j);
        sptrans cp = sp_child_parent(c, j);
        sptrans cpi = inv_trans(&cp);
        cpF[j] = build_spatial_mat_f(&cp);
        pcM[j] = build_spatial_mat_m(&cpi);
    }
    for (int j = N - 1; j >= 0; --j) {
        if (!trlo_valid_body(m, j)) continue;
        int p = trlo_parent(m, j);
        if (p != -1) {
            spmat t = spmat_mul(&cpF[j], &Is[j]);
            spmat u = spmat_mul(&t, &pcM[j]);
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) Is[p].m[a][b] += u.m[a][b];
        }
        int off = trlo_param_offset(m, j), dim = trlo_param_size(m, j);
        double F[6][8]; /* 6 x dim, dim <= 7 */
        for (int r = 0; r < 6; ++r)
            for (int k = 0; k < dim; ++k) {
                double s = Is[j].m[r][0] * c->S[0][off + k];
                for (int q = 1; q < 6; ++q) s += Is[j].m[r][q] * c->S[q][off + k];
                F[r][k] = s;
            }
        for (int a = 0; a < dim; ++a)
            for (int b = 0; b < dim; ++b) {
                double s = c->S[0][off + a] * F[0][b];
                for (int q = 1; q < 6; ++q) s += c->S[q][off + a] * F[q][b];
                H[(off + a) * n + (off + b)] = s;
            }
        int cur = j;
        while (trlo_parent(m, cur) != -1) {
            double F2[6][8];
            for (int r = 0; r < 6; ++r)
                for (int k = 0; k < dim; ++k) {
                    double s = cpF[cur].m[r][0] * F[0][k];
                    for (int q = 1; q < 6; ++q) s += cpF[cur].m[r][q] * F[q][k];
                    F2[r][k] = s;
                }
            memcpy(F, F2, sizeof(F));
            cur = trlo_parent(m, cur);
            int coff = trlo_param_offset(m, cur), cdim = trlo_param_size(m, cur);
            for (int a = 0; a < dim; ++a)
                for (int b = 0; b < cdim; ++b) {
                    double s = F[0][a] * c->S[0][coff + b];
                    for (int q = 1; q < 6; ++q) s += F[q][a] * c->S[q][coff + b];
                    H[(off + a) * n + (coff + b)] = s;
                    H[(coff + b) * n + (off + a)] = s;
                }
        }
    }
}

void trlo_rbd_update(const trlo_model* m, const double* pose, const double* vel, double* out_mass, double* out_bias)
{
    rbd_cache c;
    rbd_build_cache(m, pose, &c);
    if (out_mass) build_mass_mat(m, &c, out_mass);
    if (out_bias) { /* BuildBiasForce = SolveInvDyna(acc = 0), RBDUtil.cpp:931-935 */
        double zero[TRLO_MAX_DOF];
        memset(zero, 0, sizeof(zero));
        solve_inv_dyna(m, &c, pose, vel, zero, out_bias);
    }
}

void trlo_solve_inv_dyna(const trlo_model* m, const double* pose, const double* vel, const double* acc, double* out_tau)
{
    rbd_cache c;
    rbd_build_cache(m, pose, &c);
    solve_inv_dyna(m, &c, pose, vel, acc, out_tau);
}

/* ------------------------------------------------------------------------------------------
 * RBD extras used by the FSM controllers every substep
 * ---------------------------------------------------------------------------------------- */
static spvec apply_inv_trans_m(const sptrans* X, const spvec* sv) /* SpAlg.cpp:285-297 */
{
    double o1[3], v1[3], c[3];
    mat3T_mulv(X->E, sv->v, o1);
    mat3T_mulv(X->E, sv->v + 3, v1);
    cross3(X->r, o1, c);
    spvec out = { { o1[0], o1[1], o1[2], v1[0] + c[0], v1[1] + c[1], v1[2] + c[2] } };
    return out;
}

/* cRBDUtil::BuildJacobian, RBDUtil.cpp:256-275: column block of joint j = ApplyInvTransM(world_joint_trans, S_j).
 * out_J row-major [6][n]; the dead slots (root 6, spherical 3) stay zero columns. */
static void build_jacobian(const trlo_model* m, const rbd_cache* c, double* J)
{
    int n = m->num_dof, N = m->num_joints;
    for (int i = 0; i < 6 * n; ++i) J[i] = 0;
    for (int j = 0; j < N; ++j) {
        int off = trlo_param_offset(m, j), dim = trlo_param_size(m, j);
        for (int k = 0; k < dim; ++k) {
            spvec s, w;
            for (int r = 0; r < 6; ++r) s.v[r] = c->S[r][off + k];
            w = apply_inv_trans_m(&c->world_joint[j], &s);
            for (int r = 0; r < 6; ++r) J[r * n + off + k] = w.v[r];
        }
    }
}

static mat4 trans_to_mat(const sptrans* X) /* SpAlg.cpp:164-172 */
{
    mat4 mt = mat4_identity();
    double Er[3];
    mat3_mulv(X->E, X->r, Er);
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) mt.m[r][k] = X->E[r][k];
        mt.m[r][3] = -Er[r];
    }
    return mt;
}

/* cRBDUtil::BuildCOMJacobian, RBDUtil.cpp:294-345 (the comment there says "origin at the com": each body's
 * block is the world Jacobian shifted to that body's position, weighted by its mass fraction) */
static void build_com_jacobian(const trlo_model* m, const rbd_cache* c, const double* J, double* Jc)
{
    int n = m->num_dof, N = m->num_joints;
    for (int i = 0; i < 6 * n; ++i) Jc[i] = 0;
    double total_mass = 0; /* cKinTree::CalcTotalMass, KinTree.cpp:376-388 */
    for (int j = 0; j < N; ++j)
        if (trlo_valid_body(m, j)) total_mass += brow(m, j)[BD_MASS];
    for (int j = N - 1; j >= 0; --j) {
        if (!trlo_valid_body(m, j)) continue;
        double mass_frac = brow(m, j)[BD_MASS] / total_mass;
        mat4 bj = body_joint_trans(m, j);
        sptrans inv_w = inv_trans(&c->world_joint[j]);
        sptrans bjt = mat_to_trans(&bj);
        sptrans bw = comp_trans(&inv_w, &bjt);
        mat4 bwm = trans_to_mat(&bw);
        sptrans shift = identity_trans();
        shift.r[0] = bwm.m[0][3]; shift.r[1] = bwm.m[1][3]; shift.r[2] = bwm.m[2][3];
        for (int cur = j; cur != -1; cur = trlo_parent(m, cur)) {
            int off = trlo_param_offset(m, cur), dim = trlo_param_size(m, cur);
            for (int k = 0; k < dim; ++k) {
                spvec col, t;
                for (int r = 0; r < 6; ++r) col.v[r] = J[r * n + off + k];
                t = apply_trans_m(&shift, &col);
                for (int r = 0; r < 6; ++r) Jc[r * n + off + k] += mass_frac * t.v[r];
            }
        }
    }
}

/* cRBDUtil::CalcGravityForce, RBDUtil.cpp:937-982. Note acc0 = (0, +gravity), and that an invalid body neither gets a
 * torque nor hands its children's forces on to its parent. */
static void calc_gravity_force(const trlo_model* m, const rbd_cache* c, double* out)
{
    int n = m->num_dof, N = m->num_joints;
    spvec fs[TRLO_MAX_JOINTS];
    memset(fs, 0, sizeof(fs)); /* the reference leaves the rows of invalid bodies uninitialised; they are never read */
    spvec acc0 = { { 0, 0, 0, m->gravity[0], m->gravity[1], m->gravity[2] } };
    for (int j = 0; j < N; ++j) {
        if (!trlo_valid_body(m, j)) continue;
        spmat I = build_inertia_spatial_mat(m, j);
        spvec a = apply_trans_m(&c->world_joint[j], &acc0);
        fs[j] = spmat_mulv(&I, &a);
    }
    for (int i = 0; i < n; ++i) out[i] = 0;
    for (int j = N - 1; j >= 0; --j) {
        if (!trlo_valid_body(m, j)) continue;
        int off = trlo_param_offset(m, j), dim = trlo_param_size(m, j);
        for (int k = 0; k < dim; ++k) { /* S^T f */
            double sacc = 0;
            for (int r = 0; r < 6; ++r) {
                if (r == 0) sacc = c->S[0][off + k] * fs[j].v[0];
                else sacc += c->S[r][off + k] * fs[j].v[r];
            }
            out[off + k] = sacc;
        }
        int p = trlo_parent(m, j);
        if (p != -1) {
            sptrans cp = sp_child_parent(c, j);
            spvec t = apply_trans_f(&cp, &fs[j]);
            for (int r = 0; r < 6; ++r) fs[p].v[r] += t.v[r];
        }
    }
}

/* cRBDUtil::BuildEndEffectorJacobian(model, joint_id), RBDUtil.cpp:181-206: blocks of the joints on the chain
 * root .. joint_id, expressed in world coordinates (the other columns are zero). */
static void build_end_eff_jacobian(const trlo_model* m, const rbd_cache* c, int joint_id, double* J)
{
    int n = m->num_dof;
    for (int i = 0; i < 6 * n; ++i) J[i] = 0;
    sptrans cur = identity_trans();
    for (int id = joint_id; id != -1; id = trlo_parent(m, id)) {
        int off = trlo_param_offset(m, id), dim = trlo_param_size(m, id);
        for (int k = 0; k < dim; ++k) {
            spvec s, t;
            for (int r = 0; r < 6; ++r) s.v[r] = c->S[r][off + k];
            t = apply_trans_m(&cur, &s);
            for (int r = 0; r < 6; ++r) J[r * n + off + k] = t.v[r];
        }
        sptrans pc = sp_parent_child(c, id);
        cur = comp_trans(&cur, &pc);
    }
    for (int col = 0; col < n; ++col) { /* out_J = ApplyInvTransM(curr_trans, out_J) */
        spvec s, t;
        for (int r = 0; r < 6; ++r) s.v[r] = J[r * n + col];
        t = apply_inv_trans_m(&cur, &s);
        for (int r = 0; r < 6; ++r) J[r * n + col] = t.v[r];
    }
}

void trlo_rbd_extras(const trlo_model* m, const double* pose, double* out_J, double* out_Jcom, double* out_gravity_force)
{
    rbd_cache c;
    double Jtmp[6 * TRLO_MAX_DOF];
    rbd_build_cache(m, pose, &c);
    double* J = out_J ? out_J : Jtmp;
    if (out_J || out_Jcom) build_jacobian(m, &c, J);
    if (out_Jcom) build_com_jacobian(m, &c, J, out_Jcom);
    if (out_gravity_force) calc_gravity_force(m, &c, out_gravity_force);
}

void trlo_end_eff_jacobian(const trlo_model* m, const double* pose, int joint_id, double* out_J)
{
    rbd_cache c;
    rbd_build_cache(m, pose, &c);
    build_end_eff_jacobian(m, &c, joint_id, out_J);
}

/* ------------------------------------------------------------------------------------------
 * Torque post-processing: cSimCharacter::ApplyControlForces (sim/SimCharacter.cpp:589-603) ->
 * cJoint::AddTau* (sim/Joint.cpp:536-562) -> cJoint::ApplyTau / ApplyTorque / ApplyForce with ClampTotalTorque /
 * ClampTotalForce (sim/Joint.cpp:162-221,300-318): the generalized torques become clamped joint wrenches and then
 * equal-and-opposite world-frame torques / forces on the parent and child bodies (what Bullet consumes).
 * The joint's world rotation (cJoint::BuildWorldTrans = child body transform x joint-child transform) is taken from
 * FK of `pose`, which is what it equals when the simulated bodies are consistent with the pose.
 * out_tau [n]: the clamped generalized torques (same layout as tau; untouched slots are copied);
 * out_body_wrench [N][6]: per body (torque xyz, force xyz) in world coordinates.
 * ---------------------------------------------------------------------------------------- */
enum { JD_TORQUE_LIM = 14, JD_FORCE_LIM = 15 };

void trlo_apply_control_forces(const trlo_model* m, const double* pose, const double* tau, double* out_tau,
                               double* out_body_wrench)
{
    int N = m->num_joints, n = m->num_dof;
    if (out_tau) for (int i = 0; i < n; ++i) out_tau[i] = tau[i];
    if (out_body_wrench) for (int i = 0; i < 6 * N; ++i) out_body_wrench[i] = 0;
    for (int j = 0; j < N; ++j) {
        int p = trlo_parent(m, j);
        if (p == -1 || !trlo_valid_body(m, j)) continue; /* cJoint::IsValid: a constraint with a child body */
        int off = trlo_param_offset(m, j), type = trlo_joint_type(m, j);
        double tot[6] = { 0, 0, 0, 0, 0, 0 }; /* mTotalTau = [torque ; force] */
        int has_torque = 0, has_force = 0;
        switch (type) {
        case TRLO_JOINT_REVOLUTE: tot[2] += tau[off]; has_torque = 1; break;
        case TRLO_JOINT_PLANAR: /* note: the planar torque lands in slot 0 (Joint.cpp:541-546) */
            tot[3] += tau[off]; tot[4] += tau[off + 1]; tot[0] += tau[off + 2]; has_torque = 1; has_force = 1; break;
        case TRLO_JOINT_PRISMATIC: tot[3] += tau[off]; has_force = 1; break;
        case TRLO_JOINT_SPHERICAL: tot[0] += tau[off]; tot[1] += tau[off + 1]; tot[2] += tau[off + 2]; has_torque = 1; break;
        default: break; /* fixed */
        }
        mat4 W = joint_world_trans(m, pose, j);
        if (has_torque) { /* ApplyTorque */
            double mag = sqrt(tot[0] * tot[0] + tot[1] * tot[1] + tot[2] * tot[2]);
            double lim = jrow(m, j)[JD_TORQUE_LIM];
            if (mag > lim) { double sc = lim / mag; tot[0] *= sc; tot[1] *= sc; tot[2] *= sc; }
            vec4 t = { { tot[0], tot[1], tot[2], 0 } };
            vec4 w = mat4_mulv(&W, &t);
            if (out_body_wrench)
                for (int k = 0; k < 3; ++k) { out_body_wrench[6 * p + k] += -w.v[k]; out_body_wrench[6 * j + k] += w.v[k]; }
        }
        if (has_force) { /* ApplyForce */
            double mag = sqrt(tot[3] * tot[3] + tot[4] * tot[4] + tot[5] * tot[5]);
            double lim = jrow(m, j)[JD_FORCE_LIM];
            if (mag > lim) { double sc = lim / mag; tot[3] *= sc; tot[4] *= sc; tot[5] *= sc; }
            vec4 f = { { tot[3], tot[4], tot[5], 0 } };
            vec4 w = mat4_mulv(&W, &f);
            if (out_body_wrench)
                for (int k = 0; k < 3; ++k) { out_body_wrench[6 * p + 3 + k] += -w.v[k]; out_body_wrench[6 * j + 3 + k] += w.v[k]; }
        }
        if (out_tau) { /* SetTotalTorque / SetTotalForce written back in the generalized layout */
            switch (type) {
            case TRLO_JOINT_REVOLUTE: out_tau[off] = tot[2]; break;
            case TRLO_JOINT_PLANAR: out_tau[off] = tot[3]; out_tau[off + 1] = tot[4]; out_tau[off + 2] = tot[0]; break;
            case TRLO_JOINT_PRISMATIC: out_tau[off] = tot[3]; break;
            case TRLO_JOINT_SPHERICAL: out_tau[off] = tot[0]; out_tau[off + 1] = tot[1]; out_tau[off + 2] = tot[2]; break;
            default: break;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Eigen 3.3 LDLT restated (Eigen/src/Cholesky/LDLT.h: ldlt_inplace<Lower>::unblocked + _solve_impl).
 * Lower triangle, diagonal pivoting, zero pivot => solution component 0 (Q1).
 * ---------------------------------------------------------------------------------------- */
int trlo_ldlt_solve(int n, const double* A, const double* b, double* x)
{
    double a[TRLO_MAX_DOF][TRLO_MAX_DOF];
    int tr[TRLO_MAX_DOF];
    double temp[TRLO_MAX_DOF], y[TRLO_MAX_DOF];
    int zero_pivots = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i][j] = A[i * n + j];

    if (n <= 1) {
        tr[0] = 0;
    } else {
        for (int k = 0; k < n; ++k) {
            int big = k;
            double bigv = fabs(a[k][k]);
            for (int i = k + 1; i < n; ++i) /* maxCoeff: first maximum wins */
                if (fabs(a[i][i]) > bigv) { bigv = fabs(a[i][i]); big = i; }
            tr[k] = big;
            if (k != big) {
                int s = n - big - 1;
                for (int q = 0; q < k; ++q) { double t = a[k][q]; a[k][q] = a[big][q]; a[big][q] = t; }
                for (int q = 0; q < s; ++q) { double t = a[big + 1 + q][k]; a[big + 1 + q][k] = a[big + 1 + q][big]; a[big + 1 + q][big] = t; }
                { double t = a[k][k]; a[k][k] = a[big][big]; a[big][big] = t; }
                for (int i = k + 1; i < big; ++i) { double t = a[i][k]; a[i][k] = a[big][i]; a[big][i] = t; }
            }
            int rs = n - k - 1;
            if (k > 0) {
                for (int q = 0; q < k; ++q) temp[q] = a[q][q] * a[k][q];
                double s = 0;
                for (int q = 0; q < k; ++q) { if (q == 0) s = a[k][0] * temp[0]; else s += a[k][q] * temp[q]; }
                a[k][k] -= s;
                for (int i = 0; i < rs; ++i) {
                    double t = 0;
                    for (int q = 0; q < k; ++q) { if (q == 0) t = a[k + 1 + i][0] * temp[0]; else t += a[k + 1 + i][q] * temp[q]; }
                    a[k + 1 + i][k] -= t;
                }
            }
            double akk = a[k][k];
            int pivot_valid = fabs(akk) > 0.0;
            if (k == 0 && !pivot_valid) {
                for (int j = 0; j < n; ++j) tr[j] = j;
                break;
            }
            if (rs > 0 && pivot_valid)
                for (int i = 0; i < rs; ++i) a[k + 1 + i][k] /= akk;
        }
    }
    /* solve: dst = P b */
    for (int i = 0; i < n; ++i) y[i] = b[i];
    for (int k = 0; k < n; ++k)
        if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    /* L^-1 (unit lower), column-oriented forward substitution */
    for (int k = 0; k < n; ++k)
        for (int i = k + 1; i < n; ++i) y[i] -= a[i][k] * y[k];
    /* pseudo-inverse of D with tolerance numeric_limits<double>::min() */
    for (int i = 0; i < n; ++i) {
        if (fabs(a[i][i]) > DBL_MIN) y[i] /= a[i][i];
        else { y[i] = 0; ++zero_pivots; }
    }
    /* L^-T */
    for (int k = n - 1; k >= 0; --k)
        for (int i = k + 1; i < n; ++i) y[k] -= a[i][k] * y[i];
    /* P^-1 */
    for (int k = n - 1; k >= 0; --k)
        if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < n; ++i) x[i] = y[i];
    return zero_pivots;
}

/* ------------------------------------------------------------------------------------------
 * PD targets as the controllers see them: cPDController::SetTargetTheta / SetTargetVel post-processing
 * (sim/PDController.cpp:191-211,430-461: spherical targets normalised, zero quaternion -> identity, vel slot 3 = 0)
 * followed by GetTargetTheta (:223-273), which converts WORLD-coordinate targets (pd_params UseWorldCoord != 0)
 * of revolute and spherical joints into joint coordinates from the current joint rotations. The joint's world
 * transform (cJoint::BuildWorldTrans, sim/Joint.cpp:126-131: child body transform x joint-in-body) equals
 * cKinTree::JointWorldTrans of the pose; CalcRotation (sim/Joint.cpp:70-93,565-569,624-628) is the joint's own
 * coordinate. Used by cImpPDController::BuildTargetPose / BuildTargetVel (sim/ImpPDController.cpp:198-243) and by
 * cPDController::CalcJointTau* (sim/PDController.cpp:302-428). Invalid controllers (root, null bodies) get zeros.
 * ---------------------------------------------------------------------------------------- */
static void pd_build_targets(const trlo_model* m, const double* pose, const double* tar_pose_in,
                             const double* tar_vel_in, double* tar_pose, double* tar_vel)
{
    int N = m->num_joints, n = m->num_dof;
    for (int i = 0; i < n; ++i) tar_pose[i] = tar_vel[i] = 0;
    for (int j = 0; j < N; ++j) {
        if (!trlo_pd_valid(m, j)) continue;
        int off = trlo_param_offset(m, j), size = trlo_param_size(m, j);
        int type = trlo_joint_type(m, j);
        int world = m->pd_params[(size_t)j * TRLO_PD_COLS + PD_USE_WORLD] != 0;
        for (int i = 0; i < size; ++i) {
            tar_pose[off + i] = tar_pose_in[off + i];
            tar_vel[off + i] = tar_vel_in[off + i];
        }
        if (type == TRLO_JOINT_SPHERICAL) {
            double* t = tar_pose + off;
            if (t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3] == 0) t[0] = 1;
            else normalize4(t);
            tar_vel[off + size - 1] = 0;
        }
        if (world && type == TRLO_JOINT_REVOLUTE) {      /* PDController.cpp:238-252 */
            double W[16];
            trlo_joint_world_trans(m, pose, j, W);
            mat4 Wm;
            memcpy(Wm.m, W, sizeof(W));
            double axis_world[3], theta_world;
            rot_mat_to_axis_angle(&Wm, axis_world, &theta_world);
            double axis_rel[3] = { 0, 0, 1 }, theta_rel = pose[off];
            /* CalcAxisWorld: trans * (0,0,1,0), Joint.cpp:46-53 */
            double axis_ref[3] = { Wm.m[0][2], Wm.m[1][2], Wm.m[2][2] };
            double dw = axis_world[0] * axis_ref[0] + axis_world[1] * axis_ref[1] + axis_world[2] * axis_ref[2] + 0.0;
            double dr = axis_rel[0] * axis_ref[0] + axis_rel[1] * axis_ref[1] + axis_rel[2] * axis_ref[2] + 0.0;
            double theta_offset = dw * theta_world - dr * theta_rel;
            tar_pose[off] -= theta_offset;
        }
        if (world && type == TRLO_JOINT_SPHERICAL) {     /* PDController.cpp:254-267 */
            double W[16];
            trlo_joint_world_trans(m, pose, j, W);
            mat4 Wm;
            memcpy(Wm.m, W, sizeof(W));
            double axis_rel[3], theta_rel;
            quat_to_axis_angle(pose_quat(pose + off), axis_rel, &theta_rel);
            quat world_rot = rot_mat_to_quat(&Wm);
            quat rel_rot = axis_angle_to_quat(axis_rel, theta_rel);
            quat tar_q = pose_quat(tar_pose + off);
            quat diff = quat_mul(quat_conj(world_rot), tar_q);
            tar_q = quat_mul(rel_rot, diff);
            tar_pose[off] = tar_q.w; tar_pose[off + 1] = tar_q.x; tar_pose[off + 2] = tar_q.y; tar_pose[off + 3] = tar_q.z;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * cExpPDController::UpdateControlForce -> cPDController::UpdateControlForce / CalcJointTau*
 * (sim/ExpPDController.cpp:59-66,148-159; sim/PDController.cpp:134-149,302-428): explicit PD per joint.
 * Joint coordinates / rates are the pose / vel entries (cJoint::CalcRotation, BuildPose, BuildVel,
 * CalcDisplacementPrismatic: the data contract of sim/Joint.cpp:353-417,565-628).
 * ---------------------------------------------------------------------------------------- */
void trlo_exp_pd_control_force(const trlo_model* m, double dt, const double* pose, const double* vel,
                               const double* tar_pose_in, const double* tar_vel_in,
                               const unsigned char* joint_active, const double* kp_joint, const double* kd_joint,
                               double* inout_tau)
{
    int N = m->num_joints;
    double tar_pose[TRLO_MAX_DOF], tar_vel[TRLO_MAX_DOF];
    if (!(dt > 0)) return; /* ExpPDController.cpp:62 */
    pd_build_targets(m, pose, tar_pose_in, tar_vel_in, tar_pose, tar_vel);
    for (int j = 0; j < N; ++j) {
        if (!trlo_pd_valid(m, j)) continue;
        if (joint_active && !joint_active[j]) continue; /* PDController.cpp:137 */
        int off = trlo_param_offset(m, j);
        double kp = kp_joint ? kp_joint[j] : m->pd_params[(size_t)j * TRLO_PD_COLS + PD_KP];
        double kd = kd_joint ? kd_joint[j] : m->pd_params[(size_t)j * TRLO_PD_COLS + PD_KD];
        switch (trlo_joint_type(m, j)) {
        case TRLO_JOINT_REVOLUTE:   /* :329-356 */
        case TRLO_JOINT_PRISMATIC: { /* :366-389 */
            double theta_err = tar_pose[off] - pose[off];
            double vel_err = tar_vel[off] - vel[off];
            inout_tau[off] += kp * theta_err + kd * vel_err;
            break;
        }
        case TRLO_JOINT_SPHERICAL: { /* :397-428 */
            quat tar_q = pose_quat(tar_pose + off);
            quat q = pose_quat(pose + off);
            quat q_err = quat_mul(quat_conj(q), tar_q);
            double axis[3], theta;
            quat_to_axis_angle(q_err, axis, &theta);
            double torque[4] = { (kp * theta) * axis[0], (kp * theta) * axis[1], (kp * theta) * axis[2], (kp * theta) * 0.0 };
            for (int i = 0; i < 3; ++i) torque[i] += kd * (tar_vel[off + i] - vel[off + i]);
            for (int i = 0; i < 4; ++i) inout_tau[off + i] += torque[i];
            break;
        }
        default: /* planar (unsupported: zero, :358-364), fixed (:391-395) */
            break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * cImpPDController::CalcControlForces, sim/ImpPDController.cpp:137-196
 * ---------------------------------------------------------------------------------------- */
int trlo_imp_pd_control_force(const trlo_model* m, double dt, const double* pose, const double* vel,
                              const double* tar_pose_in, const double* tar_vel_in,
                              const unsigned char* joint_active, const double* kp_joint, const double* kd_joint,
                              double* inout_tau, double* out_mass, double* out_bias)
{
    int N = m->num_joints, n = m->num_dof;
    double M[TRLO_MAX_DOF * TRLO_MAX_DOF];
    double C[TRLO_MAX_DOF], mKp[TRLO_MAX_DOF], mKd[TRLO_MAX_DOF], Kp[TRLO_MAX_DOF], Kd[TRLO_MAX_DOF];
    double tar_pose[TRLO_MAX_DOF], tar_vel[TRLO_MAX_DOF];
    double pose_inc[TRLO_MAX_DOF], pose_err[TRLO_MAX_DOF], vel_err[TRLO_MAX_DOF], acc[TRLO_MAX_DOF];
    if (!(dt > 0)) return 0; /* UpdateControlForce :56 */

    trlo_rbd_update(m, pose, vel, M, C); /* UpdateRBDModel :58-61 (model owned or shared: same numbers) */
    if (out_mass) memcpy(out_mass, M, sizeof(double) * (size_t)(n * n));
    if (out_bias) memcpy(out_bias, C, sizeof(double) * (size_t)n);

    /* InitGains :100-121, BuildTargetPose/Vel :198-243: zero for invalid controllers (root) */
    for (int i = 0; i < n; ++i) { mKp[i] = mKd[i] = 0; }
    for (int j = 0; j < N; ++j) {
        if (!trlo_pd_valid(m, j)) continue;
        int off = trlo_param_offset(m, j), size = trlo_param_size(m, j);
        double kp = kp_joint ? kp_joint[j] : m->pd_params[(size_t)j * TRLO_PD_COLS + PD_KP];
        double kd = kd_joint ? kd_joint[j] : m->pd_params[(size_t)j * TRLO_PD_COLS + PD_KD];
        for (int i = 0; i < size; ++i) { mKp[off + i] = 1.0 * kp; mKd[off + i] = 1.0 * kd; }
    }
    pd_build_targets(m, pose, tar_pose_in, tar_vel_in, tar_pose, tar_vel);
    /* mask gains of invalid / inactive controllers :151-161 */
    for (int i = 0; i < n; ++i) { Kp[i] = mKp[i]; Kd[i] = mKd[i]; }
    for (int j = 0; j < N; ++j) {
        int active = joint_active ? (joint_active[j] != 0) : 1;
        if (!trlo_pd_valid(m, j) || !active) {
            int off = trlo_param_offset(m, j), size = trlo_param_size(m, j);
            for (int i = 0; i < size; ++i) Kp[off + i] = Kd[off + i] = 0;
        }
    }
    /* M.diagonal() += t * mKd  -- the UNMASKED gains (Q2) :163-165 */
    for (int i = 0; i < n; ++i) M[i * n + i] += dt * mKd[i];

    trlo_vel_to_pose_diff(m, pose, vel, pose_inc);        /* :167-169 */
    for (int i = 0; i < n; ++i) pose_inc[i] = pose[i] + dt * pose_inc[i]; /* :171 */
    trlo_post_process_pose(m, pose_inc);                  /* :172 */
    trlo_calc_vel(m, pose_inc, tar_pose, 1, pose_err);    /* :175 (Q10) */
    for (int i = 0; i < n; ++i) vel_err[i] = tar_vel[i] - vel[i];
    for (int i = 0; i < n; ++i) acc[i] = (Kp[i] * pose_err[i] + Kd[i] * vel_err[i]) - C[i]; /* :177 */
    int zp = trlo_ldlt_solve(n, M, acc, acc);             /* :189 */
    for (int i = 0; i < n; ++i)
        inout_tau[i] += Kp[i] * pose_err[i] + Kd[i] * (vel_err[i] - dt * acc[i]); /* :195 */
    return zp;
}

/* ------------------------------------------------------------------------------------------
 * cMotion / cKinCharacter (anim/Motion.cpp, anim/KinCharacter.cpp)
 * ---------------------------------------------------------------------------------------- */
static const double* frame_row(const trlo_model* m, int f) { return m->frames + (size_t)f * (size_t)(m->num_dof + 1); }
static double motion_duration(const trlo_model* m) { return frame_row(m, m->num_frames - 1)[0]; } /* Motion.cpp:219-224 */

void trlo_motion_index_phase(const trlo_model* m, double time, int* out_idx, double* out_phase)
{ /* Motion.cpp:240-273 */
    double max_time = motion_duration(m);
    if (!m->loop) {
        if (time <= 0) { *out_idx = 0; *out_phase = 0; return; }
        else if (time >= max_time) { *out_idx = m->num_frames - 2; *out_phase = 1; return; }
    }
    time = fmod(time, max_time);
    if (time < 0) time += max_time;
    /* std::upper_bound over the frame start times */
    int lo = 0, hi = m->num_frames;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (!(time < frame_row(m, mid)[0])) lo = mid + 1; else hi = mid;
    }
    int idx = lo - 1;
    double t0 = frame_row(m, idx)[0];
    double t1 = frame_row(m, idx + 1)[0];
    *out_idx = idx;
    *out_phase = (time - t0) / (t1 - t0);
}

void trlo_motion_calc_frame(const trlo_model* m, double time, double* out_pose)
{ /* Motion.cpp:112-139 with the cKinCharacter blend func (KinCharacter.cpp:355-360) */
    int idx; double phase;
    trlo_motion_index_phase(m, time, &idx, &phase);
    double lerp = clampd(phase, 0.0, 1.0); /* Saturate */
    trlo_lerp_poses(m, frame_row(m, idx) + 1, frame_row(m, idx + 1) + 1, lerp, out_pose);
}

static int motion_cycle_count(const trlo_model* m, double time) /* Motion.cpp:231-238 */
{
    double dur = motion_duration(m);
    double phases = time / dur;
    int count = (int)floor(phases);
    if (!m->loop) count = (count < 0) ? 0 : (count > 1 ? 1 : count);
    return count;
}

static void kin_calc_pose(const trlo_model* m, double time, const double origin[3], quat origin_rot, double* out_pose)
{ /* KinCharacter.cpp:280-315 */
    double root_delta[3] = { 0, 0, 0 };
    trlo_motion_calc_frame(m, time, out_pose);
    if (m->loop) {
        int cycle = motion_cycle_count(m, time);
        const double* fb = frame_row(m, 0) + 1;
        const double* fe = frame_row(m, m->num_frames - 1) + 1;
        for (int i = 0; i < 3; ++i) root_delta[i] = cycle * (fe[i] - fb[i]); /* mCycleRootDelta :264-277 */
    }
    quat ident = { 1, 0, 0, 0 };
    quat rdr = quat_mul(origin_rot, ident);
    quat rr = quat_mul(rdr, pose_quat(out_pose + 3));
    double rp[3];
    quat_rot_vec(rdr, out_pose, rp);
    for (int i = 0; i < 3; ++i) {
        root_delta[i] += origin[i];
        rp[i] += root_delta[i];
        out_pose[i] = rp[i];
    }
    out_pose[3] = rr.w; out_pose[4] = rr.x; out_pose[5] = rr.y; out_pose[6] = rr.z;
}

void trlo_kin_char_pose(const trlo_model* m, double time, const double origin_pos[3], const double origin_rot[4],
                        double* out_pose, double* out_vel)
{ /* cKinCharacter::Pose, KinCharacter.cpp:111-123; gDiffTimeStep = 1/60.0 (:5) */
    const double diff_dt = 1 / 60.0;
    double zero3[3] = { 0, 0, 0 };
    quat orot = { 1, 0, 0, 0 };
    double pose1[TRLO_MAX_DOF];
    if (origin_rot) { orot.w = origin_rot[0]; orot.x = origin_rot[1]; orot.y = origin_rot[2]; orot.z = origin_rot[3]; }
    const double* op = origin_pos ? origin_pos : zero3;
    kin_calc_pose(m, time, op, orot, out_pose);
    if (out_vel) {
        kin_calc_pose(m, time + diff_dt, op, orot, pose1);
        trlo_calc_vel(m, out_pose, pose1, diff_dt, out_vel);
    }
}

/* ------------------------------------------------------------------------------------------
 * ground sampling (sim/GroundVar2D.cpp:109-124,621-700; sim/GroundVar3D.cpp:154-170,564-684)
 * ---------------------------------------------------------------------------------------- */
static double neg_inf(void) { return -INFINITY; }

static double sample_segment_2d(const trlo_ground* g, int s, const double pos[3], int* valid, int cell[3])
{
    const double tol = 0.0001;
    int w = g->res[s][0], l = g->res[s][1];
    /* CalcGridCoord, GroundVar2D.cpp:652-681 */
    double c0 = pos[0] - g->origin[s][0];
    double c1 = pos[2] - g->origin[s][2];
    c0 /= g->scale[s][0];
    c1 /= g->scale[s][2];
    c0 += ((w - 1) * 0.5);
    c1 += ((l - 1) * 0.5);
    if (c0 > -tol && c0 < w - 1 + tol && c1 > -tol && c1 < l - 1 + tol) {
        c0 = clampd(c0, 0.0, w - 1.0);
        c1 = clampd(c1, 0.0, l - 1.0);
    }
    /* SampleHeight, :621-638 */
    int outside = (c0 < 0 || c0 > w - 1 || c1 < 0 || c1 > l - 1);
    *valid = !outside;
    c0 = clampd(c0, 0.0, w - 1.0);
    c1 = clampd(c1, 0.0, l - 1.0);
    int i = (int)c0;
    int j = (w - 1 < i + 1) ? (w - 1) : (i + 1);
    double lerp = c0 - i;
    double a = (double)g->data[s][i] * g->scale[s][1]; /* GetVertex(i,0)[1], :594-611 (row 0, Q8) */
    double b = (double)g->data[s][j] * g->scale[s][1];
    if (cell) { cell[0] = s; cell[1] = i; cell[2] = j; }
    return (1 - lerp) * a + lerp * b;
}

static double sample_slab_3d(const trlo_ground* g, int s, const double pos[3], int* valid, int cell[3])
{
    const double tol = 0.0001;
    int rx = g->res[s][0], rz = g->res[s][1];
    double c0 = pos[0] - g->origin[s][0];
    double c1 = pos[2] - g->origin[s][2];
    c0 /= g->scale[s][0];
    c1 /= g->scale[s][2];
    c0 += ((rx - 1) * 0.5);
    c1 += ((rz - 1) * 0.5);
    c0 = clampd(c0, 0.0, rx - 1.0 - tol);
    c1 = clampd(c1, 0.0, rz - 1.0 - tol);
    int i = (int)c0, j = (int)c1;
    double lx = c0 - i, lz = c1 - j;
    int lower = (lz <= lx);
    int i0 = lower ? i : i + 1, j0 = lower ? j : j + 1;
    int i1 = lower ? i + 1 : i, j1 = lower ? j : j + 1;
    int i2 = lower ? i + 1 : i, j2 = lower ? j + 1 : j;
    double v0 = (double)g->data[s][j0 * rx + i0] * g->scale[s][1];
    double v1 = (double)g->data[s][j1 * rx + i1] * g->scale[s][1];
    double v2 = (double)g->data[s][j2 * rx + i2] * g->scale[s][1];
    double p[2] = { lx, lz };
    double ca[2], cb[2], cc[2];
    if (lower) { ca[0] = 0; ca[1] = 0; cb[0] = 1; cb[1] = 0; cc[0] = 1; cc[1] = 1; }
    else { ca[0] = 1; ca[1] = 1; cb[0] = 0; cb[1] = 1; cc[0] = 0; cc[1] = 0; }
    double bc[3];
    calc_barycentric(p, ca, cb, cc, bc);
    *valid = 1;
    if (cell) { cell[0] = s; cell[1] = i; cell[2] = j; }
    return bc[0] * v0 + bc[1] * v1 + bc[2] * v2;
}

double trlo_sample_height(const trlo_ground* g, const double pos[3], int* out_valid, int out_cell[3])
{
    int valid = 0;
    double h = neg_inf();
    if (out_cell) { out_cell[0] = -1; out_cell[1] = 0; out_cell[2] = 0; }
    if (!g || g->kind == TRLO_GROUND_NONE || g->kind == TRLO_GROUND_PLANE) { /* cGround::SampleHeight, Ground.cpp:176-186 */
        if (out_valid) *out_valid = 1;
        return 0;
    }
    for (int s = 0; s < g->num_pieces; ++s) {
        if (g->kind == TRLO_GROUND_VAR2D) {
            if (pos[0] >= g->bounds[s][0] && pos[0] <= g->bounds[s][1]) { /* ContainsPos :697-700 */
                h = sample_segment_2d(g, s, pos, &valid, out_cell);
                break;
            }
        } else {
            if (pos[0] >= g->bounds[s][0] && pos[0] <= g->bounds[s][1] &&
                pos[2] >= g->bounds[s][2] && pos[2] <= g->bounds[s][3]) { /* tSlab::ContainsPos :596-600 */
                h = sample_slab_3d(g, s, pos, &valid, out_cell);
                break;
            }
        }
    }
    if (out_valid) *out_valid = valid;
    return h;
}

/* batch form of cGround::SampleHeight(MatrixXd, VectorXd&) with the scalar overload's numerics (test helper) */
void trlo_sample_height_batch(const trlo_ground* g, int count, const double* pos, double* out_h, int* out_valid, int* out_cell)
{
    for (int i = 0; i < count; ++i) {
        int v, c[3];
        out_h[i] = trlo_sample_height(g, pos + 3 * (size_t)i, &v, c);
        if (out_valid) out_valid[i] = v;
        if (out_cell) { out_cell[3 * i] = c[0]; out_cell[3 * i + 1] = c[1]; out_cell[3 * i + 2] = c[2]; }
    }
}

/* ------------------------------------------------------------------------------------------
 * observation (sim/TerrainRLCharController.cpp:261-436, sim/CtController.cpp:189-228,482-632,
 *              sim/CtPhaseController.cpp:120-139, sim/BipedController3D.cpp:1927-1965,
 *              sim/MonopedHopperController.cpp:1533-1552)
 * ---------------------------------------------------------------------------------------- */
static int has_ground(const trlo_ground* g) { return g && g->kind != TRLO_GROUND_NONE; }

static mat4 ground_sample_trans(const trlo_ground* g, const double* pose) /* TerrainRLCharController.cpp:261-279 */
{
    double ground_h = 0;
    if (has_ground(g)) {
        int v;
        ground_h = trlo_sample_height(g, pose, &v, 0);
        if (!isfinite(ground_h)) ground_h = 0; /* Q11 */
    }
    mat4 t = build_origin_trans(pose);
    t = inv_rigid_mat(&t);
    t.m[1][3] += ground_h;
    return t;
}

void trlo_ground_sample_trans(const trlo_ground* g, const double* pose, double out[16])
{
    mat4 t = ground_sample_trans(g, pose);
    memcpy(out, t.m, sizeof(t.m));
}

static void ground_sample_pos(const trlo_obs_cfg* cfg, const mat4* trans, int s, int flip, double out[3])
{
    vec4 p;
    if (cfg->sampler == TRLO_SAMPLER_LINE) { /* TerrainRLCharController.cpp:310-319 */
        double dist = ((cfg->view_max - cfg->view_min) * s) / (cfg->num_samples - 1) + cfg->view_min;
        p.v[0] = dist; p.v[1] = 0; p.v[2] = 0; p.v[3] = 1;
    } else { /* CtController.cpp:189-228 / BipedController3D.cpp:1927-1965 */
        int res = cfg->grid_res;
        double ox = 0.5 * (cfg->view_max + cfg->view_min);
        double w = cfg->view_max - cfg->view_min;
        double u = (double)(s % res) / (res - 1);
        double v = (double)(s / res) / (res - 1);
        double x = (u - 0.5) * w;
        double z = (v - 0.5) * w;
        if (flip && (cfg->obs_kind == TRLO_OBS_CT || cfg->obs_kind == TRLO_OBS_CT_3D)) z = -z; /* cCtController::FlipStance */
        p.v[0] = ox + x; p.v[1] = 0 + 0; p.v[2] = 0 + z; p.v[3] = 1;
    }
    vec4 r = mat4_mulv(trans, &p);
    out[0] = r.v[0]; out[1] = r.v[1]; out[2] = r.v[2];
}

void trlo_parse_ground(const trlo_ground* g, const trlo_obs_cfg* cfg, const double* pose, int flip,
                       double* out_samples, int* out_cells)
{ /* ParseGround + SampleGround, TerrainRLCharController.cpp:281-308 */
    mat4 trans = ground_sample_trans(g, pose);
    for (int s = 0; s < cfg->num_samples; ++s) {
        if (!has_ground(g)) { out_samples[s] = 0; continue; }
        double p[3];
        int v;
        ground_sample_pos(cfg, &trans, s, flip, p);
        double h = trlo_sample_height(g, p, &v, out_cells ? out_cells + 3 * s : 0);
        h -= trans.m[1][3];
        out_samples[s] = h;
    }
}

static int pose_feature_size(const trlo_model* m, const trlo_obs_cfg* cfg)
{
    int N = m->num_joints;
    switch (cfg->obs_kind) {
    case TRLO_OBS_CT: return N * 2 + 1;          /* CtController.cpp:482-487 */
    case TRLO_OBS_CT_3D: return N * (3 + 4) + 1;
    default: return N * 2 - 1;                   /* TerrainRLCharController.cpp:485-488 */
    }
}

static int vel_feature_size(const trlo_model* m, const trlo_obs_cfg* cfg)
{
    int N = m->num_joints;
    switch (cfg->obs_kind) {
    case TRLO_OBS_CT: return N * 2;              /* CtController.cpp:489-504 */
    case TRLO_OBS_CT_3D: return N * (3 + 4 - 1);
    default: return N * 2;                       /* TerrainRLCharController.cpp:490-493 */
    }
}

int trlo_poli_state_size(const trlo_model* m, const trlo_obs_cfg* cfg)
{
    int F = cfg->num_samples + pose_feature_size(m, cfg) + vel_feature_size(m, cfg);
    if (cfg->num_phase_bins >= 0) F += 1 + cfg->num_phase_bins; /* CtPhaseController.cpp:66-82 */
    return F;
}

void trlo_build_poli_state(const trlo_model* m, const trlo_ground* g, const trlo_obs_cfg* cfg,
                           const double* pose, const double* vel, const double* body_pos, const double* body_vel,
                           const double* body_quat, const double* body_angvel, const double* ground_vel,
                           double phase, int flip, double* out_state);

static void build_poli_state_impl(const trlo_model* m, const trlo_ground* g, const trlo_obs_cfg* cfg,
                                  const double* pose, const double* vel, const double* body_pos, const double* body_vel,
                                  const double* body_quat, const double* body_angvel, const double* ground_vel,
                                  double phase, int flip, double* out_state)
{
    int N = m->num_joints;
    int G = cfg->num_samples;
    int psize = pose_feature_size(m, cfg), vsize = vel_feature_size(m, cfg);
    double* out_pose = out_state + G;
    double* out_vel = out_state + G + psize;
    int is_ct = (cfg->obs_kind == TRLO_OBS_CT || cfg->obs_kind == TRLO_OBS_CT_3D);
    int is3d = (cfg->obs_kind == TRLO_OBS_CT_3D);
    int pos_dim = is3d ? 3 : 2;

    /* ground block */
    trlo_parse_ground(g, cfg, pose, flip, out_state, 0);

    /* pose block: BuildPoliStatePose (TerrainRLCharController.cpp:370-406 / CtController.cpp:507-576) */
    mat4 origin = build_origin_trans(pose);
    quat origin_quat = rot_mat_to_quat(&origin);
    if (is_ct && flip)
        for (int c = 0; c < 4; ++c) origin.m[2][c] *= -1;
    double ground_h = 0;
    if (has_ground(g)) { int v; ground_h = trlo_sample_height(g, pose, &v, 0); }
    vec4 rp = { { pose[0], pose[1] - ground_h, pose[2], 1 } };
    vec4 root_rel = mat4_mulv(&origin, &rp);
    root_rel.v[3] = 0;
    for (int i = 0; i < psize; ++i) out_pose[i] = 0;
    out_pose[0] = root_rel.v[1];
    int idx = 1;
    for (int i = (is_ct ? 0 : 1); i < N; ++i) {
        if (!trlo_valid_body(m, i)) continue;
        vec4 cp = { { body_pos[3 * i], body_pos[3 * i + 1] - ground_h, body_pos[3 * i + 2], 1 } };
        vec4 r = mat4_mulv(&origin, &cp);
        r.v[3] = 0;
        for (int c = 0; c < pos_dim; ++c) out_pose[idx + c] = r.v[c] - root_rel.v[c];
        idx += pos_dim;
        if (is3d) { /* ENABLE_MAX_STATE, CtController.cpp:553-573 */
            quat q = { body_quat[4 * i], body_quat[4 * i + 1], body_quat[4 * i + 2], body_quat[4 * i + 3] };
            q = quat_mul(origin_quat, q);
            if (flip) { q.x = -q.x; q.y = -q.y; } /* MirrorQuaternion(eAxisZ), MathUtil.cpp:491-499 */
            if (q.w < 0) { q.w *= -1; q.x *= -1; q.y *= -1; q.z *= -1; }
            out_pose[idx] = q.w; out_pose[idx + 1] = q.x; out_pose[idx + 2] = q.y; out_pose[idx + 3] = q.z;
            idx += 4;
        }
    }

    /* vel block: BuildPoliStateVel (TerrainRLCharController.cpp:408-436 / CtController.cpp:578-632) */
    mat4 origin_v = build_origin_trans(pose);
    if (is_ct && flip)
        for (int c = 0; c < 4; ++c) origin_v.m[2][c] *= -1;
    double gv[3] = { 0, 0, 0 };
    if (ground_vel) { gv[0] = ground_vel[0]; gv[1] = ground_vel[1]; gv[2] = ground_vel[2]; }
    idx = 0;
    for (int i = 0; i < N; ++i) {
        vec4 cv = { { body_vel[3 * i] - gv[0], body_vel[3 * i + 1] - gv[1], body_vel[3 * i + 2] - gv[2], 0 } };
        vec4 r = mat4_mulv(&origin_v, &cv);
        for (int c = 0; c < pos_dim; ++c) out_vel[idx + c] = r.v[c];
        idx += pos_dim;
        if (is3d) {
            vec4 av = { { body_angvel[3 * i], body_angvel[3 * i + 1], body_angvel[3 * i + 2], 0 } };
            vec4 ra = mat4_mulv(&origin_v, &av);
            if (flip) for (int c = 0; c < 4; ++c) ra.v[c] = -ra.v[c];
            for (int c = 0; c < 3; ++c) out_vel[idx + c] = ra.v[c];
            idx += 3;
        }
    }

    if (cfg->obs_kind == TRLO_OBS_TERRAIN_RL_HOPPER) { /* MonopedHopperController.cpp:1533-1552 */
        mat4 R = rotate_mat_quat(pose_quat(pose + 3));
        double axis[3], theta;
        rot_mat_to_axis_angle(&R, axis, &theta);
        if (axis[2] < 0) theta = -theta; /* ref_axis (0,0,1) . root_axis < 0 */
        out_pose[psize - 1] = theta;
        out_vel[vsize - 1] = vel[5];
    }

    if (!is_ct && flip) { /* FlipPoliPoseStance, BipedController3D.cpp:1811-1830: swap the two leg blocks at the tail */
        int legs = (N - 1) / 2; /* joints per leg: (#joints - root) / 2 for the biped */
        int num_leg_params = legs * 2;
        for (int i = 0; i < num_leg_params; ++i) {
            int a = psize - i - 1, b = psize - num_leg_params - i - 1;
            double t = out_pose[a]; out_pose[a] = out_pose[b]; out_pose[b] = t;
            a = vsize - i - 1; b = vsize - num_leg_params - i - 1;
            t = out_vel[a]; out_vel[a] = out_vel[b]; out_vel[b] = t;
        }
    }

    if (cfg->num_phase_bins >= 0) { /* BuildPhaseState, CtPhaseController.cpp:120-139 (bins on, cos off) */
        double* ph = out_state + G + psize + vsize;
        for (int i = 0; i < 1 + cfg->num_phase_bins; ++i) ph[i] = 0;
        ph[0] = phase;
        if (cfg->num_phase_bins > 0) {
            int bin = (int)(phase * cfg->num_phase_bins);
            if (bin >= 0 && bin < cfg->num_phase_bins) ph[bin + 1] = 1;
        }
    }
}

void trlo_build_poli_state(const trlo_model* m, const trlo_ground* g, const trlo_obs_cfg* cfg,
                           const double* pose, const double* vel, const double* body_pos, const double* body_vel,
                           const double* body_quat, const double* body_angvel, const double* ground_vel,
                           double phase, int flip, double* out_state)
{
    build_poli_state_impl(m, g, cfg, pose, vel, body_pos, body_vel, body_quat, body_angvel, ground_vel, phase, flip, out_state);
}

/* ------------------------------------------------------------------------------------------
 * imitation reward: cScenarioExpImitate::CalcReward (scenarios/ScenarioExpImitate.cpp:13-159) with
 * cKinTree::NormalizePoseHeading / CalcPoseErr / CalcVelErr (anim/KinTree.cpp:1319-1426,1647-1672),
 * cRBDUtil::CalcCoM / CalcCoMVel (pose-only "hack" overload, sim/RBDUtil.cpp:499-505,557-598),
 * cRBDUtil::BuildEndEffectorJacobian (:209-233), cSimCharacter::CalcCOMVel (sim/SimCharacter.cpp:359-377)
 * ---------------------------------------------------------------------------------------- */
static double quat_theta(quat dq) /* cMathUtil::QuatTheta, util/MathUtil.cpp:468-477 */
{
    if (fabs(dq.w) > 1) dq = quat_normalize(dq);
    return 2 * acos(dq.w);
}

static sptrans build_trans_r(const double r[3]) /* cSpAlg::BuildTrans(r): identity rotation */
{
    sptrans X = identity_trans();
    X.r[0] = r[0]; X.r[1] = r[1]; X.r[2] = r[2];
    return X;
}

/* cRBDUtil::CalcWorldVel(joint_mat, pose, vel, joint_id) = BuildEndEffectorJacobian(...) * vel (RBDUtil.cpp:209-233,474-480) */
static spvec calc_world_vel(const trlo_model* m, const double* pose, const double* vel, int joint_id)
{
    int n = m->num_dof;
    static const spvec zero_sv;
    spvec J[TRLO_MAX_DOF];
    for (int i = 0; i < n; ++i) J[i] = zero_sv;
    rbd_cache c; /* only the subspace columns are used */
    for (int j = 0; j < m->num_joints; ++j) build_joint_subspace(m, pose, j, &c);
    int cur = joint_id;
    sptrans cur_trans = identity_trans();
    while (cur != -1) {
        int off = trlo_param_offset(m, cur), size = trlo_param_size(m, cur);
        for (int k = 0; k < size; ++k) {
            spvec col;
            for (int r = 0; r < 6; ++r) col.v[r] = c.S[r][off + k];
            J[off + k] = apply_trans_m(&cur_trans, &col);
        }
        /* BuildParentChildTransform = InvTrans(MatToTrans(ChildParentTrans)) (RBDUtil.cpp:712-724) */
        mat4 cp = child_parent_trans(m, pose, cur);
        sptrans cpt = mat_to_trans(&cp);
        sptrans pc = inv_trans(&cpt);
        cur_trans = comp_trans(&cur_trans, &pc);
        cur = trlo_parent(m, cur);
    }
    spvec sv = zero_sv;
    for (int i = 0; i < n; ++i) {
        spvec col = apply_inv_trans_m(&cur_trans, &J[i]);
        for (int r = 0; r < 6; ++r) {
            if (i == 0) sv.v[r] = col.v[r] * vel[0];
            else sv.v[r] += col.v[r] * vel[i];
        }
    }
    return sv;
}

/* cRBDUtil::CalcCoM(joint_mat, body_defs, pose, vel, com, com_vel), RBDUtil.cpp:557-598 */
static void calc_com(const trlo_model* m, const double* pose, const double* vel, double com[3], double com_vel[3])
{
    double total = 0;
    for (int i = 0; i < 3; ++i) { com[i] = 0; com_vel[i] = 0; }
    for (int j = 0; j < m->num_joints; ++j) {
        if (!trlo_valid_body(m, j)) continue;
        mat4 jw = joint_world_trans(m, pose, j);
        mat4 bj = body_joint_trans(m, j);
        mat4 bw = mat4_mul(&jw, &bj);
        vec4 lc = { { 0, 0, 0, 1 } };
        vec4 wc = mat4_mulv(&bw, &lc);
        double wcom[3] = { wc.v[0], wc.v[1], wc.v[2] };
        sptrans ct = build_trans_r(wcom);
        spvec sv = calc_world_vel(m, pose, vel, j);
        sv = apply_trans_m(&ct, &sv);
        double mass = brow(m, j)[BD_MASS];
        for (int i = 0; i < 3; ++i) { com[i] += mass * wcom[i]; com_vel[i] += mass * sv.v[3 + i]; }
        total += mass;
    }
    for (int i = 0; i < 3; ++i) { com[i] /= total; com_vel[i] /= total; }
}

void trlo_calc_com(const trlo_model* m, const double* pose, const double* vel, double com[3], double com_vel[3])
{
    calc_com(m, pose, vel, com, com_vel);
}

/* cKinTree::NormalizePoseHeading(joint_mat, pose, vel), KinTree.cpp:1647-1672 */
static void normalize_pose_heading(double* pose, double* vel)
{
    double heading = trlo_calc_heading(pose);
    const double yaxis[3] = { 0, 1, 0 };
    quat hq = axis_angle_to_quat(yaxis, -heading);
    quat rr = quat_mul(hq, pose_quat(pose + 3));
    pose[0] = 0; pose[2] = 0;
    pose[3] = rr.w; pose[4] = rr.x; pose[5] = rr.y; pose[6] = rr.z;
    double rv[3], ra[3];
    quat_rot_vec(hq, vel, rv);
    quat_rot_vec(hq, vel + 3, ra);
    for (int i = 0; i < 3; ++i) vel[i] = rv[i];
    vel[3] = ra[0]; vel[4] = ra[1]; vel[5] = ra[2]; vel[6] = 0; /* SetRootAngVel writes the 4-segment of a w=0 vector */
}

double trlo_imitate_reward(const trlo_model* m, const trlo_ground* g, const double* pose0_in, const double* vel0_in,
                           const double* pose1_in, const double* vel1_in, const double* body_vel, int fallen,
                           double* out_terms /* [5] pose, vel, end_eff, root, com errors or NULL */)
{
    int N = m->num_joints, n = m->num_dof;
    double pose_w = 0.5, vel_w = 0.05, end_eff_w = 0.15, root_w = 0.1, com_w = 0.2;
    double total_w = pose_w + vel_w + end_eff_w + root_w + com_w;
    pose_w /= total_w; vel_w /= total_w; end_eff_w /= total_w; root_w /= total_w; com_w /= total_w;
    const double pose_scale = 2, vel_scale = 0.005, end_eff_scale = 40, root_scale = 10, com_scale = 10, err_scale = 1;
    if (out_terms) for (int i = 0; i < 5; ++i) out_terms[i] = 0;
    if (fallen) return 0;

    double pose0[TRLO_MAX_DOF], vel0[TRLO_MAX_DOF], pose1[TRLO_MAX_DOF], vel1[TRLO_MAX_DOF];
    memcpy(pose0, pose0_in, sizeof(double) * (size_t)n); memcpy(vel0, vel0_in, sizeof(double) * (size_t)n);
    memcpy(pose1, pose1_in, sizeof(double) * (size_t)n); memcpy(vel1, vel1_in, sizeof(double) * (size_t)n);
    mat4 origin_trans = build_origin_trans(pose0_in);

    /* cSimCharacter::CalcCOMVel: mass-weighted body velocities */
    double cv0w[3] = { 0, 0, 0 }, tm = 0;
    for (int j = 0; j < N; ++j) {
        if (!trlo_valid_body(m, j)) continue;
        double mass = brow(m, j)[BD_MASS];
        for (int i = 0; i < 3; ++i) cv0w[i] += mass * body_vel[3 * j + i];
        tm += mass;
    }
    for (int i = 0; i < 3; ++i) cv0w[i] /= tm;
    double tmp_com[3], cv1w[3];
    calc_com(m, pose1, vel1, tmp_com, cv1w); /* cRBDUtil::CalcCoMVel(joint_mat, body_defs, pose1, vel1) */

    normalize_pose_heading(pose0, vel0);
    normalize_pose_heading(pose1, vel1);
    double com0[3], com1[3], cvdummy[3];
    calc_com(m, pose0, vel0, com0, cvdummy);
    calc_com(m, pose1, vel1, com1, cvdummy);

    /* joint weights: DiffWeight normalised by their 1-norm (ScenarioExpImitate.cpp:232-244) */
    double wsum = 0;
    for (int j = 0; j < N; ++j) wsum += fabs(jrow(m, j)[17]);
    double pose_err = 0, vel_err = 0, end_eff_err = 0, root_err = 0, com_err = 0;
    int num_end_effs = 0;
    {
        double w = jrow(m, 0)[17] / wsum;
        quat d = quat_mul(pose_quat(pose1 + 3), quat_conj(pose_quat(pose0 + 3))); /* QuatDiff(q0,q1) = q1 * conj(q0) */
        double th = quat_theta(d);
        pose_err += w * (th * th);
        double s = 0;
        for (int i = 0; i < 4; ++i) { double dv = vel1[3 + i] - vel0[3 + i]; s += dv * dv; }
        vel_err += w * s;
    }
    for (int j = 1; j < N; ++j) {
        double w = jrow(m, j)[17] / wsum;
        int off = trlo_param_offset(m, j), size = trlo_param_size(m, j);
        double cpe = 0, cve = 0;
        if (trlo_joint_type(m, j) == TRLO_JOINT_SPHERICAL) {
            quat d = quat_mul(pose_quat(pose1 + off), quat_conj(pose_quat(pose0 + off)));
            double th = quat_theta(d);
            cpe = th * th;
        } else {
            for (int i = 0; i < size; ++i) { double dv = pose1[off + i] - pose0[off + i]; cpe += dv * dv; }
        }
        for (int i = 0; i < size; ++i) { double dv = vel1[off + i] - vel0[off + i]; cve += dv * dv; }
        pose_err += w * cpe;
        vel_err += w * cve;
        if (jrow(m, j)[16] != 0) { /* IsEndEffector */
            mat4 jw0 = joint_world_trans(m, pose0_in, j); /* mChar->CalcJointPos(j): world position in the sim pose */
            double p0w[3] = { jw0.m[0][3], jw0.m[1][3], jw0.m[2][3] };
            int v;
            double gh0 = (g && g->kind != TRLO_GROUND_NONE) ? trlo_sample_height(g, p0w, &v, 0) : 0.0;
            vec4 p0 = { { p0w[0], p0w[1], p0w[2], 1 } };
            p0 = mat4_mulv(&origin_trans, &p0);
            mat4 jw1 = joint_world_trans(m, pose1, j);
            double p1[3] = { jw1.m[0][3], jw1.m[1][3], jw1.m[2][3] };
            double r0[3] = { p0.v[0] - com0[0], p0.v[1] - gh0, p0.v[2] - com0[2] };
            double r1[3] = { p1[0] - com1[0], p1[1] - 0.0, p1[2] - com1[2] };
            double e2 = 0;
            for (int i = 0; i < 3; ++i) { double dv = r1[i] - r0[i]; e2 += dv * dv; }
            end_eff_err += e2;
            ++num_end_effs;
        }
    }
    if (num_end_effs > 0) end_eff_err /= num_end_effs;
    {
        int v;
        double rgh0 = (g && g->kind != TRLO_GROUND_NONE) ? trlo_sample_height(g, pose0_in, &v, 0) : 0.0;
        double h0 = pose0[1] - rgh0, h1 = pose1[1] - 0.0;
        double dv2 = 0;
        for (int i = 0; i < 3; ++i) { double dv = vel1[i] - vel0[i]; dv2 += dv * dv; }
        root_err = (h1 - h0) * (h1 - h0) + 0 * 0.01 * dv2;
    }
    {
        double s = 0;
        for (int i = 0; i < 3; ++i) { double dv = cv1w[i] - cv0w[i]; s += dv * dv; }
        com_err = 0.1 * s;
    }
    if (out_terms) { out_terms[0] = pose_err; out_terms[1] = vel_err; out_terms[2] = end_eff_err; out_terms[3] = root_err; out_terms[4] = com_err; }
    double pose_reward = exp(-err_scale * pose_scale * pose_err);
    double vel_reward = exp(-err_scale * vel_scale * vel_err);
    double end_eff_reward = exp(-err_scale * end_eff_scale * end_eff_err);
    double root_reward = exp(-err_scale * root_scale * root_err);
    double com_reward = exp(-err_scale * com_scale * com_err);
    return pose_w * pose_reward + vel_w * vel_reward + end_eff_w * end_eff_reward + root_w * root_reward + com_w * com_reward;
}

/* ------------------------------------------------------------------------------------------
 * batched driver: one env-step = stable-PD torque + kin-char pose/vel + FK + terrain + pack
 * 
 * ---------------------------------------------------------------------------------------- */
void trlo_step_range(const trlo_model* m, const trlo_obs_cfg* cfg, const trlo_step_io* io, int env_begin, int env_end)
{
    int n = m->num_dof, N = m->num_joints;
    int F = trlo_poli_state_size(m, cfg);
    for (int e = env_begin; e < env_end; ++e) {
        const double* pose = io->pose + (size_t)e * n;
        const double* vel = io->vel + (size_t)e * n;
        /* (1) stable PD */
        if (io->tau) {
            double* tau = io->tau + (size_t)e * n;
            for (int i = 0; i < n; ++i) tau[i] = 0;
            trlo_imp_pd_control_force(m, io->dt, pose, vel, io->tar_pose + (size_t)e * n, io->tar_vel + (size_t)e * n,
                                      io->joint_active ? io->joint_active + (size_t)e * N : 0, 0, 0, tau, 0, 0);
        }
        /* (2) kinematic reference character and (3) its FK */
        if (io->kin_pose && m->num_frames > 0) {
            double* kp = io->kin_pose + (size_t)e * n;
            trlo_kin_char_pose(m, io->kin_time[e], io->origin_pos ? io->origin_pos + 3 * (size_t)e : 0,
                               io->origin_rot ? io->origin_rot + 4 * (size_t)e : 0, kp,
                               io->kin_vel ? io->kin_vel + (size_t)e * n : 0);
            if (io->kin_body_pos)
                for (int j = 0; j < N; ++j) {
                    double* bp = io->kin_body_pos + ((size_t)e * N + j) * 3;
                    if (trlo_valid_body(m, j)) trlo_body_part_pos(m, kp, j, bp);
                    else bp[0] = bp[1] = bp[2] = 0;
                }
        }
        /* (4) terrain features + pack */
        if (io->state) {
            const trlo_ground* g = 0;
            if (io->num_grounds > 0) g = &io->grounds[io->env_to_ground ? io->env_to_ground[e] : (e % io->num_grounds)];
            trlo_build_poli_state(m, g, cfg, pose, vel, io->body_pos + (size_t)e * N * 3, io->body_vel + (size_t)e * N * 3,
                                  0, 0, 0, io->phase ? io->phase[e] : 0.0, io->flip ? io->flip[e] : 0,
                                  io->state + (size_t)e * F);
        }
    }
}
