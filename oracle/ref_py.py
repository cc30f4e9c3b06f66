"""ctypes binding of oracle/_ref/libtrl_ref.so -- the reference's OWN translation units, compiled unmodified
(oracle/Makefile target `ref`, wrapper oracle/ref_wrap.cpp).

TEST INFRASTRUCTURE: only tests/ and tools/make_golden.py import this module. It is how the C oracle
(oracle/trl_oracle.c) is pinned to the reference: same inputs through both, compared at 1e-13.
The library is built only where /root/reference exists (this container); the built .so travels to the GPU box.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libtrl_ref.so")
REF_ROOT = os.environ.get("TRL_REFERENCE_ROOT", "/root/reference")

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


def have_reference_tree():
    return os.path.isdir(os.path.join(REF_ROOT, "sim"))


def build_lib(force=False):
    """Build oracle/_ref when the reference tree is present; returns the path or None."""
    if have_reference_tree():
        srcs = [os.path.join(_HERE, "ref_wrap.cpp"), os.path.join(_HERE, "eigen_shim", "Eigen", "Dense")]
        stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(LIB_PATH) < os.path.getmtime(s) for s in srcs)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "ref", "REF=" + REF_ROOT])
    return LIB_PATH if os.path.exists(LIB_PATH) else None


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_lib()
        if path is None:
            raise RuntimeError("oracle/_ref/libtrl_ref.so is not built (needs /root/reference)")
        _lib = C.CDLL(path)
        _lib.trlr_model_create.restype = C.c_void_p
        _lib.trlr_kin_char_create.restype = C.c_void_p
        _lib.trlr_ground_create.restype = C.c_void_p
        _lib.trlr_calc_heading.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _u(a):
    return a.ctypes.data_as(_up) if a is not None else None


def _c(a, dtype=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def load_character(char_file):
    """cKinTree::Load / LoadBodyDefs / cPDController::LoadParams on a reference character file."""
    path = char_file.encode()
    N = lib().trlr_load_character(path, None, None, None)
    if N < 0:
        raise RuntimeError("reference failed to load " + char_file)
    jm, bd, pd = np.zeros((N, 19)), np.zeros((N, 17)), np.zeros((N, 18))
    rc = lib().trlr_load_character(path, _d(jm), _d(bd), _d(pd))
    if rc < 0:
        raise RuntimeError("reference failed to load " + char_file)
    return jm, bd, pd


class RefModel:
    """cRBDModel + cKinTree statics + the PD controllers of the reference on flat model tables."""

    def __init__(self, joint_mat, body_defs, pd_params=None, gravity=(0.0, -9.8, 0.0)):
        self.joint_mat = _c(joint_mat)
        self.body_defs = _c(body_defs)
        self.pd_params = _c(pd_params)
        self.N = self.joint_mat.shape[0]
        g = np.array(gravity, dtype=np.float64)
        self.h = C.c_void_p(lib().trlr_model_create(C.c_int(self.N), _d(self.joint_mat), _d(self.body_defs),
                                                   _d(self.pd_params), _d(g)))
        self.n = lib().trlr_num_dof(self.h)

    def __del__(self):
        if getattr(self, "h", None) is not None and _lib is not None:
            _lib.trlr_model_free(self.h)
            self.h = None

    def rbd_update(self, pose, vel):
        pose, vel = _c(pose), _c(vel)
        M, Cb = np.zeros((self.n, self.n)), np.zeros(self.n)
        lib().trlr_rbd_update(self.h, _d(pose), _d(vel), _d(M), _d(Cb))
        return M, Cb

    def inv_dyna(self, pose, vel, acc):
        pose, vel, acc = _c(pose), _c(vel), _c(acc)
        tau = np.zeros(self.n)
        lib().trlr_solve_inv_dyna(self.h, _d(pose), _d(vel), _d(acc), _d(tau))
        return tau

    def inertia_spatial(self, j):
        out = np.zeros((6, 6))
        lib().trlr_inertia_spatial(self.h, C.c_int(j), _d(out))
        return out

    def rbd_extras(self, pose):
        pose = _c(pose)
        J, Jc, g = np.zeros((6, self.n)), np.zeros((6, self.n)), np.zeros(self.n)
        lib().trlr_rbd_extras(self.h, _d(pose), _d(J), _d(Jc), _d(g))
        return J, Jc, g

    def end_eff_jacobian(self, pose, joint_id):
        pose = _c(pose)
        J = np.zeros((6, self.n))
        lib().trlr_end_eff_jacobian(self.h, _d(pose), C.c_int(joint_id), _d(J))
        return J

    def calc_com(self, pose, vel):
        pose, vel = _c(pose), _c(vel)
        com, cv = np.zeros(3), np.zeros(3)
        lib().trlr_calc_com(self.h, _d(pose), _d(vel), _d(com), _d(cv))
        return com, cv

    def child_parent_trans(self, pose, j):
        pose = _c(pose)
        out = np.zeros((4, 4))
        lib().trlr_child_parent_trans(self.h, _d(pose), C.c_int(j), _d(out))
        return out

    def joint_world_trans(self, pose, j):
        pose = _c(pose)
        out = np.zeros((4, 4))
        lib().trlr_joint_world_trans(self.h, _d(pose), C.c_int(j), _d(out))
        return out

    def body_part_pos(self, pose, j):
        pose = _c(pose)
        out = np.zeros(3)
        lib().trlr_body_part_pos(self.h, _d(pose), C.c_int(j), _d(out))
        return out

    def vel_to_pose_diff(self, pose, vel):
        pose, vel = _c(pose), _c(vel)
        out = np.zeros(self.n)
        lib().trlr_vel_to_pose_diff(self.h, _d(pose), _d(vel), _d(out))
        return out

    def post_process_pose(self, pose):
        pose = np.array(pose, dtype=np.float64)
        lib().trlr_post_process_pose(self.h, _d(pose))
        return pose

    def calc_vel(self, pose0, pose1, dt):
        pose0, pose1 = _c(pose0), _c(pose1)
        out = np.zeros(self.n)
        lib().trlr_calc_vel(self.h, _d(pose0), _d(pose1), C.c_double(dt), _d(out))
        return out

    def lerp_poses(self, pose0, pose1, t):
        pose0, pose1 = _c(pose0), _c(pose1)
        out = np.zeros(self.n)
        lib().trlr_lerp_poses(self.h, _d(pose0), _d(pose1), C.c_double(t), _d(out))
        return out

    def calc_heading(self, pose):
        pose = _c(pose)
        return lib().trlr_calc_heading(self.h, _d(pose))

    def build_origin_trans(self, pose):
        pose = _c(pose)
        out = np.zeros((4, 4))
        lib().trlr_build_origin_trans(self.h, _d(pose), _d(out))
        return out

    def pd_valid(self, j):
        return bool(lib().trlr_pd_valid(self.h, C.c_int(j)))

    def imp_pd(self, dt, pose, vel, tar_pose, tar_vel, active=None, kp=None, kd=None, tau0=None):
        """cImpPDController::UpdateControlForce(dt, tau) -> tau, M, C"""
        pose, vel, tar_pose, tar_vel = _c(pose), _c(vel), _c(tar_pose), _c(tar_vel)
        active, kp, kd = _c(active, np.uint8), _c(kp), _c(kd)
        tau = np.zeros(self.n) if tau0 is None else np.array(tau0, dtype=np.float64)
        M, Cb = np.zeros((self.n, self.n)), np.zeros(self.n)
        rc = lib().trlr_imp_pd_control_force(self.h, C.c_double(dt), _d(pose), _d(vel), _d(tar_pose), _d(tar_vel),
                                             _u(active), _d(kp), _d(kd), _d(tau), _d(M), _d(Cb))
        assert rc == 0
        return tau, M, Cb

    def exp_pd(self, dt, pose, vel, tar_pose, tar_vel, active=None, kp=None, kd=None, tau0=None):
        """cExpPDController::UpdateControlForce(dt, tau) -> tau"""
        pose, vel, tar_pose, tar_vel = _c(pose), _c(vel), _c(tar_pose), _c(tar_vel)
        active, kp, kd = _c(active, np.uint8), _c(kp), _c(kd)
        tau = np.zeros(self.n) if tau0 is None else np.array(tau0, dtype=np.float64)
        rc = lib().trlr_exp_pd_control_force(self.h, C.c_double(dt), _d(pose), _d(vel), _d(tar_pose), _d(tar_vel),
                                             _u(active), _d(kp), _d(kd), _d(tau))
        assert rc == 0
        return tau


def quat_to_axis_angle(q):
    q = _c(q)
    axis = np.zeros(3)
    theta = C.c_double()
    lib().trlr_quat_to_axis_angle(_d(q), _d(axis), C.byref(theta))
    return axis, theta.value


class RefKinChar:
    """cKinCharacter of the reference on its own character / motion files."""

    def __init__(self, char_file, motion_file=None):
        h = lib().trlr_kin_char_create(char_file.encode(), motion_file.encode() if motion_file else None)
        if not h:
            raise RuntimeError("reference failed to load %s / %s" % (char_file, motion_file))
        self.h = C.c_void_p(h)
        N, n, F, loop = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        dur = C.c_double()
        lib().trlr_kin_char_dims(self.h, C.byref(N), C.byref(n), C.byref(F), C.byref(loop), C.byref(dur))
        self.N, self.n, self.num_frames, self.loop, self.duration = N.value, n.value, F.value, bool(loop.value), dur.value

    def __del__(self):
        if getattr(self, "h", None) is not None and _lib is not None:
            _lib.trlr_kin_char_free(self.h)
            self.h = None

    def frames(self):
        out = np.zeros((self.num_frames, 1 + self.n))
        lib().trlr_motion_frames(self.h, _d(out))
        return out

    def index_phase(self, t):
        idx, ph = C.c_int(), C.c_double()
        lib().trlr_motion_index_phase(self.h, C.c_double(t), C.byref(idx), C.byref(ph))
        return idx.value, ph.value

    def pose(self, t, origin_pos=None, origin_rot=None):
        op, orr = _c(origin_pos), _c(origin_rot)
        pose, vel = np.zeros(self.n), np.zeros(self.n)
        lib().trlr_kin_char_pose(self.h, C.c_double(t), _d(op), _d(orr), _d(pose), _d(vel))
        return pose, vel

    def read_state(self, state_file):
        pose, vel = np.zeros(self.n), np.zeros(self.n)
        rc = lib().trlr_read_state(self.h, state_file.encode(), _d(pose), _d(vel))
        if rc < 0:
            raise RuntimeError("reference failed to read " + state_file)
        return pose, vel


class RefGround:
    """cGroundVar2D / cGroundVar3D of the reference, built like cGroundFactory from a terrain-parameter file."""

    def __init__(self, param_file, world_scale=1.0, seed=0, blend=0.0):
        h = lib().trlr_ground_create(param_file.encode(), C.c_double(world_scale), C.c_ulong(seed), C.c_double(blend))
        if not h:
            raise RuntimeError("reference failed to build ground from " + param_file)
        self.h = C.c_void_p(h)
        self.kind = lib().trlr_ground_kind(self.h)
        self.num_pieces = lib().trlr_ground_num_pieces(self.h)

    def __del__(self):
        if getattr(self, "h", None) is not None and _lib is not None:
            _lib.trlr_ground_free(self.h)
            self.h = None

    def update(self, bound_min, bound_max, time_elapsed=0.0):
        bmin, bmax = _c(bound_min), _c(bound_max)
        lib().trlr_ground_update(self.h, C.c_double(time_elapsed), _d(bmin), _d(bmax))

    def pieces(self):
        """[{data float32 [res_z][res_x], res (x, z), origin, scale, bounds}] in the reference's scan order."""
        out = []
        for s in range(self.num_pieces):
            res = (C.c_int * 2)()
            origin, scale, bounds = np.zeros(3), np.zeros(3), np.zeros(4)
            count = lib().trlr_ground_piece_info(self.h, C.c_int(s), res, _d(origin), _d(scale), _d(bounds))
            data = np.zeros(count, dtype=np.float32)
            if count:
                lib().trlr_ground_piece_data(self.h, C.c_int(s), data.ctypes.data_as(_fp))
            out.append({"data": data.reshape(res[1], res[0]) if count else data, "res": (res[0], res[1]),
                        "origin": origin, "scale": scale, "bounds": bounds})
        return out

    def sample_height(self, pos):
        pos = _c(pos).reshape(-1, 3)
        n = pos.shape[0]
        h = np.zeros(n)
        valid = np.zeros(n, dtype=np.int32)
        lib().trlr_ground_sample_height(self.h, C.c_int(n), _d(pos), _d(h), valid.ctypes.data_as(_ip))
        return h, valid


def terrain_gen_2d(type_str, width, params=None, seed=0, cap=1 << 16):
    """cTerrainGen2D::Build*(width, params, cRand(seed), out_data) -> float32 heights"""
    params = _c(params)
    out = np.zeros(cap, dtype=np.float32)
    n = lib().trlr_terrain_gen_2d(type_str.encode(), C.c_double(width), _d(params), C.c_ulong(seed),
                                  out.ctypes.data_as(_fp), C.c_int(cap))
    assert n <= cap
    return out[:n].copy()


def terrain_params_2d(param_file, idx=0):
    npar = lib().trlr_terrain_gen_2d_num_params()
    out = np.zeros(npar)
    rc = lib().trlr_terrain_params_2d(param_file.encode(), C.c_int(idx), _d(out))
    if rc < 0:
        raise RuntimeError("reference failed to parse " + param_file)
    return out


def terrain_params_2d_default():
    """cTerrainGen2D::GetDefaultParams through a terrain file without a Params block"""
    return _default_2d()


def _default_2d():
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write('{"Type": "var2d_flat", "Params": [{}]}')
        path = f.name
    try:
        return terrain_params_2d(path)
    finally:
        os.unlink(path)
