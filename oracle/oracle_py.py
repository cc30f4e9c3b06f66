"""ctypes binding of the CPU oracle (oracle/libtrl_oracle.so).

TEST INFRASTRUCTURE -- may be imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. The product never imports this module.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtrl_oracle.so")
_DATA = os.path.join(_HERE, "..", "terrainrlsim_b200", "data")

JOINT_COLS, BODY_COLS, PD_COLS = 19, 17, 18
GROUND_NONE, GROUND_VAR2D, GROUND_VAR3D, GROUND_PLANE = 0, 1, 2, 3
OBS_TERRAIN_RL, OBS_TERRAIN_RL_HOPPER, OBS_CT, OBS_CT_3D = 0, 1, 2, 3
SAMPLER_LINE, SAMPLER_GRID = 0, 1

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


class CModel(C.Structure):
    _fields_ = [("num_joints", C.c_int), ("num_dof", C.c_int), ("joint_mat", _dp), ("body_defs", _dp),
                ("pd_params", _dp), ("gravity", C.c_double * 3), ("num_frames", C.c_int), ("frames", _dp),
                ("loop", C.c_int)]


class CGround(C.Structure):
    _fields_ = [("kind", C.c_int), ("num_pieces", C.c_int), ("data", _fp * 4), ("res", (C.c_int * 2) * 4),
                ("origin", (C.c_double * 3) * 4), ("scale", (C.c_double * 3) * 4), ("bounds", (C.c_double * 4) * 4)]


class CObsCfg(C.Structure):
    _fields_ = [("obs_kind", C.c_int), ("sampler", C.c_int), ("num_samples", C.c_int), ("grid_res", C.c_int),
                ("view_min", C.c_double), ("view_max", C.c_double), ("num_phase_bins", C.c_int)]


class CStepIO(C.Structure):
    _fields_ = [("num_envs", C.c_int), ("dt", C.c_double),
                ("pose", _dp), ("vel", _dp), ("tar_pose", _dp), ("tar_vel", _dp), ("joint_active", _up),
                ("kin_time", _dp), ("origin_pos", _dp), ("origin_rot", _dp), ("body_pos", _dp), ("body_vel", _dp),
                ("phase", _dp), ("flip", _up), ("grounds", C.POINTER(CGround)), ("env_to_ground", _ip),
                ("num_grounds", C.c_int),
                ("tau", _dp), ("kin_pose", _dp), ("kin_vel", _dp), ("kin_body_pos", _dp), ("state", _dp)]


def build_lib(force=False):
    if force or not os.path.exists(_LIB_PATH) or \
            any(os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, f))
                for f in ("trl_oracle.c", "trl_oracle_terrain.c", "trl_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_lib())
        _lib.trlo_sample_height.restype = C.c_double
        _lib.trlo_calc_heading.restype = C.c_double
        _lib.trlo_imitate_reward.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _u(a):
    return a.ctypes.data_as(_up) if a is not None else None


def _c(a, dtype=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def load_model_json(name):
    with open(os.path.join(_DATA, name + ".json")) as f:
        return json.load(f)


class Model:
    """Flat model arrays + the C struct the oracle consumes."""

    def __init__(self, name=None, joint_mat=None, body_defs=None, pd_params=None, frames=None, loop=False,
                 gravity=(0.0, -9.8, 0.0), with_motion=True):
        self.meta = {}
        if name is not None:
            d = load_model_json(name)
            self.meta = d
            joint_mat, body_defs, pd_params = d["joint_mat"], d["body_defs"], d["pd_params"]
            if with_motion and "motion" in d:
                frames = np.array(d["motion"]["frames"], dtype=np.float64)
                loop = d["motion"]["loop"]
        self.name = name
        self.joint_mat = np.ascontiguousarray(joint_mat, dtype=np.float64)
        self.body_defs = np.ascontiguousarray(body_defs, dtype=np.float64)
        self.pd_params = np.ascontiguousarray(pd_params, dtype=np.float64)
        self.N = self.joint_mat.shape[0]
        assert self.joint_mat.shape == (self.N, JOINT_COLS)
        assert self.body_defs.shape == (self.N, BODY_COLS)
        assert self.pd_params.shape == (self.N, PD_COLS)
        self.gravity = np.array(gravity, dtype=np.float64)
        self.loop = bool(loop)
        self.frames_raw = None
        self.frames = None
        if frames is not None:
            self.frames_raw = np.ascontiguousarray(frames, dtype=np.float64)
            fr = self.frames_raw.copy()
            # cMotion::PostProcessFrames (anim/Motion.cpp:207-217): durations -> cumulative start times
            t = 0.0
            for f in range(fr.shape[0]):
                dur = fr[f, 0]
                fr[f, 0] = t
                t += dur
            self.frames = fr
        self.parent = self.joint_mat[:, 1].astype(int)
        self.jtype = self.joint_mat[:, 0].astype(int)
        self.offset = self.joint_mat[:, 18].astype(int)
        sizes = {0: 1, 1: 3, 2: 1, 3: 0, 4: 4, 5: 0}
        self.size = np.array([7 if self.parent[j] == -1 else sizes[self.jtype[j]] for j in range(self.N)])
        self.n = int(self.offset[-1] + self.size[-1])
        self.c = CModel()
        self.c.num_joints = self.N
        self.c.num_dof = self.n
        self.c.joint_mat = _d(self.joint_mat)
        self.c.body_defs = _d(self.body_defs)
        self.c.pd_params = _d(self.pd_params)
        self.c.gravity[:] = list(self.gravity)
        self.c.num_frames = 0 if self.frames is None else self.frames.shape[0]
        self.c.frames = _d(self.frames)
        self.c.loop = int(self.loop)

    @property
    def ref(self):
        return C.byref(self.c)

    def quat_slots(self):
        """Offsets of quaternion (w,x,y,z) blocks in a pose vector."""
        out = [3]
        for j in range(1, self.N):
            if self.jtype[j] == 4:
                out.append(int(self.offset[j]))
        return out

    # ---- single-env wrappers ----
    def rbd_update(self, pose, vel):
        pose, vel = _c(pose), _c(vel)
        M = np.zeros((self.n, self.n))
        Cb = np.zeros(self.n)
        lib().trlo_rbd_update(self.ref, _d(pose), _d(vel), _d(M), _d(Cb))
        return M, Cb

    def rbd_extras(self, pose):
        """cRBDUtil::BuildJacobian, BuildCOMJacobian, CalcGravityForce -> J [6][n], Jcom [6][n], tau_g [n]"""
        pose = _c(pose)
        J, Jc, g = np.zeros((6, self.n)), np.zeros((6, self.n)), np.zeros(self.n)
        lib().trlo_rbd_extras(self.ref, _d(pose), _d(J), _d(Jc), _d(g))
        return J, Jc, g

    def apply_control_forces(self, pose, tau):
        """cSimCharacter::ApplyControlForces + cJoint::ApplyTau -> clamped tau [n], body wrenches [N][6] (world)"""
        pose, tau = _c(pose), _c(tau)
        out_tau, wr = np.zeros(self.n), np.zeros((self.N, 6))
        lib().trlo_apply_control_forces(self.ref, _d(pose), _d(tau), _d(out_tau), _d(wr))
        return out_tau, wr

    def end_eff_jacobian(self, pose, joint_id):
        pose = _c(pose)
        J = np.zeros((6, self.n))
        lib().trlo_end_eff_jacobian(self.ref, _d(pose), C.c_int(joint_id), _d(J))
        return J

    def inv_dyna(self, pose, vel, acc):
        pose, vel, acc = _c(pose), _c(vel), _c(acc)
        tau = np.zeros(self.n)
        lib().trlo_solve_inv_dyna(self.ref, _d(pose), _d(vel), _d(acc), _d(tau))
        return tau

    def imp_pd(self, dt, pose, vel, tar_pose, tar_vel, active=None, kp=None, kd=None, tau0=None):
        pose, vel, tar_pose, tar_vel = _c(pose), _c(vel), _c(tar_pose), _c(tar_vel)
        active = _c(active, np.uint8)
        kp, kd = _c(kp), _c(kd)
        tau = np.zeros(self.n) if tau0 is None else np.array(tau0, dtype=np.float64)
        M = np.zeros((self.n, self.n))
        Cb = np.zeros(self.n)
        zp = lib().trlo_imp_pd_control_force(self.ref, C.c_double(dt), _d(pose), _d(vel), _d(tar_pose), _d(tar_vel),
                                             _u(active), _d(kp), _d(kd), _d(tau), _d(M), _d(Cb))
        return tau, M, Cb, zp

    def exp_pd(self, dt, pose, vel, tar_pose, tar_vel, active=None, kp=None, kd=None, tau0=None):
        """cExpPDController::UpdateControlForce(dt, tau) -> tau"""
        pose, vel, tar_pose, tar_vel = _c(pose), _c(vel), _c(tar_pose), _c(tar_vel)
        active = _c(active, np.uint8)
        kp, kd = _c(kp), _c(kd)
        tau = np.zeros(self.n) if tau0 is None else np.array(tau0, dtype=np.float64)
        lib().trlo_exp_pd_control_force(self.ref, C.c_double(dt), _d(pose), _d(vel), _d(tar_pose), _d(tar_vel),
                                        _u(active), _d(kp), _d(kd), _d(tau))
        return tau

    def joint_world_trans(self, pose, j):
        pose = _c(pose)
        out = np.zeros((4, 4))
        lib().trlo_joint_world_trans(self.ref, _d(pose), C.c_int(j), _d(out))
        return out

    def child_parent_trans(self, pose, j):
        pose = _c(pose)
        out = np.zeros((4, 4))
        lib().trlo_child_parent_trans(self.ref, _d(pose), C.c_int(j), _d(out))
        return out

    def body_part_pos(self, pose, j):
        pose = _c(pose)
        out = np.zeros(3)
        lib().trlo_body_part_pos(self.ref, _d(pose), C.c_int(j), _d(out))
        return out

    def inertia_spatial(self, j):
        out = np.zeros((6, 6))
        lib().trlo_inertia_spatial(self.ref, C.c_int(j), _d(out))
        return out

    def vel_to_pose_diff(self, pose, vel):
        pose, vel = _c(pose), _c(vel)
        out = np.zeros(self.n)
        lib().trlo_vel_to_pose_diff(self.ref, _d(pose), _d(vel), _d(out))
        return out

    def post_process_pose(self, pose):
        pose = np.array(pose, dtype=np.float64)
        lib().trlo_post_process_pose(self.ref, _d(pose))
        return pose

    def calc_vel(self, pose0, pose1, dt):
        pose0, pose1 = _c(pose0), _c(pose1)
        out = np.zeros(self.n)
        lib().trlo_calc_vel(self.ref, _d(pose0), _d(pose1), C.c_double(dt), _d(out))
        return out

    def lerp_poses(self, pose0, pose1, t):
        pose0, pose1 = _c(pose0), _c(pose1)
        out = np.zeros(self.n)
        lib().trlo_lerp_poses(self.ref, _d(pose0), _d(pose1), C.c_double(t), _d(out))
        return out

    def motion_index_phase(self, t):
        idx = C.c_int()
        ph = C.c_double()
        lib().trlo_motion_index_phase(self.ref, C.c_double(t), C.byref(idx), C.byref(ph))
        return idx.value, ph.value

    def kin_char_pose(self, t, origin_pos=None, origin_rot=None):
        pose = np.zeros(self.n)
        vel = np.zeros(self.n)
        op, orr = _c(origin_pos), _c(origin_rot)
        lib().trlo_kin_char_pose(self.ref, C.c_double(t), _d(op), _d(orr), _d(pose), _d(vel))
        return pose, vel


def ldlt_solve(A, b):
    A, b = _c(A), _c(b)
    n = A.shape[0]
    x = np.zeros(n)
    zp = lib().trlo_ldlt_solve(C.c_int(n), _d(A), _d(b), _d(x))
    return x, zp


class Ground:
    """One env's terrain: list of pieces (data float32 2-D array [res_z][res_x] or 1-D row)."""

    def __init__(self, kind, pieces=()):
        self.kind = kind
        self.c = CGround()
        self.c.kind = kind
        self.c.num_pieces = len(pieces)
        self._keep = []
        self.pieces = list(pieces)
        for s, p in enumerate(pieces):
            data = np.ascontiguousarray(p["data"], dtype=np.float32)
            self._keep.append(data)
            self.c.data[s] = data.ctypes.data_as(_fp)
            self.c.res[s][0] = int(p["res"][0])
            self.c.res[s][1] = int(p["res"][1])
            for k in range(3):
                self.c.origin[s][k] = float(p["origin"][k])
                self.c.scale[s][k] = float(p["scale"][k])
            for k in range(4):
                self.c.bounds[s][k] = float(p["bounds"][k])

    def sample_height(self, pos):
        pos = _c(pos)
        valid = C.c_int()
        cell = (C.c_int * 3)()
        h = lib().trlo_sample_height(C.byref(self.c), _d(pos), C.byref(valid), cell)
        return h, bool(valid.value), (cell[0], cell[1], cell[2])


def sample_height_batch(ground, pos):
    pos = _c(pos).reshape(-1, 3)
    n = pos.shape[0]
    h = np.zeros(n)
    valid = np.zeros(n, dtype=np.int32)
    cell = np.zeros((n, 3), dtype=np.int32)
    lib().trlo_sample_height_batch(C.byref(ground.c), C.c_int(n), _d(pos), _d(h), valid.ctypes.data_as(_ip),
                                   cell.ctypes.data_as(_ip))
    return h, valid, cell


def obs_cfg(obs_kind=OBS_TERRAIN_RL, sampler=SAMPLER_LINE, num_samples=200, grid_res=0, view_min=-0.5,
            view_max=10.0, num_phase_bins=-1):
    c = CObsCfg()
    c.obs_kind, c.sampler, c.num_samples, c.grid_res = obs_kind, sampler, num_samples, grid_res
    c.view_min, c.view_max, c.num_phase_bins = view_min, view_max, num_phase_bins
    return c


def poli_state_size(model, cfg):
    return lib().trlo_poli_state_size(model.ref, C.byref(cfg))


def parse_ground(ground, cfg, pose, flip=0):
    pose = _c(pose)
    G = cfg.num_samples
    out = np.zeros(G)
    cells = np.zeros((G, 3), dtype=np.int32)
    lib().trlo_parse_ground(C.byref(ground.c) if ground is not None else None, C.byref(cfg), _d(pose), C.c_int(flip),
                            _d(out), cells.ctypes.data_as(_ip))
    return out, cells


def build_poli_state(model, ground, cfg, pose, vel, body_pos, body_vel, body_quat=None, body_angvel=None,
                     ground_vel=None, phase=0.0, flip=0):
    F = poli_state_size(model, cfg)
    out = np.zeros(F)
    pose, vel, body_pos, body_vel = _c(pose), _c(vel), _c(body_pos), _c(body_vel)
    body_quat, body_angvel, ground_vel = _c(body_quat), _c(body_angvel), _c(ground_vel)
    lib().trlo_build_poli_state(model.ref, C.byref(ground.c) if ground is not None else None, C.byref(cfg),
                                _d(pose), _d(vel), _d(body_pos), _d(body_vel), _d(body_quat), _d(body_angvel),
                                _d(ground_vel), C.c_double(phase), C.c_int(flip), _d(out))
    return out


def calc_com(model, pose, vel):
    pose, vel = _c(pose), _c(vel)
    com, cv = np.zeros(3), np.zeros(3)
    lib().trlo_calc_com(model.ref, _d(pose), _d(vel), _d(com), _d(cv))
    return com, cv


def imitate_reward(model, ground, pose0, vel0, pose1, vel1, body_vel, fallen=0):
    pose0, vel0, pose1, vel1, body_vel = _c(pose0), _c(vel0), _c(pose1), _c(vel1), _c(body_vel)
    terms = np.zeros(5)
    r = lib().trlo_imitate_reward(model.ref, C.byref(ground.c) if ground is not None else None, _d(pose0), _d(vel0),
                                  _d(pose1), _d(vel1), _d(body_vel), C.c_int(int(fallen)), _d(terms))
    return r, terms


def step_range(model, cfg, arrays, grounds, dt, env_begin=0, env_end=None, env_to_ground=None):
    """Run the fused oracle step over envs [env_begin, env_end). `arrays` is a dict of numpy arrays
    (inputs: pose, vel, tar_pose, tar_vel, joint_active, kin_time, origin_pos, origin_rot, body_pos, body_vel,
    phase, flip; outputs are allocated if absent: tau, kin_pose, kin_vel, kin_body_pos, state)."""
    E = arrays["pose"].shape[0]
    n, N = model.n, model.N
    F = poli_state_size(model, cfg)
    io = CStepIO()
    io.num_envs = E
    io.dt = dt
    keep = []

    def get(name, dtype=np.float64):
        a = arrays.get(name)
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        arrays[name] = a
        return a

    for name in ("pose", "vel", "tar_pose", "tar_vel", "kin_time", "origin_pos", "origin_rot", "body_pos",
                 "body_vel", "phase"):
        setattr(io, name, _d(get(name)))
    io.joint_active = _u(get("joint_active", np.uint8))
    io.flip = _u(get("flip", np.uint8))
    garr = None
    if grounds:
        garr = (CGround * len(grounds))(*[g.c for g in grounds])
        io.grounds = garr
        io.num_grounds = len(grounds)
        if env_to_ground is not None:
            e2g = np.ascontiguousarray(env_to_ground, dtype=np.int32)
            keep.append(e2g)
            io.env_to_ground = e2g.ctypes.data_as(_ip)
    shapes = {"tau": (E, n), "kin_pose": (E, n), "kin_vel": (E, n), "kin_body_pos": (E, N, 3), "state": (E, F)}
    for name, shp in shapes.items():
        if name not in arrays or arrays[name] is None:
            arrays[name] = np.zeros(shp)
        setattr(io, name, _d(arrays[name]))
    if env_end is None:
        env_end = E
    lib().trlo_step_range(model.ref, C.byref(cfg), C.byref(io), C.c_int(env_begin), C.c_int(env_end))
    return arrays


# ----------------------------------------------------------------------------------------------
# terrain generation and scrolling (oracle/trl_oracle_terrain.c)
# ----------------------------------------------------------------------------------------------
GROUND_TYPE_NAMES = ["plane", "var2d_flat", "var2d_gaps", "var2d_steps", "var2d_walls", "var2d_bumps", "var2d_mixed",
                     "var2d_narrow_gaps", "var2d_slopes", "var2d_slopes_gaps", "var2d_slopes_walls",
                     "var2d_slopes_steps", "var2d_slopes_mixed", "var2d_slopes_narrow_gaps", "var2d_cliffs",
                     "var3d_flat"]  # cGround::eType order, sim/Ground.cpp:4-33
TERRAIN2D_PARAM_NAMES = [  # cTerrainGen2D::gParamDefs, sim/TerrainGen2D.cpp:12-63
    "GapSpacingMin", "GapSpacingMax", "GapWMin", "GapWMax", "GapHMin", "GapHMax",
    "WallSpacingMin", "WallSpacingMax", "WallWMin", "WallWMax", "WallHMin", "WallHMax",
    "StepSpacingMin", "StepSpacingMax", "StepH0Min", "StepH0Max", "StepH1Min", "StepH1Max",
    "BumpHMin", "BumpHMax",
    "NarrowGapSpacingMin", "NarrowGapSpacingMax", "NarrowGapDistMin", "NarrowGapDistMax", "NarrowGapWMin",
    "NarrowGapWMax", "NarrowGapDepthMin", "NarrowGapDepthMax", "NarrowGapCountMin", "NarrowGapCountMax",
    "CliffSpacingMin", "CliffSpacingMax", "CliffH0Min", "CliffH0Max", "CliffH1Min", "CliffH1Max", "CliffMiniCountMax",
    "SlopeDeltaRange", "SlopeDeltaMin", "SlopeDeltaMax"]
PIECE_CAP = 49152


class CRand(C.Structure):
    _fields_ = [("x", C.c_ulonglong)]


class CTerrainPiece(C.Structure):
    _fields_ = [("data", C.c_float * PIECE_CAP), ("count", C.c_int), ("res", C.c_int * 2), ("origin", C.c_double * 3),
                ("scale", C.c_double * 3), ("aabb", C.c_double * 4), ("min", C.c_double * 3), ("max", C.c_double * 3)]


class CGroundVar2D(C.Structure):
    _fields_ = [("type", C.c_int), ("params", C.c_double * 40), ("ground_width", C.c_double),
                ("world_scale", C.c_double), ("rand", CRand), ("flip_seg", C.c_int), ("piece", CTerrainPiece * 2)]


class CGroundVar3D(C.Structure):
    _fields_ = [("ground_width", C.c_double), ("spacing_x", C.c_double), ("spacing_z", C.c_double),
                ("world_scale", C.c_double), ("slab_order", C.c_int * 4), ("piece", CTerrainPiece * 4)]


def terrain2d_default_params():
    out = np.zeros(lib().trlo_terrain2d_num_params())
    lib().trlo_terrain2d_default_params(_d(out))
    return out


def load_terrain_file(path):
    """cGroundFactory::ParseParamsJson (sim/GroundFactory.cpp:11-61) + cTerrainGen2D::LoadParams: returns
    {type (int), type_name, ground_width, spacing_x, spacing_z, params [num_sets][40] (2-D) }"""
    with open(path) as f:
        d = json.load(f)
    name = d.get("Type", "var2d_flat")
    out = {"type_name": name, "type": GROUND_TYPE_NAMES.index(name), "ground_width": float(d.get("GroundWidth", 20)),
           "spacing_x": float(d.get("VertSpacingX", 0.2)), "spacing_z": float(d.get("VertSpacingZ", 0.2)), "params": None}
    if name.startswith("var2d") and d.get("Params"):
        rows = []
        for entry in d["Params"]:
            p = terrain2d_default_params()
            for i, key in enumerate(TERRAIN2D_PARAM_NAMES):
                if key in entry and entry[key] is not None:
                    p[i] = float(entry[key])
            rows.append(p)
        out["params"] = np.array(rows)
    return out


def terrain_gen_2d(type_id, width, params=None, seed=0, cap=1 << 16):
    """cTerrainGen2D::Build*(width, params, cRand(seed), out_data) -> float32 heights"""
    rnd = CRand()
    lib().trlo_rand_seed(C.byref(rnd), C.c_ulong(seed))
    params = _c(params)
    out = np.zeros(cap, dtype=np.float32)
    count = C.c_int(0)
    lib().trlo_terrain_gen_2d.restype = C.c_double
    lib().trlo_terrain_gen_2d(C.c_int(type_id), C.c_double(width), _d(params), C.byref(rnd), out.ctypes.data_as(_fp),
                              C.byref(count), C.c_int(cap))
    return out[:count.value].copy()


def _pieces_from_view(view, n):
    out = []
    for s in range(n):
        rx, rz = view.res[s][0], view.res[s][1]
        data = np.ctypeslib.as_array(view.data[s], shape=(rz * rx,)).copy().reshape(rz, rx)
        out.append({"data": data, "res": (rx, rz), "origin": np.array(view.origin[s][:]),
                    "scale": np.array(view.scale[s][:]), "bounds": np.array(view.bounds[s][:])})
    return out


class GroundVar2D:
    """cGroundVar2D restated: two scrolling segments generated by cTerrainGen2D"""

    def __init__(self, type_id, params=None, ground_width=20.0, world_scale=1.0, seed=0, blend=0.0):
        self.c = CGroundVar2D()
        if params is not None:
            params = np.atleast_2d(np.asarray(params, dtype=np.float64))
            # cGround::CalcBlendParams (sim/Ground.cpp:255-278): lerp of two consecutive parameter sets
            b = min(max(blend, 0.0), params.shape[0] - 1.0)
            i0 = int(b)
            i1 = min(i0 + 1, params.shape[0] - 1)
            b -= i0
            params = np.ascontiguousarray((1 - b) * params[i0] + b * params[i1])
        half = ground_width / 2
        lib().trlo_ground2d_init(C.byref(self.c), C.c_int(type_id), _d(params), C.c_double(ground_width),
                                 C.c_double(world_scale), C.c_ulong(seed), C.c_double(-half - 1.0), C.c_double(half - 1.0))

    def update(self, bound_min_x, bound_max_x):
        lib().trlo_ground2d_update(C.byref(self.c), C.c_double(bound_min_x), C.c_double(bound_max_x))

    def pieces(self):
        view = CGround()
        lib().trlo_ground2d_view(C.byref(self.c), C.byref(view))
        return _pieces_from_view(view, 2)


class GroundVar3D:
    """cGroundVar3D restated with cTerrainGen3D::BuildFlat: four scrolling slabs"""

    def __init__(self, spacing_x=0.2, spacing_z=0.2, world_scale=1.0, ground_width=20.0):
        self.c = CGroundVar3D()
        half = ground_width / 2  # cGroundFactory::BuildGroundVar3D bounds use the *file's* width (default 20)
        bmin = np.array([-half, 0.0, -half])
        bmax = np.array([half, 0.0, half])
        lib().trlo_ground3d_init(C.byref(self.c), C.c_double(spacing_x), C.c_double(spacing_z), C.c_double(world_scale),
                                 _d(bmin), _d(bmax))

    def update(self, bound_min, bound_max):
        bmin, bmax = _c(bound_min), _c(bound_max)
        lib().trlo_ground3d_update(C.byref(self.c), _d(bmin), _d(bmax))

    def pieces(self):
        view = CGround()
        lib().trlo_ground3d_view(C.byref(self.c), C.byref(view))
        return _pieces_from_view(view, 4)
