"""Deterministic synthetic inputs for the batched control-and-observation step.

Host-side numpy only (counter-based Philox streams), so the CPU oracle and the CUDA path
consume *identical* arrays. Nothing here touches the oracle.

Layouts follow the reference's conventions:
  pose  [E][n]  root (x y z | qw qx qy qz) then joints in index order   (sim/SimCharacter.cpp:1037-1115)
  vel   [E][n]  root (vx vy vz | wx wy wz 0), spherical (wx wy wz 0)     (sim/Joint.cpp:353-417)
  terrain pieces: float heights already multiplied by world_scale, origin/scale/bounds as the doubles the
  reference would read back from Bullet's float state (sim/GroundVar2D.cpp:425-492,586-591; GroundVar3D.cpp:438-483)
"""
import json
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
WORLD_SCALE = 4.0  # -world_scale= 4 in every config's arg file

JOINT_REVOLUTE, JOINT_PLANAR, JOINT_PRISMATIC, JOINT_FIXED, JOINT_SPHERICAL = 0, 1, 2, 3, 4
_SIZES = {0: 1, 1: 3, 2: 1, 3: 0, 4: 4, 5: 0}

CONFIG_IDS = {"biped2d": 0, "hopper": 1, "dog": 2, "raptor": 3, "biped3d": 4}


def load_model_dict(name):
    with open(os.path.join(DATA_DIR, name + ".json")) as f:
        return json.load(f)


class ModelInfo:
    """Joint layout derived from a joint_mat (host-side helper for building inputs)."""

    def __init__(self, d):
        self.d = d
        jm = np.array(d["joint_mat"], dtype=np.float64)
        self.joint_mat = jm
        self.N = jm.shape[0]
        self.parent = jm[:, 1].astype(int)
        self.jtype = jm[:, 0].astype(int)
        self.offset = jm[:, 18].astype(int)
        self.size = np.array([7 if self.parent[j] == -1 else _SIZES[self.jtype[j]] for j in range(self.N)])
        self.n = int(self.offset[-1] + self.size[-1])
        self.lim_low = jm[:, 8]
        self.lim_high = jm[:, 11]
        self.is3d = bool(np.any(self.jtype[1:] == JOINT_SPHERICAL))


def _rng(seed, stream):
    return np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream)]))


def quat_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx], axis=-1)


def axis_angle_quat(axis, theta):
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    h = 0.5 * theta
    return np.concatenate([np.cos(h)[..., None], np.sin(h)[..., None] * axis], axis=-1)


def _rand_rot(rng, E, max_angle, planar):
    if planar:
        axis = np.tile(np.array([0.0, 0.0, 1.0]), (E, 1))
    else:
        axis = rng.normal(size=(E, 3))
    theta = rng.uniform(-max_angle, max_angle, size=E)
    return axis_angle_quat(axis, theta)


def make_inputs(name_or_dict, E, seed=None, rank=0):
    """Seeded inputs for E envs of one config. seed default = 0xB200 + 1000*config_id + rank."""
    d = load_model_dict(name_or_dict) if isinstance(name_or_dict, str) else name_or_dict
    mi = ModelInfo(d)
    n, N = mi.n, mi.N
    if seed is None:
        seed = 0xB200 + 1000 * CONFIG_IDS.get(d.get("name", ""), 9) + rank
    planar = not mi.is3d
    st = d.get("state", {})
    pose0 = np.array(st.get("pose", [0, 1, 0, 1, 0, 0, 0] + [0] * (n - 7)), dtype=np.float64)
    r = _rng(seed, 1)
    pose = np.tile(pose0, (E, 1))
    pose[:, 0] = r.uniform(-8, 8, E)
    pose[:, 2] = 0.0 if planar else r.uniform(-8, 8, E)
    pose[:, 1] = pose0[1] + r.uniform(-0.05, 0.05, E)
    q0 = np.tile(pose0[3:7] / np.linalg.norm(pose0[3:7]), (E, 1))
    pose[:, 3:7] = quat_mul(q0, _rand_rot(r, E, 0.3, planar))
    for j in range(1, N):
        o = mi.offset[j]
        t = mi.jtype[j]
        if t == JOINT_REVOLUTE:
            lo, hi = mi.lim_low[j], mi.lim_high[j]
            if lo < hi:
                pose[:, o] = np.clip(r.uniform(lo, hi, E), -2.5, 2.5)
            else:
                pose[:, o] = pose0[o] + r.uniform(-0.5, 0.5, E)
        elif t == JOINT_PRISMATIC:
            pose[:, o] = r.uniform(-0.1, 0.1, E)
        elif t == JOINT_PLANAR:
            pose[:, o:o + 3] = r.uniform(-0.1, 0.1, (E, 3))
        elif t == JOINT_SPHERICAL:
            pose[:, o:o + 4] = _rand_rot(r, E, 0.6, False)
    for o in [3] + [mi.offset[j] for j in range(1, N) if mi.jtype[j] == JOINT_SPHERICAL]:
        pose[:, o:o + 4] /= np.linalg.norm(pose[:, o:o + 4], axis=1, keepdims=True)

    r = _rng(seed, 2)
    vel = np.zeros((E, n))
    vel[:, 0:3] = r.normal(0, 1, (E, 3))
    if planar:
        vel[:, 2] = 0
        vel[:, 5] = r.normal(0, 1, E)
    else:
        vel[:, 3:6] = r.normal(0, 1, (E, 3))
    for j in range(1, N):
        o, s, t = mi.offset[j], mi.size[j], mi.jtype[j]
        if t == JOINT_SPHERICAL:
            vel[:, o:o + 3] = r.normal(0, 2, (E, 3))
        elif s > 0:
            vel[:, o:o + s] = r.normal(0, 2, (E, s))

    r = _rng(seed, 3)
    tar_pose = pose + r.normal(0, 0.2, (E, n))
    tar_pose[:, 0:7] = 0
    for j in range(1, N):
        if mi.jtype[j] == JOINT_SPHERICAL:
            o = mi.offset[j]
            tar_pose[:, o:o + 4] = quat_mul(pose[:, o:o + 4], _rand_rot(r, E, 0.4, False))
            tar_pose[:, o:o + 4] /= np.linalg.norm(tar_pose[:, o:o + 4], axis=1, keepdims=True)
    tar_vel = np.zeros((E, n))

    r = _rng(seed, 4)
    joint_active = np.ones((E, N), dtype=np.uint8)
    if d.get("ctrl", "") in ("dog", "raptor", "biped3D_cacla", "monoped_hopper") and N > 1:
        # FSM configs: the stance hip (first child of the root) is inactive in 50% of envs (exercises Q2)
        hip = 1
        joint_active[:, hip] = (r.uniform(size=E) < 0.5).astype(np.uint8)

    r = _rng(seed, 5)
    dur = 1.0
    if "motion" in d:
        dur = float(sum(f[0] for f in d["motion"]["frames"][:-1]))
    kin_time = r.uniform(0, 10 * dur, E)
    origin_pos = np.zeros((E, 3))
    origin_pos[:, 0] = pose[:, 0]
    origin_pos[:, 2] = pose[:, 2]
    origin_rot = np.tile(np.array([1.0, 0, 0, 0]), (E, 1))
    if not planar:
        origin_rot = axis_angle_quat(np.tile(np.array([0.0, 1.0, 0.0]), (E, 1)), r.uniform(-3.0, 3.0, E))

    r = _rng(seed, 6)
    body_pos = pose[:, None, 0:3] + r.normal(0, 0.4, (E, N, 3))
    body_vel = vel[:, None, 0:3] + r.normal(0, 1.0, (E, N, 3))
    if planar:
        body_pos[:, :, 2] = 0
        body_vel[:, :, 2] = 0
    phase = r.uniform(0, 1, E)
    flip = np.zeros(E, dtype=np.uint8)
    if d.get("ctrl", "") == "biped3D_cacla":
        flip = (r.uniform(size=E) < 0.5).astype(np.uint8)
    return {"pose": pose, "vel": vel, "tar_pose": tar_pose, "tar_vel": tar_vel, "joint_active": joint_active,
            "kin_time": kin_time, "origin_pos": origin_pos, "origin_rot": origin_rot,
            "body_pos": np.ascontiguousarray(body_pos), "body_vel": np.ascontiguousarray(body_vel),
            "phase": phase, "flip": flip}


# ----------------------------------------------------------------------------------------------
# body states derived from the pose ("FK of pose and Jacobian-free finite difference")
# ----------------------------------------------------------------------------------------------
def _quat_rot(q):
    """[E][4] (w, x, y, z) unit quaternions -> [E][3][3] rotation matrices."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _euler_zyx(t):
    """Rz * Ry * Rx of the attach angles (x, y, z), the reference's attach convention (util/MathUtil.cpp:128-155)."""
    cx, sx, cy, sy, cz, sz = np.cos(t[0]), np.sin(t[0]), np.cos(t[1]), np.sin(t[1]), np.cos(t[2]), np.sin(t[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def fk_body_pos(name_or_dict, pose):
    """World positions [E][N][3] of the body parts for poses [E][n]: plain numpy forward kinematics in the conventions of
    cKinTree::CalcBodyPartPos (anim/KinTree.cpp:299-308,1055-1139). Input preparation only (not a checker, not timed)."""
    d = load_model_dict(name_or_dict) if isinstance(name_or_dict, str) else name_or_dict
    mi = ModelInfo(d)
    jm, bd = mi.joint_mat, np.array(d["body_defs"], dtype=np.float64)
    E = pose.shape[0]
    Rw = [None] * mi.N
    tw = [None] * mi.N
    out = np.empty((E, mi.N, 3))
    for j in range(mi.N):
        o = mi.offset[j]
        if mi.parent[j] < 0:
            q = pose[:, 3:7] / np.linalg.norm(pose[:, 3:7], axis=1, keepdims=True)
            Rw[j], tw[j] = _quat_rot(q), pose[:, 0:3].copy()
        else:
            A = _euler_zyx(jm[j, 5:8])
            a = jm[j, 2:5]
            t = mi.jtype[j]
            Rl = np.tile(A, (E, 1, 1))
            tl = np.tile(a, (E, 1))
            if t == JOINT_REVOLUTE:
                th = pose[:, o]
                Rz = np.zeros((E, 3, 3))
                Rz[:, 0, 0] = np.cos(th); Rz[:, 0, 1] = -np.sin(th); Rz[:, 1, 0] = np.sin(th); Rz[:, 1, 1] = np.cos(th); Rz[:, 2, 2] = 1
                Rl = A @ Rz
            elif t == JOINT_PRISMATIC:
                tl = tl + pose[:, o:o + 1] * A[:, 0]
            elif t == JOINT_PLANAR:
                th = pose[:, o + 2]
                Rz = np.zeros((E, 3, 3))
                Rz[:, 0, 0] = np.cos(th); Rz[:, 0, 1] = -np.sin(th); Rz[:, 1, 0] = np.sin(th); Rz[:, 1, 1] = np.cos(th); Rz[:, 2, 2] = 1
                Rl = A @ Rz
                loc = np.zeros((E, 3))
                loc[:, 0:2] = pose[:, o:o + 2]
                tl = tl + np.einsum("eij,ej->ei", Rl, loc)
            elif t == JOINT_SPHERICAL:
                q = pose[:, o:o + 4] / np.linalg.norm(pose[:, o:o + 4], axis=1, keepdims=True)
                Rl = A @ _quat_rot(q)
            pj = mi.parent[j]
            Rw[j] = Rw[pj] @ Rl
            tw[j] = tw[pj] + np.einsum("eij,ej->ei", Rw[pj], tl)
        out[:, j] = tw[j] + np.einsum("eij,j->ei", Rw[j], bd[j, 4:7])
    return out


def derive_body_state(name_or_dict, x, h=1e-5):
    """Replaces x['body_pos'] / x['body_vel'] by the FK of x['pose'] and its finite difference along x['vel'] (root
    angular velocity in the world frame, spherical joint velocities in the joint frame), so the observation features
    are consistent with the pose. Returns x."""
    d = load_model_dict(name_or_dict) if isinstance(name_or_dict, str) else name_or_dict
    mi = ModelInfo(d)
    pose, vel = x["pose"], x["vel"]
    E = pose.shape[0]
    p2 = pose + h * vel
    zeros = np.zeros((E, 1))
    w = np.concatenate([zeros, vel[:, 3:6]], axis=1)
    p2[:, 3:7] = pose[:, 3:7] + 0.5 * h * quat_mul(w, pose[:, 3:7])
    for j in range(1, mi.N):
        if mi.jtype[j] == JOINT_SPHERICAL:
            o = mi.offset[j]
            w = np.concatenate([zeros, vel[:, o:o + 3]], axis=1)
            p2[:, o:o + 4] = pose[:, o:o + 4] + 0.5 * h * quat_mul(pose[:, o:o + 4], w)
    b0 = fk_body_pos(d, pose)
    b1 = fk_body_pos(d, p2)
    x["body_pos"] = np.ascontiguousarray(b0)
    x["body_vel"] = np.ascontiguousarray((b1 - b0) / h)
    return x


# ----------------------------------------------------------------------------------------------
# terrain
# ----------------------------------------------------------------------------------------------
def _f32(x):
    return np.float32(x)


def _profile_2d(rng, width_m, kind):
    """A float32 height profile on the 0.025f grid: flat, or 'mixed' (gaps, walls, steps, slopes).
    Our own generator (not cTerrainGen2D); keeps the reference's property that a builder appends whole
    features until the width is reached, so the vertex count varies per segment."""
    spacing = float(np.float32(0.025))
    target = int(np.ceil(width_m / spacing)) + 1
    h = []
    cur = 0.0
    if kind == "flat":
        return np.zeros(target, dtype=np.float32)
    while len(h) < target:
        feat = rng.integers(0, 5)
        run = int(rng.uniform(3.0, 7.0) / spacing)
        h += [cur] * run
        if feat == 0:      # gap
            w = int(rng.uniform(0.5, 2.0) / spacing)
            h += [cur - 2.0] * w
        elif feat == 1:    # wall
            w = max(2, int(0.2 / spacing))
            h += [cur + rng.uniform(0.25, 0.5)] * w
        elif feat == 2:    # step
            cur += rng.uniform(-0.4, 0.4)
        elif feat == 3:    # slope
            w = int(rng.uniform(1.0, 3.0) / spacing)
            slope = rng.uniform(-0.3, 0.3)
            h += [cur + slope * spacing * i for i in range(w)]
            cur += slope * spacing * w
        else:              # bumps
            w = int(rng.uniform(1.0, 2.0) / spacing)
            h += list(cur + 0.05 * rng.normal(size=w))
    return np.array(h, dtype=np.float32)


def make_ground_2d(seed, idx, kind="mixed", ground_width=20.0):
    """Two segments [min | max] meeting at x = 0 (cGroundVar2D::InitSegments, sim/GroundVar2D.cpp:271-291)."""
    rng = _rng(seed, 1000 + idx)
    spacing = float(np.float32(0.025))
    pieces = []
    for s in range(2):
        prof = _profile_2d(rng, ground_width, kind)
        if s == 0:
            prof = (prof - prof[-1]).astype(np.float32)   # eAlignMax: end height pinned to fix point 0
        else:
            prof = (prof - prof[0]).astype(np.float32)    # eAlignMin
        w = prof.shape[0]
        width = (w - 1) * spacing
        min_x = -width if s == 0 else 0.0
        data_row = (prof * _f32(WORLD_SCALE)).astype(np.float32)
        data = np.tile(data_row, (3, 1))  # gGridLength = 3 identical rows (:457-466)
        hmin, hmax = float(prof.min()), float(prof.max())
        origin = np.array([0.5 * (min_x + min_x + width), 0.5 * (hmin + hmax), 0.0])
        origin = (origin * WORLD_SCALE).astype(np.float32).astype(np.float64) / WORLD_SCALE  # through btScalar
        scale = np.array([float(_f32(spacing * WORLD_SCALE)) / WORLD_SCALE, 1.0 / WORLD_SCALE,
                          float(_f32(1.0 * WORLD_SCALE)) / WORLD_SCALE])
        margin = 0.04 / WORLD_SCALE  # bullet collision margin in world units
        bounds = np.array([float(_f32(origin[0] - 0.5 * width - margin)), float(_f32(origin[0] + 0.5 * width + margin)),
                           -1.0, 1.0])
        pieces.append({"data": data, "res": (w, 3), "origin": origin, "scale": scale, "bounds": bounds})
    return pieces


def make_ground_3d(seed, idx, noise=0.0, slab_size=40.0):
    """Four slabs around the origin (cGroundVar3D, sim/GroundVar3D.cpp:38-141,389-420): 201 x 201 floats each."""
    rng = _rng(seed, 2000 + idx)
    spacing = 0.2
    res = int(slab_size / spacing) + 1
    pieces = []
    for s in range(4):
        min_x = -slab_size if (s % 2 == 0) else 0.0
        min_z = -slab_size if (s // 2 == 0) else 0.0
        if noise > 0:
            h = rng.uniform(-noise, noise, (res, res)).astype(np.float32)
        else:
            h = np.zeros((res, res), dtype=np.float32)
        data = (h * _f32(WORLD_SCALE)).astype(np.float32)
        hmin, hmax = float(h.min()), float(h.max())
        origin = np.array([min_x + 0.5 * slab_size, 0.5 * (hmin + hmax), min_z + 0.5 * slab_size])
        origin = (origin * WORLD_SCALE).astype(np.float32).astype(np.float64) / WORLD_SCALE
        scale = np.array([float(_f32(spacing * WORLD_SCALE)) / WORLD_SCALE, 1.0 / WORLD_SCALE,
                          float(_f32(spacing * WORLD_SCALE)) / WORLD_SCALE])
        bounds = np.array([min_x, min_x + slab_size, min_z, min_z + slab_size])
        pieces.append({"data": data, "res": (res, res), "origin": origin, "scale": scale, "bounds": bounds})
    return pieces


def make_grounds(name_or_dict, num, seed=None, rank=0, stress=True):
    """`num` distinct terrains for a config. Returns (kind, list of piece lists); kind 1 = var2d, 2 = var3d."""
    d = load_model_dict(name_or_dict) if isinstance(name_or_dict, str) else name_or_dict
    if seed is None:
        seed = 0xB200 + 1000 * CONFIG_IDS.get(d.get("name", ""), 9) + rank
    terrain = d.get("terrain", "var2d_flat")
    out = []
    if terrain.startswith("var3d"):
        for i in range(num):
            out.append(make_ground_3d(seed, i, noise=(0.3 if (stress and i % 2 == 1) else 0.0)))
        return 2, out
    kind = "mixed" if "mixed" in terrain else "flat"
    for i in range(num):
        out.append(make_ground_2d(seed, i, kind=kind))
    return 1, out


def terrain_setup(name_or_dict, num, seed=None, rank=0):
    """The config's own terrain (its arg file's -terrain_file=, stored in data/<config>.json as `terrain_cfg`) for `num`
    distinct terrains: (type, params [40], ground_width, world_scale, spacing_x, spacing_z, seeds [num] uint64).
    The same tuple drives trl_ground_generate on the device and the oracle's GroundVar2D / 3D on the host, which build
    bit-identical terrains (tests/test_terrain.py)."""
    d = load_model_dict(name_or_dict) if isinstance(name_or_dict, str) else name_or_dict
    if seed is None:
        seed = 0xB200 + 1000 * CONFIG_IDS.get(d.get("name", ""), 9) + rank
    tc = d["terrain_cfg"]
    seeds = (np.uint64(seed) * np.uint64(7919) + np.arange(num, dtype=np.uint64) * np.uint64(104729) + np.uint64(1)) % np.uint64(2147483647)
    return {"type": int(tc["type"]), "params": None if tc["params"] is None else np.array(tc["params"], dtype=np.float64),
            "ground_width": float(tc["ground_width"]),
            "world_scale": WORLD_SCALE, "spacing_x": float(tc["spacing_x"]), "spacing_z": float(tc["spacing_z"]),
            "seeds": seeds.astype(np.uint64)}


def obs_config(name_or_dict):
    """Observation layout of a config."""
    d = load_model_dict(name_or_dict) if isinstance(name_or_dict, str) else name_or_dict
    obs = d.get("obs", "terrain_rl")
    G = int(d.get("ground_samples", 200))
    if obs == "ct":
        return {"obs_kind": 2, "sampler": 0, "num_samples": G, "grid_res": 0, "view_min": -0.5, "view_max": 10.0,
                "num_phase_bins": 4}
    if obs == "terrain_rl_hopper":
        return {"obs_kind": 1, "sampler": 0, "num_samples": G, "grid_res": 0, "view_min": -0.5, "view_max": 10.0,
                "num_phase_bins": -1}
    if obs == "terrain_rl_grid":
        res = int(round(G ** 0.5))
        return {"obs_kind": 0, "sampler": 1, "num_samples": res * res, "grid_res": res, "view_min": -1.0,
                "view_max": 2.0, "num_phase_bins": -1}
    return {"obs_kind": 0, "sampler": 0, "num_samples": G, "grid_res": 0, "view_min": -0.5, "view_max": 10.0,
            "num_phase_bins": -1}
