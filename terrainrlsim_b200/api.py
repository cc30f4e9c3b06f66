"""Host-side mirror of the reference interface over the C ABI (include/trl_b200.h).

Class / method names follow the reference so the parity tests read like the reference's own usage:
  CharModel            <-> cCharacter::Init + cKinTree::Load + cPDController::LoadParams + cMotion::Load
  Batch.imp_pd_update_control_force  <-> cImpPDController::UpdateControlForce (sim/ImpPDController.cpp:48-74)
  Batch.kin_tree_fk                  <-> cKinTree::JointWorldTrans / CalcBodyPartPos (anim/KinTree.cpp:1099-1112,299-308)
  Batch.kin_char_pose                <-> cKinCharacter::Pose (anim/KinCharacter.cpp:111-123)
  Batch.ground_sample_height         <-> cGround::SampleHeight (sim/Ground.h:93-97)
  Batch.build_poli_state             <-> cTerrainRLCharController::ParseGround + BuildPoliState
Arrays may be numpy (HOST pointer mode: staged through the library) or torch CUDA tensors (DEVICE pointer mode:
zero-copy, asynchronous on the batch's stream). PyTorch is plumbing only (device memory, streams).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import ObsCfg, StepArgs, StepOpts, TerrainCfg, TrlError, check
from . import synth

PTR_HOST, PTR_DEVICE = 0, 1
GROUND_NONE, GROUND_VAR2D, GROUND_VAR3D, GROUND_PLANE = 0, 1, 2, 3


def _is_torch(x):
    return x is not None and type(x).__module__.startswith("torch")


def make_obs_cfg(obs_kind=0, sampler=0, num_samples=200, grid_res=0, view_min=-0.5, view_max=10.0, num_phase_bins=-1):
    c = ObsCfg()
    c.obs_kind, c.sampler, c.num_samples, c.grid_res = obs_kind, sampler, num_samples, grid_res
    c.view_min, c.view_max, c.num_phase_bins = view_min, view_max, num_phase_bins
    return c


class CharModel:
    def __init__(self, handle, meta=None):
        if not handle:
            raise TrlError("model creation failed: " + _capi.last_error())
        self._h = C.c_void_p(handle)
        self.meta = meta or {}
        N, n, F = C.c_int(), C.c_int(), C.c_int()
        check(_capi.lib().trl_model_dims(self._h, C.byref(N), C.byref(n), C.byref(F)), "trl_model_dims")
        self.num_joints, self.num_dof, self.num_frames = N.value, n.value, F.value

    @classmethod
    def load(cls, character_file, motion_file=None, gravity=None):
        g = (C.c_double * 3)(*gravity) if gravity is not None else None
        h = _capi.lib().trl_model_load(character_file.encode(), motion_file.encode() if motion_file else None, g)
        return cls(h)

    @classmethod
    def from_dict(cls, d, gravity=None, with_motion=True):
        jm = np.ascontiguousarray(d["joint_mat"], dtype=np.float64)
        bd = np.ascontiguousarray(d["body_defs"], dtype=np.float64)
        pd = np.ascontiguousarray(d["pd_params"], dtype=np.float64)
        frames, nf, loop = None, 0, 0
        if with_motion and "motion" in d:
            frames = np.ascontiguousarray(d["motion"]["frames"], dtype=np.float64)
            nf, loop = frames.shape[0], int(d["motion"]["loop"])
        g = (C.c_double * 3)(*gravity) if gravity is not None else None
        h = _capi.lib().trl_model_create(jm.shape[0], jm.ctypes.data, bd.ctypes.data, pd.ctypes.data, nf,
                                         frames.ctypes.data if frames is not None else None, loop, g)
        return cls(h, d)

    @classmethod
    def from_name(cls, name, **kw):
        return cls.from_dict(synth.load_model_dict(name), **kw)

    def matrices(self):
        N = self.num_joints
        jm, bd, pd = np.zeros((N, 19)), np.zeros((N, 17)), np.zeros((N, 18))
        check(_capi.lib().trl_model_get_matrices(self._h, jm.ctypes.data, bd.ctypes.data, pd.ctypes.data), "get_matrices")
        return jm, bd, pd

    def motion(self):
        fr = np.zeros((self.num_frames, self.num_dof + 1))
        loop = C.c_int()
        check(_capi.lib().trl_model_get_motion(self._h, fr.ctypes.data, C.byref(loop)), "get_motion")
        return fr, bool(loop.value)

    def read_state(self, state_file, pose=None, vel=None):
        """cCharacter::ReadState: (pose [n], vel [n]); keys the file lacks keep the given defaults (zeros if none)."""
        pose = np.zeros(self.num_dof) if pose is None else np.array(pose, dtype=np.float64)
        vel = np.zeros(self.num_dof) if vel is None else np.array(vel, dtype=np.float64)
        check(_capi.lib().trl_state_read(self._h, str(state_file).encode(), pose.ctypes.data, vel.ctypes.data), "trl_state_read")
        return pose, vel

    def write_state(self, state_file, pose, vel):
        """cCharacter::WriteState (identity root transform)."""
        pose, vel = np.ascontiguousarray(pose, dtype=np.float64), np.ascontiguousarray(vel, dtype=np.float64)
        check(_capi.lib().trl_state_write(self._h, str(state_file).encode(), pose.ctypes.data, vel.ctypes.data), "trl_state_write")

    def close(self):
        if self._h:
            _capi.lib().trl_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """E independent characters of one model on one CUDA device."""

    def __init__(self, model, num_envs, device=0, obs_cfg=None):
        self.model = model
        self.E = num_envs
        self.device = device
        self.cfg = obs_cfg if obs_cfg is not None else make_obs_cfg()
        h = _capi.lib().trl_batch_create(model._h, num_envs, device, C.byref(self.cfg))
        if not h:
            raise TrlError("trl_batch_create failed: " + _capi.last_error())
        self._h = C.c_void_p(h)
        self.F = _capi.lib().trl_poli_state_size(self._h)
        self._mode = PTR_HOST
        self._keep = []

    # ---- plumbing ----
    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().trl_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        """cuda_stream_ptr: a cudaStream_t handle (0 = CUDA's legacy default stream); None = the batch's own stream."""
        h = C.c_void_p(0xFFFFFFFFFFFFFFFF) if cuda_stream_ptr is None else C.c_void_p(int(cuda_stream_ptr))
        check(_capi.lib().trl_batch_set_stream(self._h, h), "set_stream")

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        check(_capi.lib().trl_batch_sync(self._h), "trl_batch_sync")

    def launch_count(self):
        return int(_capi.lib().trl_batch_launch_count(self._h))

    def _set_mode(self, mode):
        if mode != self._mode:
            check(_capi.lib().trl_batch_set_pointer_mode(self._h, mode), "set_pointer_mode")
            self._mode = mode

    def _prep(self, arrays):
        """arrays: list of (obj, dtype). Returns raw pointers; picks the pointer mode from the array kind."""
        any_torch = any(_is_torch(a) for a, _ in arrays)
        self._set_mode(PTR_DEVICE if any_torch else PTR_HOST)
        ptrs = []
        self._keep = []
        for a, dt in arrays:
            if a is None:
                ptrs.append(None)
            elif _is_torch(a):
                if not a.is_cuda or not a.is_contiguous():
                    raise TrlError("torch arguments must be contiguous CUDA tensors")
                ptrs.append(C.c_void_p(a.data_ptr()))
            else:
                if any_torch:
                    raise TrlError("cannot mix numpy and torch CUDA arguments in one call")
                if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
                    raise TrlError("numpy arguments must be C-contiguous %s arrays (outputs are written in place)" % dt)
                self._keep.append(a)
                ptrs.append(C.c_void_p(a.ctypes.data))
        return ptrs

    def _new(self, shape, like, dtype=np.float64):
        if _is_torch(like):
            import torch
            tdt = {np.float64: torch.float64, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
            return torch.empty(shape, dtype=tdt, device=like.device)
        return np.empty(shape, dtype=dtype)

    # ---- (1) stable PD ----
    def imp_pd_update_control_force(self, dt, pose, vel, tar_pose, tar_vel, joint_active=None, kp=None, kd=None,
                                    tau=None, want_mass=False, want_bias=False):
        n = self.model.num_dof
        if tau is None:
            tau = self._new((self.E, n), pose)
            if _is_torch(tau):
                tau.zero_()
            else:
                tau[...] = 0
        mass = self._new((self.E, n, n), pose) if want_mass else None
        bias = self._new((self.E, n), pose) if want_bias else None
        f8, u1 = np.float64, np.uint8
        p = self._prep([(pose, f8), (vel, f8), (tar_pose, f8), (tar_vel, f8), (joint_active, u1), (kp, f8), (kd, f8),
                        (tau, f8), (mass, f8), (bias, f8)])
        check(_capi.lib().trl_imp_pd_update_control_force(self._h, float(dt), *p), "trl_imp_pd_update_control_force")
        return tau, mass, bias

    def exp_pd_update_control_force(self, dt, pose, vel, tar_pose, tar_vel=None, joint_active=None, kp=None, kd=None,
                                    tau=None):
        """cExpPDController::UpdateControlForce: explicit PD torques, accumulated into tau [E][n]."""
        n = self.model.num_dof
        if tau is None:
            tau = self._new((self.E, n), pose)
            if _is_torch(tau):
                tau.zero_()
            else:
                tau[...] = 0
        f8, u1 = np.float64, np.uint8
        p = self._prep([(pose, f8), (vel, f8), (tar_pose, f8), (tar_vel, f8), (joint_active, u1), (kp, f8), (kd, f8), (tau, f8)])
        check(_capi.lib().trl_exp_pd_update_control_force(self._h, float(dt), *p), "trl_exp_pd_update_control_force")
        return tau

    def status(self):
        out = np.zeros(self.E, dtype=np.int32)
        self._set_mode(PTR_HOST)
        check(_capi.lib().trl_batch_get_status(self._h, out.ctypes.data), "trl_batch_get_status")
        return out

    # ---- (2) kinematics ----
    def kin_tree_fk(self, pose, want_joint_world=True, want_body_pos=True):
        N = self.model.num_joints
        jw = self._new((self.E, N, 12), pose) if want_joint_world else None
        bp = self._new((self.E, N, 3), pose) if want_body_pos else None
        f8 = np.float64
        p = self._prep([(pose, f8), (jw, f8), (bp, f8)])
        check(_capi.lib().trl_kin_tree_fk(self._h, *p), "trl_kin_tree_fk")
        return jw, bp

    def apply_control_forces(self, pose, tau):
        """cSimCharacter::ApplyControlForces + cJoint::ApplyTau -> clamped tau [E][n], body wrenches [E][N][6] (world)"""
        n, N = self.model.num_dof, self.model.num_joints
        out_tau = self._new((self.E, n), pose)
        wr = self._new((self.E, N, 6), pose)
        f8 = np.float64
        p = self._prep([(pose, f8), (tau, f8), (out_tau, f8), (wr, f8)])
        check(_capi.lib().trl_apply_control_forces(self._h, *p), "trl_apply_control_forces")
        return out_tau, wr

    def rbd_extras(self, pose, want_J=True, want_Jcom=True, want_gravity_force=True):
        """cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce -> J [E][6][n], Jcom [E][6][n], tau_g [E][n]"""
        n = self.model.num_dof
        J = self._new((self.E, 6, n), pose) if want_J else None
        Jc = self._new((self.E, 6, n), pose) if want_Jcom else None
        g = self._new((self.E, n), pose) if want_gravity_force else None
        f8 = np.float64
        p = self._prep([(pose, f8), (J, f8), (Jc, f8), (g, f8)])
        check(_capi.lib().trl_rbd_extras(self._h, *p), "trl_rbd_extras")
        return J, Jc, g

    def rbd_end_effector_jacobian(self, pose, joint_id):
        """cRBDUtil::BuildEndEffectorJacobian -> [E][6][n]; joint_id: an int (every env) or an int32 array / tensor [E]."""
        n = self.model.num_dof
        J = self._new((self.E, 6, n), pose)
        if isinstance(joint_id, (int, np.integer)):
            p = self._prep([(pose, np.float64), (J, np.float64)])
            check(_capi.lib().trl_rbd_end_effector_jacobian(self._h, p[0], None, int(joint_id), p[1]),
                  "trl_rbd_end_effector_jacobian")
        else:
            p = self._prep([(pose, np.float64), (joint_id, np.int32), (J, np.float64)])
            check(_capi.lib().trl_rbd_end_effector_jacobian(self._h, p[0], p[1], -1, p[2]), "trl_rbd_end_effector_jacobian")
        return J

    def kin_char_pose(self, time, origin_pos=None, origin_rot=None, want_vel=True):
        n = self.model.num_dof
        pose = self._new((self.E, n), time)
        vel = self._new((self.E, n), time) if want_vel else None
        f8 = np.float64
        p = self._prep([(time, f8), (origin_pos, f8), (origin_rot, f8), (pose, f8), (vel, f8)])
        check(_capi.lib().trl_kin_char_pose(self._h, *p), "trl_kin_char_pose")
        return pose, vel

    # ---- (3) ground ----
    def ground_set_heightfields(self, kind, fields, env_to_field=None):
        """fields: list (one per distinct terrain) of lists of pieces {data, res, origin, scale, bounds}."""
        if kind in (GROUND_NONE, GROUND_PLANE):
            check(_capi.lib().trl_ground_set_heightfields(self._h, kind, 0, 0, None, 0, None, None, None, None, None, None),
                  "trl_ground_set_heightfields")
            return
        nf, pp = len(fields), len(fields[0])
        offs = np.zeros((nf, pp), dtype=np.int64)
        res = np.zeros((nf, pp, 2), dtype=np.int32)
        origin = np.zeros((nf, pp, 3))
        scale = np.zeros((nf, pp, 3))
        bounds = np.zeros((nf, pp, 4))
        total = 0
        for f, pieces in enumerate(fields):
            assert len(pieces) == pp
            for s, pc in enumerate(pieces):
                offs[f, s] = total
                total += int(np.asarray(pc["data"]).size)
                res[f, s] = pc["res"]
                origin[f, s] = pc["origin"]
                scale[f, s] = pc["scale"]
                bounds[f, s] = pc["bounds"]
        data = np.empty(total, dtype=np.float32)
        for f, pieces in enumerate(fields):
            for s, pc in enumerate(pieces):
                a = np.asarray(pc["data"], dtype=np.float32).ravel()
                data[offs[f, s]:offs[f, s] + a.size] = a
        e2f = None
        if env_to_field is not None:
            e2f = np.ascontiguousarray(env_to_field, dtype=np.int32)
        check(_capi.lib().trl_ground_set_heightfields(
            self._h, kind, nf, pp, data.ctypes.data, total, offs.ctypes.data, res.ctypes.data, origin.ctypes.data,
            scale.ctypes.data, bounds.ctypes.data, e2f.ctypes.data if e2f is not None else None),
            "trl_ground_set_heightfields")

    # ---- (3b) procedural terrain on the device ----
    def ground_generate(self, cfg, num_fields, seeds=None, bound_min=None, bound_max=None, env_to_field=None):
        """cGround::Init for `num_fields` terrains, generated on the device. cfg: TerrainCfg (terrain_cfg / terrain_load);
        seeds [F] uint64; bound_min / bound_max [F][3] view bounds (default: the reference's +-GroundWidth/2 around x = -1
        for var2d, +-10 for var3d, like cGroundFactory)."""
        F = int(num_fields)
        seeds = np.ascontiguousarray(seeds if seeds is not None else np.zeros(F), dtype=np.uint64)
        if bound_min is None or bound_max is None:
            half = 0.5 * (20.0 if cfg.type == 15 else cfg.ground_width)
            shift = 0.0 if cfg.type == 15 else -1.0
            bound_min = np.tile(np.array([-half + shift, 0.0, -half if cfg.type == 15 else 0.0]), (F, 1))
            bound_max = np.tile(np.array([half + shift, 0.0, half if cfg.type == 15 else 0.0]), (F, 1))
        bmin = np.ascontiguousarray(bound_min, dtype=np.float64).reshape(F, 3)
        bmax = np.ascontiguousarray(bound_max, dtype=np.float64).reshape(F, 3)
        e2f = np.ascontiguousarray(env_to_field, dtype=np.int32) if env_to_field is not None else None
        check(_capi.lib().trl_ground_generate(self._h, C.byref(cfg), F, seeds.ctypes.data, bmin.ctypes.data, bmax.ctypes.data,
                                              e2f.ctypes.data if e2f is not None else None), "trl_ground_generate")
        self.num_fields = F

    def ground_update(self, bound_min, bound_max, want_changed=True):
        """cGround::Update for every terrain; returns the per-terrain "rebuilt" flags [F] (numpy or torch like the bounds)."""
        F = self.num_fields
        changed = self._new((F,), bound_min, np.int32) if want_changed else None
        p = self._prep([(bound_min, np.float64), (bound_max, np.float64), (changed, np.int32)])
        check(_capi.lib().trl_ground_update(self._h, *p), "trl_ground_update")
        return changed

    def ground_read_field(self, field):
        """Terrain `field` as a list of pieces {data [res_z][res_x] float32, res, origin, scale, bounds} in scan order."""
        P = 4
        offs = np.zeros(P, dtype=np.int64)
        res = np.zeros((P, 2), dtype=np.int32)
        origin, scale, bounds = np.zeros((P, 3)), np.zeros((P, 3)), np.zeros((P, 4))
        lib = _capi.lib()
        total = lib.trl_ground_read_field(self._h, int(field), None, 0, offs.ctypes.data, res.ctypes.data, origin.ctypes.data,
                                          scale.ctypes.data, bounds.ctypes.data)
        if total < 0:
            check(int(total), "trl_ground_read_field")
        data = np.zeros(int(total), dtype=np.float32)
        total = lib.trl_ground_read_field(self._h, int(field), data.ctypes.data, int(total), offs.ctypes.data, res.ctypes.data,
                                          origin.ctypes.data, scale.ctypes.data, bounds.ctypes.data)
        if total < 0:
            check(int(total), "trl_ground_read_field")
        out = []
        for s in range(P):
            rx, rz = int(res[s, 0]), int(res[s, 1])
            if rx == 0:
                break
            out.append({"data": data[offs[s]:offs[s] + rx * rz].reshape(rz, rx).copy(), "res": (rx, rz),
                        "origin": origin[s].copy(), "scale": scale[s].copy(), "bounds": bounds[s].copy()})
        return out

    def ground_sample_height(self, pos):
        S = pos.shape[1]
        h = self._new((self.E, S), pos)
        valid = self._new((self.E, S), pos, np.uint8)
        cell = self._new((self.E, S, 3), pos, np.int32)
        p = self._prep([(pos, np.float64), (h, np.float64), (valid, np.uint8), (cell, np.int32)])
        check(_capi.lib().trl_ground_sample_height(self._h, p[0], S, p[1], p[2], p[3]), "trl_ground_sample_height")
        return h, valid, cell

    # ---- (4) observation ----
    def build_poli_state(self, pose, body_pos, body_vel, vel=None, body_quat=None, body_angvel=None, ground_vel=None,
                         phase=None, flip=None, want_cells=False):
        state = self._new((self.E, self.F), pose)
        cells = self._new((self.E, max(1, self.cfg.num_samples), 3), pose, np.int32) if want_cells else None
        f8, u1 = np.float64, np.uint8
        p = self._prep([(pose, f8), (vel, f8), (body_pos, f8), (body_vel, f8), (body_quat, f8), (body_angvel, f8),
                        (ground_vel, f8), (phase, f8), (flip, u1), (state, f8), (cells, np.int32)])
        check(_capi.lib().trl_build_poli_state(self._h, *p), "trl_build_poli_state")
        return (state, cells) if want_cells else state

    # ---- imitation reward (cScenarioExpImitate::CalcReward) ----
    def imitate_calc_reward(self, pose, vel, kin_pose, kin_vel, body_vel, fallen=None, want_terms=False):
        reward = self._new((self.E,), pose)
        terms = self._new((self.E, 5), pose) if want_terms else None
        f8, u1 = np.float64, np.uint8
        p = self._prep([(pose, f8), (vel, f8), (kin_pose, f8), (kin_vel, f8), (body_vel, f8), (fallen, u1), (reward, f8),
                        (terms, f8)])
        check(_capi.lib().trl_imitate_calc_reward(self._h, *p), "trl_imitate_calc_reward")
        return (reward, terms) if want_terms else reward

    # ---- fused step ----
    def step_fused(self, dt, inp, out=None):
        """inp: dict with pose, vel, tar_pose, tar_vel, joint_active, kin_time, origin_pos, origin_rot, body_pos,
        body_vel, phase, flip (missing keys = NULL). out: dict of preallocated outputs (tau, kin_pose, kin_vel,
        kin_body_pos, state) or None to allocate all. Returns the out dict."""
        n, N = self.model.num_dof, self.model.num_joints
        like = inp["pose"]
        if out is None:
            out = {"tau": self._new((self.E, n), like), "state": self._new((self.E, self.F), like)}
            if self.model.num_frames > 1 and inp.get("kin_time") is not None:
                out.update({"kin_pose": self._new((self.E, n), like), "kin_vel": self._new((self.E, n), like),
                            "kin_body_pos": self._new((self.E, N, 3), like)})
        f8, u1 = np.float64, np.uint8
        names_in = [("pose", f8), ("vel", f8), ("tar_pose", f8), ("tar_vel", f8), ("joint_active", u1), ("kin_time", f8),
                    ("origin_pos", f8), ("origin_rot", f8), ("body_pos", f8), ("body_vel", f8), ("phase", f8), ("flip", u1)]
        names_out = [("tau", f8), ("kin_pose", f8), ("kin_vel", f8), ("kin_body_pos", f8), ("state", f8)]
        p = self._prep([(inp.get(k), d) for k, d in names_in] + [(out.get(k), d) for k, d in names_out])
        a = StepArgs()
        a.dt = float(dt)
        for (k, _), ptr in zip(names_in, p[:len(names_in)]):
            setattr(a, k, ptr)
        for (k, _), ptr in zip(names_out, p[len(names_in):]):
            setattr(a, "out_" + k, ptr)
        check(_capi.lib().trl_step_fused(self._h, C.byref(a)), "trl_step_fused")
        return out

    def make_step_args(self, dt, inp, out):
        """Pre-resolved trl_step_args for a hot loop over fixed device tensors (DEVICE pointer mode)."""
        f8, u1 = np.float64, np.uint8
        names_in = [("pose", f8), ("vel", f8), ("tar_pose", f8), ("tar_vel", f8), ("joint_active", u1), ("kin_time", f8),
                    ("origin_pos", f8), ("origin_rot", f8), ("body_pos", f8), ("body_vel", f8), ("phase", f8), ("flip", u1)]
        names_out = [("tau", f8), ("kin_pose", f8), ("kin_vel", f8), ("kin_body_pos", f8), ("state", f8)]
        p = self._prep([(inp.get(k), d) for k, d in names_in] + [(out.get(k), d) for k, d in names_out])
        a = StepArgs()
        a.dt = float(dt)
        for (k, _), ptr in zip(names_in, p[:len(names_in)]):
            setattr(a, k, ptr)
        for (k, _), ptr in zip(names_out, p[len(names_in):]):
            setattr(a, "out_" + k, ptr)
        return a

    def set_host_async(self, on):
        """HOST pointer mode: step_fused_args returns after enqueueing; sync() completes the step (pinned buffers only)."""
        check(_capi.lib().trl_batch_set_host_async(self._h, 1 if on else 0), "trl_batch_set_host_async")

    def step_fused_args(self, args, opts=None):
        if opts is None:
            return _capi.lib().trl_step_fused(self._h, C.byref(args))
        return _capi.lib().trl_step_fused_ex(self._h, C.byref(args), C.byref(opts))

    def make_step_opts(self, reward=None, fallen=None, state_f32=None):
        """trl_step_opts for step_fused_args: imitation-reward epilogue [E], fallen flags [E], fp32 state [E][F]."""
        arrs = [(reward, np.float64), (fallen, np.uint8), (state_f32, np.float32)]
        o = StepOpts()
        if all(a is None for a, _ in arrs):
            return o
        keep = self._keep
        o.out_reward, o.fallen, o.out_state_f32 = self._prep(arrs)
        self._keep = keep + self._keep
        return o

    def step_fused_ex(self, dt, inp, out, reward=None, fallen=None, state_f32=None):
        """trl_step_fused_ex: `out` holds only the outputs the caller wants back (missing keys = NULL); reward [E] and
        state_f32 [E][F] are preallocated arrays / tensors or None."""
        a = self.make_step_args(dt, inp, out)
        o = self.make_step_opts(reward, fallen, state_f32)
        check(self.step_fused_args(a, o), "trl_step_fused_ex")
        return out


def terrain_cfg(type_id=1, params=None, ground_width=20.0, world_scale=4.0, spacing_x=0.2, spacing_z=0.2):
    """trl_terrain_cfg with the reference's defaults (cGround::tParams, cTerrainGen2D::gParamDefs)."""
    cfg = TerrainCfg()
    check(_capi.lib().trl_terrain_default_cfg(C.byref(cfg)), "trl_terrain_default_cfg")
    cfg.type = int(type_id)
    if params is not None:
        for i, v in enumerate(params):
            cfg.params[i] = float(v)
    cfg.ground_width, cfg.world_scale, cfg.spacing_x, cfg.spacing_z = ground_width, world_scale, spacing_x, spacing_z
    return cfg


def terrain_load(path, blend=0.0, world_scale=4.0):
    """A data/terrain/*.txt parameter file -> TerrainCfg (cGroundFactory::ParseParamsJson + cGround::SetParamBlend)."""
    cfg = TerrainCfg()
    check(_capi.lib().trl_terrain_load(str(path).encode(), float(blend), C.byref(cfg)), "trl_terrain_load")
    cfg.world_scale = world_scale
    return cfg
