"""cSimAdapter-shaped batched front end (simAdapter/SimAdapter.h:29-81, simAdapter/terrainRLSim.py:28-163).

The reference's SWIG class drives ONE scenario (Bullet world + character + ground + controller). Here the same
method names drive E characters at once; every `std::vector<double>` of the reference becomes an `[E, ...]` array.
What the adapter owns is exactly this repo's scope: the controller's numerical core (stable PD), the kinematic
reference character, the terrain features and the packed policy state. The rigid-body world itself (Bullet in the
reference) stays external: it is a pluggable `world` object with three methods

    world.reset(adapter)                      -> None      (cScenarioSimChar::Reset)
    world.state() -> pose, vel, body_pos, body_vel         (cSimCharacter::BuildPose/BuildVel, cSimObj::GetPos/...)
    world.apply(tau, dt)                      -> None      (cSimCharacter::ApplyControlForces + cWorld::Update)
    world.has_fallen() -> [E] bool                          (cSimCharacter::HasFallen; contact queries live in Bullet)

`KinematicWorld` below is a stand-in used by the tests and the smoke demo: it moves every character along its
mocap reference (so all kernels see consistent, moving inputs) and ignores torques.

Method map (reference -> here):
    init()                 arg parsing + scenario/character/controller construction      SimAdapter.cpp:40-253
    initEpoch()            scenario Reset                                                :255-260
    update()               one animation frame = num_update_steps control substeps       :262-268, ScenarioSimChar.cpp:274-325
    updateAction(a)        PD targets of the actuated dofs (cCtPDController)             :441-456, CtPDController.cpp:70-100
    getState()             ParseGround + BuildPoliState                                  :493-508
    calcReward()           imitation reward vs the kinematic reference                   ScenarioExpImitate.cpp:13-159
    agentHasFallen()       delegated to the world                                        :458-465
    needUpdatedAction()    true on the frames where the controller queries the policy    :467-475
    getActionSpaceSize() / getObservationSpaceSize()                                     :477-491
    setRandomSeed(s)                                                                     :510-515
"""
import ctypes as C
import os

import numpy as np

from . import _capi, synth
from .api import Batch, CharModel, TrlError, make_obs_cfg

ANIM_STEP = 1.0 / 30.0  # gAnimStep, Main.h:76-77


def _arg(path, key, default=None):
    buf = C.create_string_buffer(1024)
    rc = _capi.lib().trl_arg_file_get(path.encode(), key.encode(), buf, 1024)
    return buf.value.decode() if rc == 0 else default


class KinematicWorld:
    """Stand-in for the physics world: the simulated character follows its own mocap clip (pose, finite-difference
    velocity, FK body positions), all computed by this library's kernels. Torques are ignored."""

    def __init__(self, seed=0):
        self.seed = seed

    def reset(self, adapter):
        self.a = adapter
        rng = np.random.default_rng(self.seed)
        E = adapter.num_envs
        self.time = rng.uniform(0.0, 1.0, E)
        self.origin = np.zeros((E, 3))
        self.origin[:, 0] = rng.uniform(-6.0, 6.0, E)
        self._refresh(0.0)

    def _refresh(self, dt):
        b = self.a.batch
        self.time = self.time + dt
        self.pose, self.vel = b.kin_char_pose(np.ascontiguousarray(self.time), self.origin, None)
        _, self.body_pos = b.kin_tree_fk(self.pose, want_joint_world=False)
        if dt > 0 and getattr(self, "_prev_bp", None) is not None:
            self.body_vel = (self.body_pos - self._prev_bp) / dt
        else:
            self.body_vel = np.zeros_like(self.body_pos)
        self._prev_bp = self.body_pos.copy()

    def state(self):
        return self.pose, self.vel, self.body_pos, self.body_vel

    def apply(self, tau, dt):
        self._refresh(dt)

    def has_fallen(self):
        return self.pose[:, 1] < 0.3


class BatchedSimAdapter:
    def __init__(self, args, num_envs, device=0, world=None):
        """args: the reference's argv list (e.g. ['-arg_file=', 'args/test_dog_args.txt']) or the name of one of
        the bundled configs ('biped2d', 'hopper', 'dog', 'raptor', 'biped3d')."""
        self.args = args
        self.num_envs = int(num_envs)
        self.device = device
        self.world = world if world is not None else KinematicWorld()
        self.batch = None
        self.seed = 0
        self._frame = 0

    @staticmethod
    def _data_root(arg_file):
        """The reference resolves data/... paths against its working directory (the repository root): the first ancestor
        of the arg file that has a data/ directory."""
        d = os.path.dirname(os.path.abspath(arg_file))
        while d and d != os.path.dirname(d):
            if os.path.isdir(os.path.join(d, "data")):
                return d
            d = os.path.dirname(d)
        return os.getcwd()

    # ---- construction ----
    def init(self):
        cfg_name = None
        if isinstance(self.args, str):
            cfg_name = self.args
            meta = synth.load_model_dict(cfg_name)
            self.model = CharModel.from_dict(meta)
            self.num_update_steps = int(meta["num_update_steps"])
            self.obs_cfg = make_obs_cfg(**synth.obs_config(meta))
            self._terrain_meta = meta
            self._terrain_file = None
        else:
            arg_file = rel = None
            a = list(self.args)
            for i, tok in enumerate(a):
                if tok.startswith("-arg_file=") and i + 1 < len(a):
                    arg_file = a[i + 1]
                if tok.startswith("-relative_file_path=") and i + 1 < len(a):
                    rel = a[i + 1]  # prefix of every data path (SimAdapter.cpp:104-118, cCharacter::setRelativeFilePath)
            if arg_file is None:
                raise TrlError("BatchedSimAdapter.init: no -arg_file= given")
            root = rel if rel is not None else (self._data_root(arg_file) + "/")

            def path(p):
                if not p or os.path.isabs(p) or os.path.exists(p):
                    return p
                return os.path.join(root, p)

            char_file = path(_arg(arg_file, "character_file"))
            motion_file = path(_arg(arg_file, "motion_file"))
            if char_file is None:
                raise TrlError("arg file has no -character_file=")
            self.model = CharModel.load(char_file, motion_file)
            self.state_file = path(_arg(arg_file, "state_file"))
            self.num_update_steps = int(_arg(arg_file, "num_update_steps", "20"))
            ctrl = _arg(arg_file, "char_ctrl", "")
            if ctrl.startswith("ct"):
                self.obs_cfg = make_obs_cfg(obs_kind=2, num_samples=0, num_phase_bins=4 if "phase" in ctrl else -1)
            elif ctrl.startswith("biped3D"):
                self.obs_cfg = make_obs_cfg(obs_kind=0, sampler=1, num_samples=256, grid_res=16, view_min=-1.0, view_max=2.0)
            else:
                self.obs_cfg = make_obs_cfg(obs_kind=1 if "hopper" in ctrl else 0)
            terrain = _arg(arg_file, "terrain_file", "") or ""
            self._terrain_meta = None
            self._terrain_file = path(terrain) if terrain else None
            if self._terrain_file is None:  # no -terrain_file=: cGround::tParams defaults (flat 2-D ground)
                self._terrain_meta = {"name": ctrl, "terrain_cfg": {"type": 1, "ground_width": 20.0, "spacing_x": 0.2,
                                                                     "spacing_z": 0.2, "params": None}}
            self._world_scale = float(_arg(arg_file, "world_scale", "1"))
        self.batch = Batch(self.model, self.num_envs, self.device, self.obs_cfg)
        self.n = self.model.num_dof
        self.N = self.model.num_joints
        self.dt = ANIM_STEP / self.num_update_steps
        self._build_ground()
        E = self.num_envs
        self.tar_pose = np.zeros((E, self.n))
        self.tar_vel = np.zeros((E, self.n))
        self.tau = np.zeros((E, self.n))
        self.initEpoch()

    def _build_ground(self):
        """cScenarioSimChar::BuildGround: one procedural terrain per env (up to 4096), generated on the device from the
        config's terrain parameter file with the reference's generator (trl_ground_generate)."""
        from .api import terrain_cfg, terrain_load
        F = min(self.num_envs, 4096)
        if self._terrain_file is not None:
            cfg = terrain_load(self._terrain_file, world_scale=self._world_scale)
            seeds = (np.arange(F, dtype=np.uint64) * np.uint64(104729) + np.uint64(self.seed) + np.uint64(1)) % np.uint64(2147483647)
        else:
            t = synth.terrain_setup(self._terrain_meta, F, seed=0xB200 + self.seed)
            cfg = terrain_cfg(t["type"], t["params"], t["ground_width"], t["world_scale"], t["spacing_x"], t["spacing_z"])
            seeds = t["seeds"]
        self.batch.ground_generate(cfg, F, seeds=seeds)
        self._ground_fields = F

    def updateGround(self, root_pos, view_half=10.0):
        """cScenarioSimChar::UpdateGround -> cGround::Update: scroll the terrains under the characters. root_pos [E][3];
        terrain f follows env f (envs beyond the terrain count share terrains and do not move them)."""
        F = self._ground_fields
        p = np.asarray(root_pos, dtype=np.float64)[:F]
        return self.batch.ground_update(np.ascontiguousarray(p - [view_half, 0, view_half]),
                                        np.ascontiguousarray(p + [view_half, 0, view_half]))

    # ---- cSimAdapter methods ----
    def setRandomSeed(self, seed):
        self.seed = int(seed)
        if hasattr(self.world, "seed"):
            self.world.seed = self.seed

    def initEpoch(self):
        self.world.reset(self)
        self._frame = 0
        pose, _, _, _ = self.world.state()
        self.tar_pose[...] = pose  # default action: hold the current pose
        self.tar_pose[:, :7] = 0

    def getActionSpaceSize(self):
        return self.n - 7  # cCtController::GetPoliActionSize (sim/CtController.cpp:82-88)

    def getObservationSpaceSize(self):
        return self.batch.F

    def needUpdatedAction(self):
        return True  # every adapter frame (policy query rate = 30 Hz)

    def updateAction(self, action):
        """action [E][n-7]: PD target pose of the actuated dofs (SetTargetTheta for every joint)."""
        action = np.asarray(action, dtype=np.float64)
        if action.shape != (self.num_envs, self.n - 7):
            raise TrlError("updateAction: expected shape %s" % ((self.num_envs, self.n - 7),))
        self.tar_pose[:, 7:] = action

    def handleUpdatedAction(self):
        """cSimAdapter::handleUpdatedAction: the new action has been consumed by updateAction; nothing is pending."""
        return None

    def update(self):
        """One animation frame: num_update_steps x (read state, stable-PD torques, hand them to the world)."""
        for _ in range(self.num_update_steps):
            pose, vel, _, _ = self.world.state()
            self.tau[...] = 0  # cSimCharacter::ClearJointTorques
            self.batch.imp_pd_update_control_force(self.dt, np.ascontiguousarray(pose), np.ascontiguousarray(vel),
                                                   self.tar_pose, self.tar_vel, tau=self.tau)
            self.world.apply(self.tau, self.dt)
        self._frame += 1

    def getState(self):
        pose, vel, body_pos, body_vel = self.world.state()
        phase = None
        if self.obs_cfg.num_phase_bins >= 0:
            phase = np.ascontiguousarray((self._frame * ANIM_STEP) % 1.0 * np.ones(self.num_envs))
        return self.batch.build_poli_state(np.ascontiguousarray(pose), np.ascontiguousarray(body_pos),
                                           np.ascontiguousarray(body_vel), vel=np.ascontiguousarray(vel), phase=phase)

    def calcReward(self):
        """cScenarioExpImitate::CalcReward against the kinematic reference character at the current time."""
        if self.model.num_frames < 2:
            raise TrlError("calcReward needs a motion (imitation scenarios)")
        pose, vel, _, body_vel = self.world.state()
        t = np.ascontiguousarray(getattr(self.world, "time", np.full(self.num_envs, self._frame * ANIM_STEP)), dtype=np.float64)
        origin = getattr(self.world, "origin", None)
        kin_pose, kin_vel = self.batch.kin_char_pose(t, origin, None)
        fallen = np.ascontiguousarray(self.agentHasFallen(), dtype=np.uint8)
        return self.batch.imitate_calc_reward(np.ascontiguousarray(pose), np.ascontiguousarray(vel), kin_pose, kin_vel,
                                              np.ascontiguousarray(body_vel), fallen)

    def agentHasFallen(self):
        return np.asarray(self.world.has_fallen(), dtype=bool)

    def finish(self):
        if self.batch is not None:
            self.batch.close()
            self.batch = None
