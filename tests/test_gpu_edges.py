"""GPU edge cases the bundled characters never reach (VERDICT r01 "untested edges"), CUDA path vs the oracle through the C ABI:

  * a synthetic character with non-root PLANAR and FIXED joints (and, in one variant, a shape-less body) through every
    entry point -- explicit PD, torque post-processing (planar "torque in slot 0", sim/Joint.cpp:541-546), Jacobians,
    end-effector Jacobian, FK;
  * stable PD exactly at the threshold of cMathUtil::QuaternionToAxisAngle (|v| <= 1e-4, util/MathUtil.cpp:408-414): the
    "hold pose" target and targets a hair below / above the branch;
  * the kinematic character across two identical / antipodal consecutive frames (Eigen slerp's |d| >= 1 - eps branch and the
    d < 0 sign flip).
"""
import numpy as np
import pytest

from helpers import assert_close, synthetic_character, synthetic_state
from terrainrlsim_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trl():
    import terrainrlsim_b200 as t
    return t


@pytest.mark.parametrize("null_body", [False, True])
def test_planar_and_fixed_joints_every_entry_point(trl, oracle, null_body):
    jm, bd, pd = synthetic_character(with_null_body=null_body)
    om = oracle.Model(joint_mat=jm, body_defs=bd, pd_params=pd)
    pm = trl.CharModel.from_dict({"joint_mat": jm, "body_defs": bd, "pd_params": pd})
    E = 45
    pose, vel, tar, tar_vel = synthetic_state(jm, E)
    rng = np.random.default_rng(3)
    tar_vel = rng.normal(size=tar_vel.shape)
    batch = trl.Batch(pm, E, 0)
    dt = 1.0 / 600
    # stable PD with M / C
    tau, M, C = batch.imp_pd_update_control_force(dt, pose, vel, tar, tar_vel, want_mass=True, want_bias=True)
    rt, rM, rC = np.zeros_like(tau), np.zeros_like(M), np.zeros_like(C)
    for e in range(E):
        rt[e], rM[e], rC[e], _ = om.imp_pd(dt, pose[e], vel[e], tar[e], tar_vel[e])
    assert_close(M.reshape(E, -1), rM.reshape(E, -1), "mass matrix (planar / fixed joints)")
    assert_close(C, rC, "bias force (planar / fixed joints)")
    assert_close(tau, rt, "stable-PD torque (planar / fixed joints)")
    # explicit PD: planar and fixed joints add nothing (PDController.cpp:358-364,391-395)
    tau0 = rng.normal(size=tau.shape)
    ex = batch.exp_pd_update_control_force(dt, pose, vel, tar, tar_vel, tau=tau0.copy())
    rex = np.stack([om.exp_pd(dt, pose[e], vel[e], tar[e], tar_vel[e], tau0=tau0[e]) for e in range(E)])
    assert_close(ex, rex, "explicit PD (planar / fixed joints)")
    for j in range(om.N):
        if int(jm[j, 0]) in (1, 3) and jm[j, 1] >= 0:
            o, s = int(om.offset[j]), int(om.size[j])
            assert np.array_equal(ex[:, o:o + s], tau0[:, o:o + s])
    # torque post-processing: clamps, the planar joint's torque slot, world wrenches
    big = tau * 3.0 + rng.normal(size=tau.shape) * 50.0
    ct, wr = batch.apply_control_forces(pose, big)
    rct, rwr = np.zeros_like(ct), np.zeros_like(wr)
    for e in range(E):
        rct[e], rwr[e] = om.apply_control_forces(pose[e], big[e])
    assert_close(ct, rct, "clamped generalized torques (planar / fixed joints)")
    assert_close(wr.reshape(E, -1), rwr.reshape(E, -1), "body wrenches (planar / fixed joints)")
    # Jacobians, gravity force, end-effector Jacobian of every joint, FK
    J, Jc, g = batch.rbd_extras(pose)
    rJ, rJc, rg = np.zeros_like(J), np.zeros_like(Jc), np.zeros_like(g)
    for e in range(E):
        rJ[e], rJc[e], rg[e] = om.rbd_extras(pose[e])
    assert_close(J.reshape(E, -1), rJ.reshape(E, -1), "Jacobian (planar / fixed joints)")
    assert_close(Jc.reshape(E, -1), rJc.reshape(E, -1), "CoM Jacobian (planar / fixed joints)")
    assert_close(g, rg, "gravity force (planar / fixed joints)")
    for jid in range(om.N):
        Je = batch.rbd_end_effector_jacobian(pose, jid)
        ref = np.stack([om.end_eff_jacobian(pose[e], jid) for e in range(E)])
        assert_close(Je.reshape(E, -1), ref.reshape(E, -1), "end-effector Jacobian of joint %d" % jid)
    jw, bp = batch.kin_tree_fk(pose)
    for e in range(0, E, 7):
        for j in range(om.N):
            W = om.joint_world_trans(pose[e], j)
            assert np.abs(jw[e, j].reshape(3, 4) - W[:3, :]).max() <= 1e-9 * max(1.0, np.abs(W).max())
            if bd[j, 0] != 3:
                assert np.abs(bp[e, j] - om.body_part_pos(pose[e], j)[:3]).max() <= 1e-9 * max(1.0, np.abs(bp[e, j]).max())
    assert np.all(batch.status() == 0)


def _rotate(q, axis, angle):
    h = 0.5 * angle
    r = np.concatenate([[np.cos(h)], np.sin(h) * np.asarray(axis) / np.linalg.norm(axis)])
    return synth.quat_mul(q[None], r[None])[0]


@pytest.mark.parametrize("name", ["biped3d", "dog"])
def test_stable_pd_at_the_axis_angle_threshold(trl, oracle, name):
    """targets equal to the integrated pose (|v| = 0), and |v| = sin(theta / 2) at 0.9e-4, 0.999e-4, 1.001e-4 and 1.1e-4
    around QuaternionToAxisAngle's 1e-4 branch, on every spherical joint; zero velocities so pose_inc == pose"""
    om = oracle.Model(name)
    pm = trl.CharModel.from_name(name)
    base = synth.make_inputs(name, 8)
    cases = [0.0, 0.9e-4, 0.999e-4, 1.001e-4, 1.1e-4, 3e-4]
    E = len(cases) * 8
    pose = np.tile(base["pose"], (len(cases), 1))
    vel = np.zeros_like(pose)
    tar = pose.copy()
    tar[:, :7] = 0
    rng = np.random.default_rng(0)
    for c, sv in enumerate(cases):
        for e in range(8):
            row = c * 8 + e
            for o in om.quat_slots():
                if o == 3:
                    continue
                tar[row, o:o + 4] = _rotate(pose[row, o:o + 4], rng.normal(size=3), 2 * np.arcsin(sv))
    batch = trl.Batch(pm, E, 0)
    dt = 1.0 / 600
    tau, _, _ = batch.imp_pd_update_control_force(dt, pose, vel, tar, np.zeros_like(pose))
    ref = np.stack([om.imp_pd(dt, pose[e], vel[e], tar[e], np.zeros(om.n))[0] for e in range(E)])
    assert_close(tau, ref, "stable PD around the 1e-4 axis-angle threshold")
    ex = batch.exp_pd_update_control_force(dt, pose, vel, tar, None)
    rex = np.stack([om.exp_pd(dt, pose[e], vel[e], tar[e], np.zeros(om.n)) for e in range(E)])
    assert_close(ex, rex, "explicit PD around the 1e-4 axis-angle threshold")
    if name == "biped3d":  # below the threshold a (joint-local) spherical target gives no explicit-PD torque at all
        local = [int(om.offset[j]) for j in range(1, om.N) if int(om.joint_mat[j, 0]) == 4 and om.pd_params[j, 17] == 0]
        assert local
        for o in local:
            assert np.array_equal(ex[0:24, o:o + 4], np.zeros((24, 4)))
            assert np.all(np.abs(ex[24:48, o:o + 3]).max(axis=1) > 0)


def test_kin_character_identical_and_antipodal_frames(trl, oracle):
    """two consecutive identical frames (Eigen slerp: |d| >= 1 - eps -> linear weights) and a frame whose quaternions are
    the negatives of the previous one (d < 0: the second weight flips sign), on the root and the spherical joints"""
    d = synth.load_model_dict("biped3d")
    frames = np.array(d["motion"]["frames"], dtype=np.float64)
    dup = frames[2].copy()
    anti = frames[4].copy()
    om0 = oracle.Model("biped3d")
    for o in om0.quat_slots():
        anti[1 + o:1 + o + 4] *= -1.0
    new = np.vstack([frames[:3], dup[None], frames[3:5], anti[None], frames[5:]])
    d2 = dict(d)
    d2["motion"] = {"frames": new.tolist(), "loop": True}
    om = oracle.Model(joint_mat=np.array(d["joint_mat"]), body_defs=np.array(d["body_defs"]), pd_params=np.array(d["pd_params"]),
                      frames=new, loop=True)
    pm = trl.CharModel.from_dict(d2)
    starts = np.concatenate([[0.0], np.cumsum(new[:-1, 0])])
    t = []
    for k in (2, 4):  # inside the duplicated / antipodal intervals, at their ends and just beside them
        a, b = starts[k], starts[k + 2]
        t += list(np.linspace(a, b, 9)) + [a + 1e-9, b - 1e-9]
    t = np.array(t)
    E = t.size
    batch = trl.Batch(pm, E, 0)
    pose, vel = batch.kin_char_pose(t, None, None)
    for e in range(E):
        rp, rv = om.kin_char_pose(float(t[e]))
        assert np.abs(pose[e] - rp).max() <= 1e-9 * max(1.0, np.abs(rp).max()), "kin pose at t = %r" % t[e]
        assert np.abs(vel[e] - rv).max() <= 1e-8 * max(1.0, np.abs(rv).max()), "kin vel at t = %r" % t[e]


def test_policy_state_zero_fills_slots_of_shape_less_bodies(trl, oracle):
    """cTerrainRLCharController::BuildPoliStatePose starts from Zero(psize): with a shape-less body the tail of the pose block
    must be zero, not whatever the output buffer held"""
    from helpers import oracle_grounds
    jm, bd, pd = synthetic_character(with_null_body=True)
    om = oracle.Model(joint_mat=jm, body_defs=bd, pd_params=pd)
    pm = trl.CharModel.from_dict({"joint_mat": jm, "body_defs": bd, "pd_params": pd})
    E = 33
    pose, vel, _, _ = synthetic_state(jm, E)
    rng = np.random.default_rng(2)
    body_pos = pose[:, None, :3] + rng.normal(0, 0.3, (E, om.N, 3))
    body_vel = rng.normal(size=(E, om.N, 3))
    cfg = trl.make_obs_cfg(obs_kind=0, num_samples=200)
    ocfg = oracle.obs_cfg(obs_kind=0, num_samples=200)
    batch = trl.Batch(pm, E, 0, cfg)
    kind, fields = synth.make_grounds("dog", 3)
    batch.ground_set_heightfields(kind, fields)
    grounds = oracle_grounds(oracle, kind, fields)
    import torch
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in
           {"pose": pose, "vel": vel, "bp": body_pos, "bv": body_vel}.items()}
    batch.use_torch_stream()
    state = batch.build_poli_state(dev["pose"], dev["bp"], dev["bv"], vel=dev["vel"])
    state.fill_(123.0)  # stale contents; the call below must overwrite every entry
    p = batch._prep([(dev["pose"], np.float64), (dev["vel"], np.float64), (dev["bp"], np.float64), (dev["bv"], np.float64),
                     (None, np.float64), (None, np.float64), (None, np.float64), (None, np.float64), (None, np.uint8),
                     (state, np.float64), (None, np.int32)])
    from terrainrlsim_b200 import _capi
    assert _capi.lib().trl_build_poli_state(batch._h, *p) == 0
    batch.sync()
    got = state.cpu().numpy()
    batch.set_stream(None)
    for e in range(E):
        ref = oracle.build_poli_state(om, grounds[e % len(grounds)], ocfg, pose[e], vel[e], body_pos[e], body_vel[e])
        assert np.abs(got[e] - ref).max() <= 1e-9 * max(1.0, np.abs(ref[np.isfinite(ref)]).max()), e
    assert not np.any(got == 123.0)
