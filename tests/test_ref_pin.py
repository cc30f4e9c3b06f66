"""Pins the C oracle (oracle/trl_oracle.c) to the REFERENCE ITSELF: the reference's own, unmodified translation
units compiled into oracle/_ref/libtrl_ref.so (oracle/Makefile target `ref`; Eigen / Bullet replaced by the
stand-in headers under oracle/eigen_shim and oracle/ref_stubs).  Same inputs through both, compared at 1e-13
relative (the stand-in Eigen sums sequentially; a vectorised Eigen build may associate reductions differently,
so bits are not promised -- although most of these agree to the last bit).

These tests run wherever the built library exists (it is built in the container that has /root/reference and
travels to the GPU box as a built .so); tests that need the reference's *data files* additionally need the tree.
The same comparisons are frozen as fixtures by tools/make_golden.py -> tests/golden/ (test_golden.py), which run
everywhere.
"""
import os

import numpy as np
import pytest

from conftest import CONFIG_NAMES
from terrainrlsim_b200 import synth

ref_py = pytest.importorskip("oracle.ref_py")
pytestmark = pytest.mark.skipif(not (ref_py.available() or ref_py.have_reference_tree()),
                                reason="oracle/_ref is not built and /root/reference is absent")
needs_tree = pytest.mark.skipif(not ref_py.have_reference_tree(), reason="needs the reference's data files")

TOL = 1e-13
REF = ref_py.REF_ROOT


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1.0)
    return float(np.abs(a - b).max() / scale) if a.size else 0.0


def models(oracle, name, world_coord=False):
    m = oracle.Model(name)
    pd = m.pd_params.copy()
    if not world_coord:
        pd[:, 17] = 0  # UseWorldCoord off: local targets (the world-target conversion has its own test)
        m = oracle.Model(joint_mat=m.joint_mat, body_defs=m.body_defs, pd_params=pd, frames=m.frames_raw, loop=m.loop)
    r = ref_py.RefModel(m.joint_mat, m.body_defs, pd)
    return m, r


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_rbd_model_update_mass_and_bias(oracle, name):
    """cRBDModel::Update -> GetMassMat / GetBiasForce (sim/RBDModel.cpp:39-55, sim/RBDUtil.cpp:113-179,931-935)"""
    m, r = models(oracle, name)
    x = synth.make_inputs(name, 8)
    for e in range(8):
        M0, C0 = m.rbd_update(x["pose"][e], x["vel"][e])
        M1, C1 = r.rbd_update(x["pose"][e], x["vel"][e])
        assert rel(M0, M1) <= TOL and rel(C0, C1) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_inverse_dynamics_and_inertia(oracle, name):
    """cRBDUtil::SolveInvDyna (sim/RBDUtil.cpp:4-87), BuildInertiaSpatialMat (:600-684, capsule quirk Q5)"""
    m, r = models(oracle, name)
    x = synth.make_inputs(name, 4)
    rng = np.random.default_rng(3)
    for e in range(4):
        acc = rng.normal(size=m.n)
        assert rel(m.inv_dyna(x["pose"][e], x["vel"][e], acc), r.inv_dyna(x["pose"][e], x["vel"][e], acc)) <= TOL
    for j in range(m.N):
        if m.body_defs[j, 0] != 3:
            assert rel(m.inertia_spatial(j), r.inertia_spatial(j)) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_jacobians_gravity_com(oracle, name):
    """cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce / BuildEndEffectorJacobian / CalcCoM
    (sim/RBDUtil.cpp:181-345,557-598,937-982)"""
    m, r = models(oracle, name)
    x = synth.make_inputs(name, 4)
    for e in range(4):
        a, b = m.rbd_extras(x["pose"][e]), r.rbd_extras(x["pose"][e])
        for u, v in zip(a, b):
            assert rel(u, v) <= TOL
        for j in range(m.N):
            assert rel(m.end_eff_jacobian(x["pose"][e], j), r.end_eff_jacobian(x["pose"][e], j)) <= TOL
        c0 = oracle.calc_com(m, x["pose"][e], x["vel"][e])
        c1 = r.calc_com(x["pose"][e], x["vel"][e])
        assert rel(c0[0], c1[0]) <= TOL and rel(c0[1], c1[1]) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_kin_tree_transforms_and_pose_algebra(oracle, name):
    """cKinTree::{ChildParentTrans, JointWorldTrans, CalcBodyPartPos, VelToPoseDiff, PoseProcessPose, CalcVel,
    LerpPoses, CalcHeading, BuildOriginTrans} (anim/KinTree.cpp:1055-1139,1470-1639)"""
    m, r = models(oracle, name)
    x = synth.make_inputs(name, 6)
    y = synth.make_inputs(name, 6, seed=77)
    for e in range(6):
        p0, p1, v = x["pose"][e], y["pose"][e], x["vel"][e]
        for j in range(m.N):
            assert rel(m.child_parent_trans(p0, j), r.child_parent_trans(p0, j)) <= TOL
            assert rel(m.joint_world_trans(p0, j), r.joint_world_trans(p0, j)) <= TOL
            if m.body_defs[j, 0] != 3:
                assert rel(m.body_part_pos(p0, j), r.body_part_pos(p0, j)) <= TOL
        assert rel(m.vel_to_pose_diff(p0, v), r.vel_to_pose_diff(p0, v)) <= TOL
        assert rel(m.post_process_pose(p0 * 1.7), r.post_process_pose(p0 * 1.7)) <= TOL
        for dt in (1.0, 1.0 / 60.0):
            assert rel(m.calc_vel(p0, p1, dt), r.calc_vel(p0, p1, dt)) <= TOL
        for t in (0.0, 0.3, 1.0):
            assert rel(m.lerp_poses(p0, p1, t), r.lerp_poses(p0, p1, t)) <= TOL
        assert abs(oracle.lib().trlo_calc_heading(oracle._d(np.ascontiguousarray(p0))) - r.calc_heading(p0)) <= TOL
        out = np.zeros((4, 4))
        oracle.lib().trlo_build_origin_trans(oracle._d(np.ascontiguousarray(p0)), oracle._d(out))
        assert rel(out, r.build_origin_trans(p0)) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_slerp_and_axis_angle_branches(oracle, name):
    """Eigen slerp |d| >= 1-eps and d < 0 branches (anim/KinTree.cpp:1544-1560); QuaternionToAxisAngle threshold
    1e-4 (util/MathUtil.cpp:399-416) through cKinTree::CalcVel"""
    m, r = models(oracle, name)
    x = synth.make_inputs(name, 3)
    for e in range(3):
        p0 = x["pose"][e]
        # identical poses: |d| = 1 -> linear weights; CalcVel of identical quaternions -> sin_theta = 0 <= 1e-4
        assert rel(m.lerp_poses(p0, p0, 0.37), r.lerp_poses(p0, p0, 0.37)) <= TOL
        assert rel(m.calc_vel(p0, p0, 1.0), r.calc_vel(p0, p0, 1.0)) <= TOL
        # antipodal representation of the same rotations: d < 0 -> second weight negated
        p1 = p0.copy()
        for q in m.quat_slots():
            p1[q:q + 4] = -p1[q:q + 4]
        assert rel(m.lerp_poses(p0, p1, 0.6), r.lerp_poses(p0, p1, 0.6)) <= TOL
        # rotations by angles just below / above the 1e-4 sine threshold
        for ang in (1.9e-4, 2.1e-4, 1e-7):
            p2 = p0.copy()
            for q in m.quat_slots():
                dq = np.array([np.cos(ang / 2), np.sin(ang / 2) * 0.6, 0.0, np.sin(ang / 2) * 0.8])
                p2[q:q + 4] = synth.quat_mul(p0[q:q + 4][None], dq[None])[0]
            assert rel(m.calc_vel(p0, p2, 1.0), r.calc_vel(p0, p2, 1.0)) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_stable_pd_torque(oracle, name):
    """cImpPDController::UpdateControlForce (sim/ImpPDController.cpp:48-74,137-196) + cPDController target
    post-processing + Eigen LDLT, through the reference's own controller classes"""
    m, r = models(oracle, name)
    E = 12
    x = synth.make_inputs(name, E)
    rng = np.random.default_rng(11)
    dts = [1.0 / 600, 1.0 / 2400, 1.0 / 30]
    for e in range(E):
        dt = dts[e % 3]
        act = x["joint_active"][e].copy()
        if e % 4 == 3:
            act = (rng.uniform(size=m.N) < 0.6).astype(np.uint8)
        kp = kd = None
        if e % 5 == 4:
            kp, kd = rng.uniform(10, 500, m.N), rng.uniform(1, 50, m.N)
        tau0 = rng.normal(size=m.n) if e % 2 else None
        t0, _, _, _ = m.imp_pd(dt, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act, kp, kd, tau0)
        t1, _, _ = r.imp_pd(dt, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act, kp, kd, tau0)
        assert rel(t0, t1) <= TOL, (name, e)
    # dt <= 0: no-op like the reference (:56)
    t0, _, _, _ = m.imp_pd(0.0, x["pose"][0], x["vel"][0], x["tar_pose"][0], x["tar_vel"][0])
    t1, _, _ = r.imp_pd(0.0, x["pose"][0], x["vel"][0], x["tar_pose"][0], x["tar_vel"][0])
    assert np.all(t0 == 0) and np.all(t1 == 0)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_stable_pd_hold_pose_threshold(oracle, name):
    """target == current pose on every joint (the adapter's default "hold pose" action): pose_err goes through the
    <= 1e-4 branch of QuaternionToAxisAngle for spherical joints"""
    m, r = models(oracle, name)
    x = synth.make_inputs(name, 4)
    for e in range(4):
        tar = x["pose"][e].copy()
        zero_vel = np.zeros(m.n)
        t0, _, _, _ = m.imp_pd(1.0 / 600, x["pose"][e], zero_vel, tar, x["tar_vel"][e])
        t1, _, _ = r.imp_pd(1.0 / 600, x["pose"][e], zero_vel, tar, x["tar_vel"][e])
        assert rel(t0, t1) <= TOL
        t0, _, _, _ = m.imp_pd(1.0 / 600, x["pose"][e], x["vel"][e], tar, x["tar_vel"][e])
        t1, _, _ = r.imp_pd(1.0 / 600, x["pose"][e], x["vel"][e], tar, x["tar_vel"][e])
        assert rel(t0, t1) <= TOL


@needs_tree
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_model_fixtures_equal_reference_loaders(oracle, name):
    """terrainrlsim_b200/data/*.json == cKinTree::Load / LoadBodyDefs / cPDController::LoadParams on the reference's
    character files (anim/KinTree.cpp:127-199,451-499,980-1041; sim/PDController.cpp:29-91)"""
    d = synth.load_model_dict(name)
    jm, bd, pd = ref_py.load_character(os.path.join(REF, d["source"]["character"]))
    m = oracle.Model(name)
    assert np.array_equal(jm, m.joint_mat)
    assert np.array_equal(bd[:, :13], m.body_defs[:, :13])  # colours are not on the path
    assert np.array_equal(pd, m.pd_params)


@needs_tree
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_kin_character_pose_and_motion(oracle, name):
    """cKinCharacter::Pose(t) = cMotion::CalcFrame + LerpPoses + cycle offset + origin, CalcVel at t + 1/60
    (anim/KinCharacter.cpp:111-123,280-325; anim/Motion.cpp:112-145,207-273)"""
    d = synth.load_model_dict(name)
    kc = ref_py.RefKinChar(os.path.join(REF, d["source"]["character"]), os.path.join(REF, d["source"]["motion"]))
    m = oracle.Model(name)
    assert kc.n == m.n and kc.num_frames == m.frames.shape[0] and kc.loop == m.loop
    assert np.array_equal(kc.frames(), m.frames)
    dur = kc.duration
    rng = np.random.default_rng(5)
    times = list(rng.uniform(0, 10 * dur, 12)) + [0.0, -0.3 * dur, -2.2 * dur, dur, 3 * dur, float(m.frames[1, 0]),
                                                 float(m.frames[2, 0]) - 1e-12, 0.5 * dur]
    for i, t in enumerate(times):
        assert kc.index_phase(t)[0] == m.motion_index_phase(t)[0]
        assert abs(kc.index_phase(t)[1] - m.motion_index_phase(t)[1]) <= 1e-12
        op = orr = None
        if i % 3 == 1:
            op = rng.normal(size=3)
        if i % 3 == 2:
            op = rng.normal(size=3)
            ang = rng.uniform(-3, 3)
            orr = np.array([np.cos(ang / 2), 0.0, np.sin(ang / 2), 0.0])
        p0, v0 = m.kin_char_pose(t, op, orr)
        p1, v1 = kc.pose(t, op, orr)
        assert rel(p0, p1) <= TOL, (name, t)
        assert rel(v0, v1) <= 60 * TOL, (name, t)  # finite difference over 1/60 s


@needs_tree
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_state_files(oracle, name):
    """cCharacter::ReadState (anim/Character.cpp:318-342: parse + PoseProcessPose) == the start state in the model
    fixtures (stored as on disk) after the oracle's PoseProcessPose"""
    d = synth.load_model_dict(name)
    kc = ref_py.RefKinChar(os.path.join(REF, d["source"]["character"]))
    pose, vel = kc.read_state(os.path.join(REF, d["source"]["state"]))
    assert np.array_equal(pose, oracle.Model(name).post_process_pose(np.array(d["state"]["pose"])))
    assert np.array_equal(vel, np.array(d["state"]["vel"]))


def _oracle_ground(oracle, rg):
    kind = oracle.GROUND_VAR2D if rg.kind == 1 else oracle.GROUND_VAR3D
    return oracle.Ground(kind, rg.pieces())


def _probe_positions(pieces, rng, count):
    lo = min(p["bounds"][0] for p in pieces) - 1.0
    hi = max(p["bounds"][1] for p in pieces) + 1.0
    pos = np.zeros((count, 3))
    pos[:, 0] = rng.uniform(lo, hi, count)
    pos[:, 1] = rng.normal(size=count)
    zlo = min(max(p["bounds"][2], -30) for p in pieces) - 1.0
    zhi = max(min(p["bounds"][3], 30) for p in pieces) + 1.0
    pos[:, 2] = rng.uniform(zlo, zhi, count)
    # grid vertices, piece boundaries and the 1e-4 tolerance band
    extra = []
    for p in pieces:
        w = p["res"][0]
        for i in (0, 1, w // 2, w - 2, w - 1):
            xv = p["origin"][0] + p["scale"][0] * (i - (w - 1) * 0.5)
            for dx in (0.0, 1e-9, -1e-9, 5e-5 * p["scale"][0], -5e-5 * p["scale"][0], 2e-4 * p["scale"][0]):
                extra.append([xv + dx, 0.0, p["origin"][2]])
        for b in (p["bounds"][0], p["bounds"][1]):
            for dx in (0.0, 1e-7, -1e-7, 1e-3, -1e-3):
                extra.append([b + dx, 0.0, p["origin"][2]])
    return np.concatenate([pos, np.array(extra)], axis=0)


@needs_tree
@pytest.mark.parametrize("terrain,scale", [("flat.txt", 1.0), ("mixed.txt", 1.0), ("mixed_hopper.txt", 4.0),
                                           ("gaps.txt", 4.0), ("flat3d.txt", 1.0), ("flat3d.txt", 4.0)])
def test_ground_sample_height_bit_exact(oracle, terrain, scale):
    """cGroundVar2D / cGroundVar3D::SampleHeight on terrain the reference generated itself
    (sim/GroundVar2D.cpp:109-135,593-700, sim/GroundVar3D.cpp:154-181,564-684), before and after scrolling"""
    rg = ref_py.RefGround(os.path.join(REF, "data", "terrain", terrain), world_scale=scale, seed=1234)
    rng = np.random.default_rng(9)
    for step in range(3):
        pieces = rg.pieces()
        og = _oracle_ground(oracle, rg)
        pos = _probe_positions(pieces, rng, 4000)
        h0, v0, _ = oracle.sample_height_batch(og, pos)
        h1, v1 = rg.sample_height(pos)
        assert np.array_equal(v0 != 0, v1 != 0)
        assert np.array_equal(h0, h1), np.abs(h0 - h1).max()
        # scroll: move the character bound past the far edge so a segment / slab is regenerated
        xs = max(p["bounds"][1] for p in pieces)
        rg.update([xs - 1.0, 0, -1.0], [xs + 1.0, 0, 1.0])


@pytest.mark.parametrize("null_body", [False, True])
def test_synthetic_character_with_planar_and_fixed_joints(oracle, null_body):
    """non-root planar / fixed joints (no bundled character has them) and an invalid (shape-null) body:
    M, C, Jacobians, FK, pose algebra and stable PD against the reference"""
    from helpers import synthetic_character, synthetic_state
    jm, bd, pd = synthetic_character(with_null_body=null_body)
    m = oracle.Model(joint_mat=jm, body_defs=bd, pd_params=pd)
    r = ref_py.RefModel(jm, bd, pd)
    assert m.n == r.n
    pose, vel, tar, tar_vel = synthetic_state(jm, 6)
    for e in range(6):
        M0, C0 = m.rbd_update(pose[e], vel[e])
        M1, C1 = r.rbd_update(pose[e], vel[e])
        assert rel(M0, M1) <= TOL and rel(C0, C1) <= TOL
        for u, v in zip(m.rbd_extras(pose[e]), r.rbd_extras(pose[e])):
            assert rel(u, v) <= TOL
        for j in range(m.N):
            assert rel(m.joint_world_trans(pose[e], j), r.joint_world_trans(pose[e], j)) <= TOL
            assert rel(m.end_eff_jacobian(pose[e], j), r.end_eff_jacobian(pose[e], j)) <= TOL
            assert bool(oracle.lib().trlo_pd_valid(m.ref, j)) == r.pd_valid(j)
        assert rel(m.vel_to_pose_diff(pose[e], vel[e]), r.vel_to_pose_diff(pose[e], vel[e])) <= TOL
        assert rel(m.calc_vel(pose[e], tar[e] + pose[e] * 0, 1.0), r.calc_vel(pose[e], tar[e], 1.0)) <= TOL
        assert rel(m.lerp_poses(pose[e], pose[(e + 1) % 6], 0.4), r.lerp_poses(pose[e], pose[(e + 1) % 6], 0.4)) <= TOL
        t0, _, _, _ = m.imp_pd(1.0 / 600, pose[e], vel[e], tar[e], tar_vel[e])
        t1, _, _ = r.imp_pd(1.0 / 600, pose[e], vel[e], tar[e], tar_vel[e])
        assert rel(t0, t1) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_world_coordinate_targets_and_explicit_pd(oracle, name):
    """UseWorldCoord targets (cPDController::GetTargetTheta, sim/PDController.cpp:223-273) through stable PD, and
    explicit PD (cExpPDController / cPDController::CalcJointTau*, sim/PDController.cpp:302-428), with the characters'
    own UseWorldCoord flags (hopper, dog, raptor, biped3D have world-coordinate joints)"""
    m, r = models(oracle, name, world_coord=True)
    E = 8
    x = synth.make_inputs(name, E)
    rng = np.random.default_rng(2)
    for e in range(E):
        act = x["joint_active"][e]
        kp = kd = None
        if e % 3 == 2:
            kp, kd = rng.uniform(10, 500, m.N), rng.uniform(1, 50, m.N)
        t0, _, _, _ = m.imp_pd(1.0 / 600, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act, kp, kd)
        t1, _, _ = r.imp_pd(1.0 / 600, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act, kp, kd)
        assert rel(t0, t1) <= TOL
        tau0 = rng.normal(size=m.n)
        e0 = m.exp_pd(1.0 / 600, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act, kp, kd, tau0)
        e1 = r.exp_pd(1.0 / 600, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act, kp, kd, tau0)
        assert rel(e0, e1) <= TOL
    assert np.array_equal(m.exp_pd(0.0, x["pose"][0], x["vel"][0], x["tar_pose"][0], x["tar_vel"][0]), np.zeros(m.n))


def _same_pieces(pa, pb):
    assert len(pa) == len(pb)
    for a, b in zip(pa, pb):
        assert tuple(a["res"]) == tuple(b["res"])
        for k in ("origin", "scale", "bounds"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(np.asarray(a["data"]).ravel(), np.asarray(b["data"]).ravel())


TERRAIN_FILES = ["flat.txt", "gaps.txt", "steps.txt", "walls.txt", "bumps.txt", "mixed.txt", "mixed_hopper.txt",
                 "narrow_gaps.txt", "slopes.txt", "slopes_mixed.txt", "cliffs_rugged.txt"]


@needs_tree
@pytest.mark.parametrize("terrain", TERRAIN_FILES)
def test_terrain_generator_and_rng_bit_exact(oracle, terrain):
    """cTerrainGen2D::Build* with cRand on std::default_random_engine (sim/TerrainGen2D.cpp:127-653, util/Rand.cpp):
    float heights bit-identical; parameter files through cTerrainGen2D::LoadParams"""
    path = os.path.join(REF, "data", "terrain", terrain)
    if not os.path.exists(path):
        pytest.skip("no such terrain file")
    tf = oracle.load_terrain_file(path)
    p = None
    if tf["params"] is not None:
        assert np.array_equal(ref_py.terrain_params_2d(path, 0), tf["params"][0])
        p = tf["params"][0]
    for seed in (0, 1, 12345, 2147483647, 2 ** 40 + 17):
        for width in (20.0, 3.3):
            a = oracle.terrain_gen_2d(tf["type"], width, p, seed)
            b = ref_py.terrain_gen_2d(tf["type_name"], width, p, seed)
            assert a.shape == b.shape and np.array_equal(a, b)


@needs_tree
@pytest.mark.parametrize("terrain,scale", [("mixed.txt", 1.0), ("mixed_hopper.txt", 4.0), ("flat.txt", 1.0),
                                           ("slopes_mixed.txt", 1.0), ("gaps.txt", 4.0)])
def test_ground_var2d_segments_and_scrolling(oracle, terrain, scale):
    """cGroundVar2D::Init / Update / BuildSegment / AddPadding + tSegment::Init (sim/GroundVar2D.cpp:33-96,261-520):
    heights, origin, scaling and AABB bounds identical through forward / backward scrolling and re-initialisation"""
    path = os.path.join(REF, "data", "terrain", terrain)
    tf = oracle.load_terrain_file(path)
    for seed in (3, 77):
        rg = ref_py.RefGround(path, world_scale=scale, seed=seed)
        og = oracle.GroundVar2D(tf["type"], tf["params"], tf["ground_width"], scale, seed)
        _same_pieces(og.pieces(), rg.pieces())
        for step in range(6):
            pcs = rg.pieces()
            hi = max(p["bounds"][1] for p in pcs)
            lo = min(p["bounds"][0] for p in pcs)
            if step in (0, 1, 2, 4):
                bmin, bmax = hi - 1.0, hi + 1.0
            elif step == 3:
                bmin, bmax = lo - 1.0, lo + 1.0
            else:
                bmin, bmax = hi + 100, hi + 102
            rg.update([bmin, 0, -1], [bmax, 0, 1])
            og.update(bmin, bmax)
            _same_pieces(og.pieces(), rg.pieces())


@needs_tree
@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_ground_var3d_slabs_and_scrolling(oracle, scale):
    """cGroundVar3D::Init / Update / BuildSlab with cTerrainGen3D::BuildFlat (sim/GroundVar3D.cpp:24-141,350-420,511-550)"""
    rg = ref_py.RefGround(os.path.join(REF, "data", "terrain", "flat3d.txt"), world_scale=scale, seed=1)
    og = oracle.GroundVar3D(0.2, 0.2, scale)
    _same_pieces(og.pieces(), rg.pieces())
    for bmin, bmax in [([19, 0, -1], [21, 0, 1]), ([39, 0, -1], [41, 0, 1]), ([30, 0, 19], [32, 0, 21]),
                       ([-5, 0, 30], [-3, 0, 32]), ([500, 0, 500], [501, 0, 501]), ([480, 0, 470], [481, 0, 471])]:
        rg.update(bmin, bmax)
        og.update(bmin, bmax)
        _same_pieces(og.pieces(), rg.pieces())
