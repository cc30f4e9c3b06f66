"""Golden-vector tests. The fixtures under tests/golden/ were produced by the REFERENCE'S OWN translation units
(tools/make_golden.py through oracle/_ref, see oracle/ref_wrap.cpp); these tests need neither the reference tree nor
oracle/_ref, so they also run on the GPU box.

  * not gpu : the C oracle against the reference's outputs (<= 1e-13; terrain heights / validity bit-exact)
  * gpu     : the CUDA path, through the C-ABI, against the reference's outputs (<= 1e-9, north_star's bar;
              terrain heights / validity bit-exact)
"""
import glob
import os

import numpy as np
import pytest

from conftest import CONFIG_NAMES
from helpers import assert_close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-13


def load(name):
    path = os.path.join(GOLDEN, name + ".npz")
    assert os.path.exists(path), "missing fixture %s (python tools/make_golden.py)" % path
    return dict(np.load(path))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1.0)
    return float(np.abs(a - b).max() / scale) if a.size else 0.0


def oracle_model(oracle, fx, world=False, frames=True):
    pd = fx["pd_params"].copy()
    if not world:
        pd[:, 17] = 0
    kw = {}
    if frames and "frames" in fx:
        fr = fx["frames"].copy()  # cumulative start times -> durations (what oracle.Model expects)
        dur = np.diff(fr[:, 0], append=fr[-1, 0])
        fr[:, 0] = dur
        kw = {"frames": fr, "loop": bool(fx["loop"])}
    return oracle.Model(joint_mat=fx["joint_mat"], body_defs=fx["body_defs"], pd_params=pd, **kw)


ALL = CONFIG_NAMES + ["synthetic0", "synthetic1"]


def test_fixture_inventory():
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    for n in ALL:
        assert "ref_%s.npz" % n in names
    assert sum(n.startswith("ref_ground_") for n in names) >= 4


@pytest.mark.parametrize("name", ALL)
def test_oracle_dynamics_against_reference_outputs(oracle, name):
    fx = load("ref_" + name)
    m = oracle_model(oracle, fx, frames=False)
    E = fx["pose"].shape[0]
    for e in range(E):
        pose, vel = fx["pose"][e], fx["vel"][e]
        tau, M, C, _ = m.imp_pd(fx["dt"][e], pose, vel, fx["tar_pose"][e], fx["tar_vel"][e], fx["joint_active"][e])
        assert rel(M, fx["M"][e]) <= TOL and rel(C, fx["C"][e]) <= TOL and rel(tau, fx["tau"][e]) <= TOL
        J, Jc, g = m.rbd_extras(pose)
        assert rel(J, fx["J"][e]) <= TOL and rel(Jc, fx["Jcom"][e]) <= TOL and rel(g, fx["g"][e]) <= TOL
        for j in range(m.N):
            assert rel(m.joint_world_trans(pose, j), fx["joint_world"][e][j]) <= TOL
            if fx["body_defs"][j, 0] != 3:
                assert rel(m.body_part_pos(pose, j), fx["body_pos"][e][j]) <= TOL
        com, cv = oracle.calc_com(m, pose, vel)
        assert rel(com, fx["com"][e]) <= TOL and rel(cv, fx["com_vel"][e]) <= TOL
        assert rel(m.vel_to_pose_diff(pose, vel), fx["pose_diff"][e]) <= TOL
        assert rel(m.calc_vel(pose, fx["pose"][(e + 1) % E], 1.0), fx["pose_err"][e]) <= TOL
        if "tau_gains" in fx:
            t2 = m.imp_pd(fx["dt"][e], pose, vel, fx["tar_pose"][e], fx["tar_vel"][e], fx["joint_active"][e],
                          fx["kp"][e], fx["kd"][e])[0]
            assert rel(t2, fx["tau_gains"][e]) <= TOL
            assert rel(m.lerp_poses(pose, fx["pose"][(e + 3) % E], 0.1 + 0.1 * e), fx["lerp"][e]) <= TOL
    for e in range(fx["end_eff_J"].shape[0]):
        for j in range(m.N):
            assert rel(m.end_eff_jacobian(fx["pose"][e], j), fx["end_eff_J"][e][j]) <= TOL


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_oracle_kin_character_against_reference_outputs(oracle, name):
    fx = load("ref_" + name)
    m = oracle_model(oracle, fx)
    assert np.array_equal(m.frames, fx["frames"])
    for i, t in enumerate(fx["kin_time"]):
        idx, ph = m.motion_index_phase(t)
        assert idx == fx["kin_idx"][i] and abs(ph - fx["kin_phase"][i]) <= 1e-12
        p, v = m.kin_char_pose(t, fx["kin_origin_pos"][i], fx["kin_origin_rot"][i])
        assert rel(p, fx["kin_pose"][i]) <= TOL
        assert rel(v, fx["kin_vel"][i]) <= 60 * TOL


def ground_from_fixture(oracle, fx, step):
    kind = oracle.GROUND_VAR2D if int(fx["kind"]) == 1 else oracle.GROUND_VAR3D
    pieces = []
    for s in range(int(fx["s%d_num_pieces" % step])):
        pre = "s%d_p%d_" % (step, s)
        pieces.append({"data": fx[pre + "data"], "res": tuple(int(v) for v in fx[pre + "res"]),
                       "origin": fx[pre + "origin"], "scale": fx[pre + "scale"], "bounds": fx[pre + "bounds"]})
    return kind, pieces


GROUNDS = ["mixed", "mixed_hopper_ws4", "flat", "flat3d"]


@pytest.mark.parametrize("tag", GROUNDS)
def test_oracle_sample_height_against_reference_outputs(oracle, tag):
    fx = load("ref_ground_" + tag)
    for step in range(2):
        kind, pieces = ground_from_fixture(oracle, fx, step)
        h, valid, _ = oracle.sample_height_batch(oracle.Ground(kind, pieces), fx["s%d_pos" % step])
        assert np.array_equal(valid != 0, fx["s%d_valid" % step] != 0)
        assert np.array_equal(h, fx["s%d_h" % step])


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_oracle_world_targets_and_explicit_pd_against_reference_outputs(oracle, name):
    fx = load("ref_" + name)
    mw = oracle_model(oracle, fx, world=True, frames=False)
    ml = oracle_model(oracle, fx, world=False, frames=False)
    for e in range(fx["pose"].shape[0]):
        args = (fx["dt"][e], fx["pose"][e], fx["vel"][e], fx["tar_pose"][e], fx["tar_vel"][e], fx["joint_active"][e])
        assert rel(mw.imp_pd(*args)[0], fx["tau_world_targets"][e]) <= TOL
        assert rel(ml.exp_pd(*args), fx["exp_tau"][e]) <= TOL


@pytest.mark.parametrize("tag", GROUNDS)
def test_oracle_terrain_generation_against_reference_outputs(oracle, tag):
    """the terrain pieces in the fixtures were GENERATED by the reference (seeded cRand + cTerrainGen2D / 3D, then one
    scroll step): the oracle's generator must reproduce heights, placement and bounds bit for bit"""
    fx = load("ref_ground_" + tag)
    scale, seed = float(fx["world_scale"]), int(fx["seed"])
    if int(fx["kind"]) == 1:
        type_id = oracle.GROUND_TYPE_NAMES.index(str(fx["type_name"]))
        og = oracle.GroundVar2D(type_id, fx["params"], float(fx["ground_width"]), scale, seed)
    else:
        og = oracle.GroundVar3D(0.2, 0.2, scale, float(fx["ground_width"]))
    for step in range(2):
        _, want = ground_from_fixture(oracle, fx, step)
        got = og.pieces()
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert tuple(a["res"]) == tuple(b["res"])
            for k in ("origin", "scale", "bounds"):
                assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(np.asarray(a["data"]).ravel(), np.asarray(b["data"]).ravel())
        lo, hi = fx["s%d_update_min" % step], fx["s%d_update_max" % step]
        if int(fx["kind"]) == 1:
            og.update(float(lo[0]), float(hi[0]))
        else:
            og.update(lo, hi)


# ------------------------------------------------------------------------------------------------------------------
# GPU: the product through the C-ABI against the reference's outputs
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def trl():
    import terrainrlsim_b200 as t
    return t


def product_model(trl, fx, world=False, frames=True):
    pd = fx["pd_params"].copy()
    if not world:
        pd[:, 17] = 0
    d = {"joint_mat": fx["joint_mat"], "body_defs": fx["body_defs"], "pd_params": pd}
    if frames and "frames" in fx:
        fr = fx["frames"].copy()
        fr[:, 0] = np.diff(fr[:, 0], append=fr[-1, 0])
        d["motion"] = {"frames": fr, "loop": bool(fx["loop"])}
    return trl.CharModel.from_dict(d)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL)
def test_gpu_dynamics_against_reference_outputs(trl, name):
    fx = load("ref_" + name)
    pm = product_model(trl, fx, frames=False)
    E = fx["pose"].shape[0]
    batch = trl.Batch(pm, E, 0)
    # the batched call takes one dt: group the envs by dt
    tau = np.zeros_like(fx["tau"])
    mass = np.zeros_like(fx["M"])
    bias = np.zeros_like(fx["C"])
    for dt in np.unique(fx["dt"]):
        t, Mm, Cc = batch.imp_pd_update_control_force(float(dt), fx["pose"], fx["vel"], fx["tar_pose"], fx["tar_vel"],
                                                      joint_active=fx["joint_active"], want_mass=True, want_bias=True)
        sel = fx["dt"] == dt
        tau[sel], mass[sel], bias[sel] = t[sel], Mm[sel], Cc[sel]
    assert_close(mass.reshape(E, -1), fx["M"].reshape(E, -1), "mass matrix vs reference")
    assert_close(bias, fx["C"], "bias force vs reference")
    assert_close(tau, fx["tau"], "stable-PD torque vs reference")
    if "tau_gains" in fx:
        t2 = np.zeros_like(tau)
        for dt in np.unique(fx["dt"]):
            t, _, _ = batch.imp_pd_update_control_force(float(dt), fx["pose"], fx["vel"], fx["tar_pose"], fx["tar_vel"],
                                                        joint_active=fx["joint_active"], kp=fx["kp"], kd=fx["kd"])
            t2[fx["dt"] == dt] = t[fx["dt"] == dt]
        assert_close(t2, fx["tau_gains"], "stable-PD torque with gain overrides vs reference")
    jw, bp = batch.kin_tree_fk(fx["pose"])
    assert_close(jw.reshape(E, -1), fx["joint_world"][:, :, :3, :].reshape(E, -1), "joint world transforms vs reference")
    valid = fx["body_defs"][:, 0] != 3
    assert_close(bp[:, valid].reshape(E, -1), fx["body_pos"][:, valid].reshape(E, -1), "body positions vs reference")
    J, Jc, g = batch.rbd_extras(fx["pose"])
    assert_close(J.reshape(E, -1), fx["J"].reshape(E, -1), "Jacobian vs reference")
    assert_close(Jc.reshape(E, -1), fx["Jcom"].reshape(E, -1), "CoM Jacobian vs reference")
    assert_close(g, fx["g"], "gravity force vs reference")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_gpu_kin_character_against_reference_outputs(trl, name):
    fx = load("ref_" + name)
    pm = product_model(trl, fx)
    T = fx["kin_time"].size
    batch = trl.Batch(pm, T, 0)
    pose, vel = batch.kin_char_pose(fx["kin_time"], fx["kin_origin_pos"], fx["kin_origin_rot"])
    assert_close(pose, fx["kin_pose"], "kin pose vs reference")
    assert_close(vel, fx["kin_vel"], "kin vel vs reference", rtol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", GROUNDS)
def test_gpu_sample_height_against_reference_outputs(trl, oracle, tag):
    fx = load("ref_ground_" + tag)
    pm = trl.CharModel.from_name("biped3d" if int(fx["kind"]) == 2 else "dog")
    for step in range(2):
        kind, pieces = ground_from_fixture(oracle, fx, step)
        pos = fx["s%d_pos" % step]
        batch = trl.Batch(pm, 1, 0)
        batch.ground_set_heightfields(kind, [pieces])
        h, valid, _ = batch.ground_sample_height(pos[None])
        assert np.array_equal(valid[0] != 0, fx["s%d_valid" % step] != 0)
        assert np.array_equal(h[0], fx["s%d_h" % step])
