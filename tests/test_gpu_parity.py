"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI
(terrainrlsim_b200.api -> ctypes -> libtrl_b200.so), against the CPU oracle on the same seeded inputs.
Bars: 1e-9 relative on torques / features / kinematics, bit-exact integer terrain cells and validity."""
import numpy as np
import pytest

from conftest import CONFIG_NAMES
from helpers import assert_close, oracle_cfg, product_cfg, oracle_grounds, rel_err
from terrainrlsim_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trl():
    import terrainrlsim_b200 as t
    return t


def _setup(trl, oracle, name, E, with_ground=True, num_fields=8):
    om = oracle.Model(name)
    pm = trl.CharModel.from_name(name)
    batch = trl.Batch(pm, E, 0, product_cfg(trl, name))
    x = synth.make_inputs(name, E)
    grounds = None
    if with_ground:
        kind, fields = synth.make_grounds(name, num_fields)
        batch.ground_set_heightfields(kind, fields)
        grounds = oracle_grounds(oracle, kind, fields)
    return om, pm, batch, x, grounds


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_imp_pd_torque_mass_bias(trl, oracle, name):
    E = 192
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    dt = 1.0 / 600
    rng = np.random.default_rng(5)
    tau0 = rng.normal(size=(E, om.n))
    tau = tau0.copy()
    tau, mass, bias = batch.imp_pd_update_control_force(dt, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"],
                                                        joint_active=x["joint_active"], tau=tau, want_mass=True,
                                                        want_bias=True)
    ref_tau = np.zeros_like(tau)
    ref_M = np.zeros_like(mass)
    ref_C = np.zeros_like(bias)
    for e in range(E):
        ref_tau[e], ref_M[e], ref_C[e], _ = om.imp_pd(dt, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e],
                                                       x["joint_active"][e], tau0=tau0[e])
    assert_close(mass.reshape(E, -1), ref_M.reshape(E, -1), "mass matrix")
    assert_close(bias, ref_C, "bias force")
    assert_close(tau, ref_tau, "torque")
    # dead dofs are exact zeros in M and C (Q1)
    assert np.all(mass[:, 6, :] == 0) and np.all(mass[:, :, 6] == 0) and np.all(bias[:, 6] == 0)
    assert np.all(batch.status() == 0)


@pytest.mark.parametrize("name", ["hopper", "biped3d"])
def test_imp_pd_gain_overrides_and_hopper_dt(trl, oracle, name):
    E = 64
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    dt = 1.0 / 2400
    rng = np.random.default_rng(7)
    kp = rng.uniform(0, 500, size=(E, om.N))
    kd = rng.uniform(0, 50, size=(E, om.N))
    kd[:, -1] = 0  # a zero Kd on a live joint: diagonal gets nothing extra
    act = (rng.uniform(size=(E, om.N)) < 0.7).astype(np.uint8)
    tau, _, _ = batch.imp_pd_update_control_force(dt, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"],
                                                  joint_active=act, kp=kp, kd=kd)
    ref = np.zeros_like(tau)
    for e in range(E):
        ref[e] = om.imp_pd(dt, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e], act[e], kp[e], kd[e])[0]
    assert_close(tau, ref, "torque with overrides")
    # dt <= 0: UpdateControlForce does nothing
    t2 = np.full((E, om.n), 3.0)
    batch.imp_pd_update_control_force(0.0, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"], tau=t2)
    assert np.all(t2 == 3.0)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_kin_tree_fk(trl, oracle, name):
    E = 96
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    jw, bp = batch.kin_tree_fk(x["pose"])
    ref_jw = np.zeros_like(jw)
    ref_bp = np.zeros_like(bp)
    for e in range(E):
        for j in range(om.N):
            ref_jw[e, j] = om.joint_world_trans(x["pose"][e], j)[:3, :].reshape(-1)
            ref_bp[e, j] = om.body_part_pos(x["pose"][e], j)
    assert_close(jw.reshape(E, -1), ref_jw.reshape(E, -1), "joint world transforms")
    assert_close(bp.reshape(E, -1), ref_bp.reshape(E, -1), "body positions")


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_kin_char_pose(trl, oracle, name):
    E = 128
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    t = x["kin_time"].copy()
    t[:8] = [0.0, -0.37, 1e-9, om.frames[-1, 0], 2 * om.frames[-1, 0], om.frames[1, 0], -5.2, 123.456]
    pose, vel = batch.kin_char_pose(t, x["origin_pos"], x["origin_rot"])
    ref_p = np.zeros_like(pose)
    ref_v = np.zeros_like(vel)
    for e in range(E):
        ref_p[e], ref_v[e] = om.kin_char_pose(t[e], x["origin_pos"][e], x["origin_rot"][e])
    assert_close(pose, ref_p, "kin pose")
    assert_close(vel, ref_v, "kin vel", rtol=1e-8)  # finite difference /(1/60) amplifies rounding by 60x
    # NULL origin = zero / identity
    pose2, _ = batch.kin_char_pose(t, None, None)
    for e in range(4):
        assert rel_err(pose2[e], om.kin_char_pose(t[e])[0]) <= 1e-9


def test_kin_char_pose_non_looping(trl, oracle):
    d = dict(synth.load_model_dict("dog"))
    d["motion"] = {"loop": False, "frames": d["motion"]["frames"]}
    om = oracle.Model(joint_mat=d["joint_mat"], body_defs=d["body_defs"], pd_params=d["pd_params"],
                      frames=np.array(d["motion"]["frames"]), loop=False)
    pm = trl.CharModel.from_dict(d)
    E = 32
    batch = trl.Batch(pm, E, 0)
    t = np.linspace(-0.2, 1.2 * om.frames[-1, 0], E)
    pose, vel = batch.kin_char_pose(t)
    for e in range(E):
        rp, rv = om.kin_char_pose(t[e])
        assert rel_err(pose[e], rp) <= 1e-9
        assert rel_err(vel[e], rv) <= 1e-8


@pytest.mark.parametrize("name", ["dog", "biped3d"])
def test_ground_sample_height_bit_exact(trl, oracle, name):
    E, S = 16, 300
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=5)
    rng = np.random.default_rng(11)
    pos = np.zeros((E, S, 3))
    pos[:, :, 0] = rng.uniform(-45, 45, (E, S))
    pos[:, :, 1] = rng.normal(size=(E, S))
    pos[:, :, 2] = rng.uniform(-45, 45, (E, S)) if name == "biped3d" else rng.uniform(-1.2, 1.2, (E, S))
    # exact grid vertices, piece boundaries, the tolerance band, and points far outside
    pc = grounds[0].pieces[1]
    sx = pc["scale"][0]
    for k in range(40):
        pos[0, k, 0] = pc["origin"][0] + sx * (k - (pc["res"][0] - 1) * 0.5)
    pos[1, :6, 0] = [pc["bounds"][0], pc["bounds"][1], pc["bounds"][0] - 1e-12, pc["bounds"][1] + 1e-12, 1e9, -1e9]
    pos[2, :4, 0] = [pc["origin"][0] + sx * ((pc["res"][0] - 1) * 0.5 + 0.00005),
                     pc["origin"][0] + sx * ((pc["res"][0] - 1) * 0.5 + 0.0002), 0.0, -0.0]
    h, valid, cell = batch.ground_sample_height(pos)
    for e in range(E):
        g = grounds[e % len(grounds)]
        for s in range(S):
            rh, rv, rc = g.sample_height(pos[e, s])
            assert tuple(cell[e, s]) == rc, (e, s, cell[e, s], rc)
            assert bool(valid[e, s]) == rv
            assert (h[e, s] == rh) or (np.isnan(h[e, s]) and np.isnan(rh)), (e, s, h[e, s], rh)
    assert np.any(~np.isfinite(h)) and np.any(valid == 0) and np.any(valid == 1)


@pytest.mark.parametrize("name", ["dog", "biped3d"])
def test_ground_sample_height_bit_exact_large(trl, oracle, name):
    """3e6 random positions per terrain kind: cell indices, validity and heights bit-identical to the oracle. This is
    also the exactness check of the FMA-based division by the (constant) grid spacing used on the GPU."""
    E, S = 32, 100000
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=4)
    rng = np.random.default_rng(17)
    pos = np.zeros((E, S, 3))
    pos[:, :, 0] = rng.uniform(-41, 41, (E, S))
    pos[:, :, 2] = rng.uniform(-41, 41, (E, S)) if name == "biped3d" else rng.uniform(-1.1, 1.1, (E, S))
    pos[:, : S // 4, 0] *= 1e-3  # many samples near the shared piece boundary at x = 0
    h, valid, cell = batch.ground_sample_height(pos)
    for e in range(E):
        rh, rv, rc = oracle.sample_height_batch(grounds[e % len(grounds)], pos[e])
        assert np.array_equal(cell[e], rc)
        assert np.array_equal(valid[e].astype(np.int32), rv)
        assert np.array_equal(h[e], rh)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_build_poli_state(trl, oracle, name):
    E = 96
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=6)
    cfg = oracle_cfg(oracle, name)
    assert batch.F == oracle.poli_state_size(om, cfg)
    state, cells = batch.build_poli_state(x["pose"], x["body_pos"], x["body_vel"], vel=x["vel"], phase=x["phase"],
                                          flip=x["flip"], want_cells=True)
    ref = np.zeros_like(state)
    mism = 0
    for e in range(E):
        g = grounds[e % len(grounds)]
        ref[e] = oracle.build_poli_state(om, g, cfg, x["pose"][e], x["vel"][e], x["body_pos"][e], x["body_vel"][e],
                                         phase=x["phase"][e], flip=int(x["flip"][e]))
        if cfg.num_samples > 0:
            _, rc = oracle.parse_ground(g, cfg, x["pose"][e], int(x["flip"][e]))
            mism += int(np.sum(np.any(rc != cells[e, :cfg.num_samples], axis=1)))
    assert_close(state, ref, "policy state")
    assert mism == 0, "terrain cell index mismatches: %d" % mism


def test_build_poli_state_ct3d_and_no_ground(trl, oracle):
    name = "biped3d"
    E = 48
    om = oracle.Model(name)
    pm = trl.CharModel.from_name(name)
    pcfg = trl.make_obs_cfg(obs_kind=3, sampler=1, num_samples=64, grid_res=8, view_min=-1.0, view_max=3.0, num_phase_bins=4)
    ocfg = oracle.obs_cfg(obs_kind=3, sampler=1, num_samples=64, grid_res=8, view_min=-1.0, view_max=3.0, num_phase_bins=4)
    batch = trl.Batch(pm, E, 0, pcfg)
    x = synth.make_inputs(name, E)
    rng = np.random.default_rng(3)
    bq = rng.normal(size=(E, om.N, 4))
    bq /= np.linalg.norm(bq, axis=2, keepdims=True)
    ba = rng.normal(size=(E, om.N, 3))
    gv = rng.normal(size=(E, 3)) * 0.1
    kind, fields = synth.make_grounds(name, 3)
    batch.ground_set_heightfields(kind, fields)
    grounds = oracle_grounds(oracle, kind, fields)
    state = batch.build_poli_state(x["pose"], x["body_pos"], x["body_vel"], vel=x["vel"], body_quat=bq, body_angvel=ba,
                                   ground_vel=gv, phase=x["phase"], flip=x["flip"])
    for e in range(E):
        ref = oracle.build_poli_state(om, grounds[e % 3], ocfg, x["pose"][e], x["vel"][e], x["body_pos"][e],
                                      x["body_vel"][e], bq[e], ba[e], gv[e], x["phase"][e], int(x["flip"][e]))
        assert rel_err(state[e], ref) <= 1e-9
    # no ground registered: samples are zeros (cTerrainRLCharController::ParseGround without a ground)
    batch.ground_set_heightfields(0, None)
    state = batch.build_poli_state(x["pose"], x["body_pos"], x["body_vel"], vel=x["vel"], body_quat=bq, body_angvel=ba,
                                   phase=x["phase"], flip=x["flip"])
    for e in range(4):
        ref = oracle.build_poli_state(om, None, ocfg, x["pose"][e], x["vel"][e], x["body_pos"][e], x["body_vel"][e],
                                      bq[e], ba[e], None, x["phase"][e], int(x["flip"][e]))
        assert rel_err(state[e], ref) <= 1e-9


def test_root_outside_terrain_features_propagate_inf(trl, oracle):
    """Q11: a root outside every piece gives -inf heights; ParseGround replaces the root height by 0 but the
    pose block still carries the non-finite values, exactly like the reference arithmetic."""
    name = "dog"
    E = 8
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=2)
    x["pose"][:, 0] = 500.0
    cfg = oracle_cfg(oracle, name)
    state = batch.build_poli_state(x["pose"], x["body_pos"], x["body_vel"], vel=x["vel"])
    for e in range(E):
        ref = oracle.build_poli_state(om, grounds[e % 2], cfg, x["pose"][e], x["vel"][e], x["body_pos"][e], x["body_vel"][e])
        assert np.array_equal(np.isnan(state[e]), np.isnan(ref))
        assert np.array_equal(np.isinf(state[e]), np.isinf(ref))
        assert rel_err(state[e], ref) <= 1e-9


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_step_fused_matches_oracle(trl, oracle, name):
    E = 128
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=7)
    cfg = oracle_cfg(oracle, name)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    out = batch.step_fused(dt, x)
    ref = oracle.step_range(om, cfg, dict(x), grounds, dt)
    assert_close(out["tau"], ref["tau"], "fused tau")
    assert_close(out["state"], ref["state"], "fused state")
    assert_close(out["kin_pose"], ref["kin_pose"], "fused kin pose")
    assert_close(out["kin_vel"], ref["kin_vel"], "fused kin vel", rtol=1e-8)
    assert_close(out["kin_body_pos"].reshape(E, -1), ref["kin_body_pos"].reshape(E, -1), "fused kin body pos")


@pytest.mark.parametrize("name,E", [("biped3d", 1), ("biped3d", 7), ("biped3d", 33), ("dog", 3), ("dog", 65), ("hopper", 9),
                                    ("raptor", 31), ("biped2d", 1)])
def test_step_fused_small_and_ragged_batches(trl, oracle, name, E):
    """Batches smaller than one 32-env tile / one 8-env group of k_obs_rec, and just past a tile boundary: every lane
    mapping (lane = env tiles, 4 lanes per env, thread per (env, body), block per env) has a ragged tail."""
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=3)
    cfg = oracle_cfg(oracle, name)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    out = batch.step_fused(dt, x)
    ref = oracle.step_range(om, cfg, dict(x), grounds, dt)
    assert_close(out["tau"], ref["tau"], "fused tau")
    assert_close(out["state"], ref["state"], "fused state")
    assert_close(out["kin_pose"], ref["kin_pose"], "fused kin pose")
    assert_close(out["kin_vel"], ref["kin_vel"], "fused kin vel", rtol=1e-8)
    assert_close(out["kin_body_pos"].reshape(E, -1), ref["kin_body_pos"].reshape(E, -1), "fused kin body pos")


def test_device_pointer_mode_equals_host_mode(trl, oracle):
    import torch
    name = "biped3d"
    E = 256
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=4)
    dt = 1.0 / 600
    host = batch.step_fused(dt, x)
    dev_in = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    batch.use_torch_stream()
    dev = batch.step_fused(dt, dev_in)
    batch.sync()
    torch.cuda.synchronize()
    for k in host:
        assert np.array_equal(host[k], dev[k].cpu().numpy(), equal_nan=True), k
    batch.set_stream(None)


@pytest.mark.parametrize("name,E", [("biped3d", 2100), ("dog", 2049)])
def test_host_pipeline_slices_equal_device_mode(trl, oracle, name, E):
    """HOST pointer mode pipelines the step over env slices (whole 32-env tiles, ragged last slice) on two lanes of
    streams; the result must be bit-identical to the single-pass DEVICE pointer mode and match the oracle."""
    import torch
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=5)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    host = batch.step_fused(dt, x)
    dev_in = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    batch.use_torch_stream()
    dev = batch.step_fused(dt, dev_in)
    batch.sync()
    torch.cuda.synchronize()
    for k in host:
        assert np.array_equal(host[k], dev[k].cpu().numpy(), equal_nan=True), k
    batch.set_stream(None)
    ref = oracle.step_range(om, oracle_cfg(oracle, name), dict(x), grounds, dt)
    assert_close(host["tau"], ref["tau"], "sliced tau")
    assert_close(host["state"], ref["state"], "sliced state")
    assert_close(host["kin_pose"], ref["kin_pose"], "sliced kin pose")
    assert np.all(batch.status() == 0)


@pytest.mark.parametrize("name,E", [("biped3d", 2100), ("dog", 2049)])
def test_host_step_graph_replay(trl, oracle, name, E):
    """A HOST-pointer step called repeatedly with the same pinned arrays is captured into a CUDA graph on the second call
    and replayed afterwards: every call must read the arrays' CURRENT contents and equal the directly submitted step
    (pageable numpy arrays, which cannot be captured) bit for bit; changing an argument falls back and re-captures."""
    import torch
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=5)
    _, _, direct, _, _ = _setup(trl, oracle, name, E, num_fields=5)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    n, N, F = pm.num_dof, pm.num_joints, batch.F
    hin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in x.items()}
    hout = {"tau": torch.empty((E, n), dtype=torch.float64).pin_memory(),
            "kin_pose": torch.empty((E, n), dtype=torch.float64).pin_memory(),
            "kin_vel": torch.empty((E, n), dtype=torch.float64).pin_memory(),
            "kin_body_pos": torch.empty((E, N, 3), dtype=torch.float64).pin_memory(),
            "state": torch.empty((E, F), dtype=torch.float64).pin_memory()}
    args = batch.make_step_args(dt, {k: t.numpy() for k, t in hin.items()}, {k: t.numpy() for k, t in hout.items()})
    rng = np.random.default_rng(11)
    l0 = batch.launch_count()
    per_call = None
    for it in range(5):
        # new inputs in the SAME arrays (a fresh pose shift, fresh targets)
        hin["pose"][:, 0] += float(rng.normal()) * 0.05
        hin["tar_pose"].add_(torch.from_numpy(rng.normal(size=tuple(hin["tar_pose"].shape)) * 1e-3))
        hin["kin_time"].add_(0.01)
        for t in hout.values():
            t.fill_(float("nan"))
        assert batch.step_fused_args(args) == 0
        l1 = batch.launch_count()
        per_call = per_call or (l1 - l0)
        assert l1 - l0 == per_call  # replays account for the same launches
        l0 = l1
        ref = direct.step_fused(dt, {k: t.numpy().copy() for k, t in hin.items()})
        for k in hout:
            assert np.array_equal(hout[k].numpy(), ref[k], equal_nan=True), (it, k)
    # a different output array: new arguments -> direct call, then a new graph
    hout["tau"] = torch.empty((E, n), dtype=torch.float64).pin_memory()
    args = batch.make_step_args(dt, {k: t.numpy() for k, t in hin.items()}, {k: t.numpy() for k, t in hout.items()})
    for it in range(3):
        assert batch.step_fused_args(args) == 0
        ref = direct.step_fused(dt, {k: t.numpy().copy() for k, t in hin.items()})
        for k in hout:
            assert np.array_equal(hout[k].numpy(), ref[k], equal_nan=True), (it, k)
    # another entry point in between (it uses the staging pool) and a new terrain set must not break the replay
    batch.kin_tree_fk(x["pose"])
    kind, fields = synth.make_grounds(name, 3)
    batch.ground_set_heightfields(kind, fields)
    direct.ground_set_heightfields(kind, fields)
    for it in range(3):
        assert batch.step_fused_args(args) == 0
        ref = direct.step_fused(dt, {k: t.numpy().copy() for k, t in hin.items()})
        for k in hout:
            assert np.array_equal(hout[k].numpy(), ref[k], equal_nan=True), (it, k)


@pytest.mark.parametrize("name,E", [("hopper", 4096), ("dog", 4096), ("raptor", 8192), ("biped3d", 16384), ("biped2d", 4096)])
def test_full_size_properties(trl, oracle, name, E):
    """BASELINE.json full sizes: determinism, env-permutation equivariance (bit-exact), and oracle parity on a
    random subset of 128 envs."""
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=16)
    cfg = oracle_cfg(oracle, name)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    out1 = {k: v.copy() for k, v in batch.step_fused(dt, x).items()}
    out2 = batch.step_fused(dt, x)
    for k in out1:
        assert np.array_equal(out1[k], out2[k], equal_nan=True), "non-deterministic " + k
    # permutation that keeps each env on its terrain (env % 16 preserved)
    rng = np.random.default_rng(2)
    perm = np.arange(E).reshape(-1, 16)
    perm = perm[rng.permutation(perm.shape[0])].reshape(-1)
    xp = {k: np.ascontiguousarray(v[perm]) for k, v in x.items()}
    outp = batch.step_fused(dt, xp)
    for k in out1:
        assert np.array_equal(out1[k][perm], outp[k], equal_nan=True), "not permutation-equivariant: " + k
    assert np.all(batch.status() == 0)
    sub = np.sort(rng.choice(E, 128, replace=False))
    xs = {k: np.ascontiguousarray(v[sub]) for k, v in x.items()}
    ref = oracle.step_range(om, cfg, xs, grounds, dt, env_to_ground=(sub % 16).astype(np.int32))
    assert_close(out1["tau"][sub], ref["tau"], "full-size tau")
    assert_close(out1["state"][sub], ref["state"], "full-size state")
    assert_close(out1["kin_pose"][sub], ref["kin_pose"], "full-size kin pose")


@pytest.mark.parametrize("name", ["biped3d", "dog", "biped2d"])
def test_batched_sim_adapter(trl, oracle, name):
    """cSimAdapter-shaped front end: shapes, a few frames with the kinematic stand-in world, and getState() parity."""
    from terrainrlsim_b200.sim_adapter import BatchedSimAdapter
    E = 48
    sim = BatchedSimAdapter(name, E)
    sim.setRandomSeed(3)
    sim.init()
    om = oracle.Model(name)
    cfg = oracle_cfg(oracle, name)
    assert sim.getActionSpaceSize() == om.n - 7
    assert sim.getObservationSpaceSize() == oracle.poli_state_size(om, cfg)
    rng = np.random.default_rng(1)
    for frame in range(2):
        if sim.needUpdatedAction():
            pose, _, _, _ = sim.world.state()
            sim.updateAction(pose[:, 7:] + 0.05 * rng.normal(size=(E, om.n - 7)))
        sim.update()
        assert np.all(np.isfinite(sim.tau)) and np.all(sim.tau[:, :7] == 0)
        assert np.all(sim.batch.status() == 0)
    ob = sim.getState()
    assert ob.shape == (E, sim.getObservationSpaceSize())
    pose, vel, body_pos, body_vel = sim.world.state()
    # the adapter's terrains are generated on the device from the config's terrain file: read them back for the oracle
    kind = 2 if name == "biped3d" else 1
    grounds = oracle_grounds(oracle, kind, [sim.batch.ground_read_field(f) for f in range(sim._ground_fields)])
    changed = sim.updateGround(pose[:, :3])  # cScenarioSimChar::UpdateGround: nothing has walked off its terrain yet
    assert changed.shape == (sim._ground_fields,)
    phase = (sim._frame * (1.0 / 30)) % 1.0
    for e in range(E):
        ref = oracle.build_poli_state(om, grounds[e % len(grounds)], cfg, pose[e], vel[e], body_pos[e], body_vel[e],
                                      phase=phase)
        assert rel_err(ob[e], ref) <= 1e-9
    # the torque of the last substep equals the oracle's stable PD on the world's state before that substep
    assert sim.agentHasFallen().shape == (E,)
    r = sim.calcReward()
    assert r.shape == (E,) and np.all((r >= 0) & (r <= 1))
    # the stand-in world tracks the reference exactly: pose / velocity terms are at their maximum
    assert np.all(r[~sim.agentHasFallen()] > 0.5)
    sim.finish()


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_imitation_reward(trl, oracle, name):
    """cScenarioExpImitate::CalcReward, GPU vs oracle (reward and its five error terms)."""
    E = 96
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=5)
    kin_pose, kin_vel = batch.kin_char_pose(x["kin_time"], x["origin_pos"], x["origin_rot"])
    # simulated character = perturbed reference (so every term is O(0.01 .. 1)), plus a few unrelated poses
    rng = np.random.default_rng(9)
    pose = kin_pose + 0.05 * rng.normal(size=kin_pose.shape)
    for o in om.quat_slots():
        pose[:, o:o + 4] /= np.linalg.norm(pose[:, o:o + 4], axis=1, keepdims=True)
    vel = kin_vel + 0.3 * rng.normal(size=kin_vel.shape)
    vel[:, 6] = 0
    pose[:8], vel[:8] = x["pose"][:8], x["vel"][:8]
    fallen = np.zeros(E, dtype=np.uint8)
    fallen[5] = 1
    reward, terms = batch.imitate_calc_reward(pose, vel, kin_pose, kin_vel, x["body_vel"], fallen, want_terms=True)
    ref_r = np.zeros(E)
    ref_t = np.zeros((E, 5))
    for e in range(E):
        ref_r[e], ref_t[e] = oracle.imitate_reward(om, grounds[e % len(grounds)], pose[e], vel[e], kin_pose[e], kin_vel[e],
                                                   x["body_vel"][e], fallen[e])
    assert reward[5] == 0 and np.all(terms[5] == 0)
    assert_close(reward[:, None], ref_r[:, None], "imitation reward")
    assert_close(terms, ref_t, "reward error terms")
    assert np.all((reward >= 0) & (reward <= 1))


_VARIANT_SCRIPT = """
import sys, numpy as np
sys.path.insert(0, {tests!r}); sys.path.insert(0, {root!r})
import terrainrlsim_b200 as trl
from terrainrlsim_b200 import synth
from oracle import oracle_py as oracle
from helpers import rel_err, product_cfg
for name in {names!r}:
    E = 77  # not a multiple of the 32-env scratch tile
    om = oracle.Model(name); pm = trl.CharModel.from_name(name)
    batch = trl.Batch(pm, E, 0, product_cfg(trl, name))
    x = synth.make_inputs(name, E)
    tau, _, _ = batch.imp_pd_update_control_force(1.0 / 600, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"],
                                                  joint_active=x["joint_active"])
    ref = np.stack([om.imp_pd(1.0 / 600, x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e],
                              x["joint_active"][e])[0] for e in range(E)])
    err = rel_err(tau, ref)
    assert err < 1e-9, (name, err)
    assert np.all(batch.status() == 0)
    # the fused step (all three chains) under the same hooks
    from helpers import oracle_cfg, oracle_grounds
    kind, fields = synth.make_grounds(name, 5)
    batch.ground_set_heightfields(kind, fields)
    grounds = oracle_grounds(oracle, kind, fields)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    out = batch.step_fused(dt, x)
    ref = oracle.step_range(om, oracle_cfg(oracle, name), dict(x), grounds, dt)
    for k, tol in (("tau", 1e-9), ("state", 1e-9), ("kin_pose", 1e-9), ("kin_vel", 1e-8), ("kin_body_pos", 1e-9)):
        err = rel_err(out[k].reshape(E, -1), ref[k].reshape(E, -1))
        assert err < tol, (name, k, err)
print("VARIANT_OK")
"""


@pytest.mark.parametrize("env", [{"TRL_PD_ABA": "0"}, {"TRL_OBS_STAGED": "0"}, {"TRL_OBS_STAGED": "1"},
                                 {"TRL_STEP_OVERLAP": "0"}, {"TRL_HOST_SLICES": "1"}, {"TRL_HOST_SLICES": "4"},
                                 {"TRL_OBS_FEAT_SIDE": "0"}, {"TRL_KIN_WARPS": "2"}, {"TRL_PD_SLIM": "1"}, {"TRL_PD_SLIM": "0"}])
def test_pd_kernel_variants(env):
    """The launcher's tuning hooks select other kernel variants (monolithic warp-per-env PD, runtime-loop solve, wider
    lane groups); each must meet the same 1e-9 bar. Hooks are read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    script = _VARIANT_SCRIPT.format(tests=here, root=os.path.dirname(here), names=["hopper", "biped2d", "dog", "biped3d"])
    out = subprocess.run([sys.executable, "-c", script], env=dict(os.environ, **env), capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "VARIANT_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_rbd_extras(trl, oracle, name):
    """world Jacobian, CoM Jacobian and gravity force vs the oracle (1e-9 per env matrix / vector);
    E is not a multiple of the 32-env tile."""
    E = 75
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    J, Jc, g = batch.rbd_extras(x["pose"])
    rJ, rJc, rg = np.zeros_like(J), np.zeros_like(Jc), np.zeros_like(g)
    for e in range(E):
        rJ[e], rJc[e], rg[e] = om.rbd_extras(x["pose"][e])
    assert_close(J.reshape(E, -1), rJ.reshape(E, -1), "world Jacobian")
    assert_close(Jc.reshape(E, -1), rJc.reshape(E, -1), "CoM Jacobian")
    assert_close(g, rg, "gravity force")
    assert np.all(J[:, :, 6] == 0)  # dead root slot
    J2, _, g2 = batch.rbd_extras(x["pose"], want_Jcom=False)
    assert np.array_equal(J2, J) and np.array_equal(g2, g)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_apply_control_forces(trl, oracle, name):
    """clamped joint torques and world-frame body wrenches vs the oracle; torques scaled so that about
    half of the joints hit their limits."""
    E = 70
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    rng = np.random.default_rng(21)
    tau = rng.normal(size=(E, om.n)) * 150.0
    ct, wr = batch.apply_control_forces(x["pose"], tau)
    rct, rwr = np.zeros_like(ct), np.zeros_like(wr)
    for e in range(E):
        rct[e], rwr[e] = om.apply_control_forces(x["pose"][e], tau[e])
    assert_close(ct, rct, "clamped tau")
    assert_close(wr.reshape(E, -1), rwr.reshape(E, -1), "body wrenches")
    assert (np.abs(ct - tau) > 0).any() and (ct == tau).any()  # some joints clamped, some not


def test_widened_rows_device_pointer_mode(trl, oracle):
    """trl_rbd_extras / trl_apply_control_forces / trl_imitate_calc_reward with device pointers (torch CUDA tensors,
    asynchronous on the batch's stream) give bit-identical results to the host-pointer calls."""
    import torch
    name, E = "biped3d", 100
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=True, num_fields=3)
    rng = np.random.default_rng(4)
    tau = rng.normal(size=(E, om.n)) * 120.0
    hJ, hJc, hg = batch.rbd_extras(x["pose"])
    hct, hwr = batch.apply_control_forces(x["pose"], tau)
    kin_pose, kin_vel = batch.kin_char_pose(x["kin_time"], x["origin_pos"], x["origin_rot"])
    hr = batch.imitate_calc_reward(x["pose"], x["vel"], kin_pose, kin_vel, x["body_vel"], None)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in
         dict(pose=x["pose"], vel=x["vel"], tau=tau, kin_pose=kin_pose, kin_vel=kin_vel, body_vel=x["body_vel"]).items()}
    batch.use_torch_stream()
    dJ, dJc, dg = batch.rbd_extras(d["pose"])
    dct, dwr = batch.apply_control_forces(d["pose"], d["tau"])
    dr = batch.imitate_calc_reward(d["pose"], d["vel"], d["kin_pose"], d["kin_vel"], d["body_vel"], None)
    batch.sync()
    torch.cuda.synchronize()
    for a, b_ in ((hJ, dJ), (hJc, dJc), (hg, dg), (hct, dct), (hwr, dwr)):
        assert np.array_equal(a, b_.cpu().numpy(), equal_nan=True)
    assert np.array_equal(hr, dr.cpu().numpy(), equal_nan=True)
    batch.set_stream(None)


def test_host_async_mode_matches_synchronous(trl, oracle):
    """trl_batch_set_host_async: the HOST-pointer step returns after enqueueing and trl_batch_sync completes it; two
    batches driven alternately give the same bits as the synchronous call (pinned buffers)."""
    import torch
    name, E = "biped3d", 2304
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    batches, args, outs, refs = [], [], [], []
    for k in range(2):
        om, pm, batch, x, _ = _setup(trl, oracle, name, E, num_fields=3)
        x = synth.make_inputs(name, E, seed=100 + k)
        refs.append(batch.step_fused(dt, x))
        hin = {kk: torch.from_numpy(v).pin_memory() for kk, v in x.items()}
        hout = {kk: torch.empty(v.shape, dtype=torch.float64).pin_memory() for kk, v in refs[-1].items()}
        batch.set_stream(None)
        args.append((hin, hout, batch.make_step_args(dt, {kk: t.numpy() for kk, t in hin.items()},
                                                      {kk: t.numpy() for kk, t in hout.items()})))
        batch.set_host_async(True)
        batches.append(batch)
        outs.append(hout)
    for rep in range(3):
        for b, a in zip(batches, args):
            assert b.step_fused_args(a[2]) == 0
        for b in batches:
            b.sync()
    for k in range(2):
        for kk, ref in refs[k].items():
            assert np.array_equal(outs[k][kk].numpy(), ref, equal_nan=True), (k, kk)
        batches[k].set_host_async(False)


@pytest.mark.parametrize("name,E", [("biped3d", 2100), ("dog", 96), ("biped2d", 2049)])
def test_step_fused_ex_reward_epilogue_f32_state_default_tar_vel(trl, oracle, name, E):
    """trl_step_fused_ex: the imitation-reward epilogue on the step's own kin pose / vel (which then stay on the device),
    the single-precision state copy and tar_vel == NULL (zero target velocity), in HOST pointer mode (sliced pipeline for
    E >= 2048, single pass below) and in DEVICE pointer mode. Bit-identical to the separate entry points."""
    import torch
    om, pm, batch, x, grounds = _setup(trl, oracle, name, E, num_fields=5)
    d = synth.load_model_dict(name)
    dt = (1.0 / 30) / d["num_update_steps"]
    assert np.all(x["tar_vel"] == 0)
    base = batch.step_fused(dt, x)  # every output, tar_vel given
    fallen = np.zeros(E, dtype=np.uint8)
    fallen[3] = 1
    ref_reward = batch.imitate_calc_reward(x["pose"], x["vel"], base["kin_pose"], base["kin_vel"], x["body_vel"], fallen)
    x0 = {k: v for k, v in x.items() if k != "tar_vel"}
    n, F = pm.num_dof, batch.F

    # HOST pointers: only tau, the fp32 state and the reward come back
    out = {"tau": np.empty((E, n))}
    reward = np.empty(E)
    s32 = np.empty((E, F), dtype=np.float32)
    batch.step_fused_ex(dt, x0, out, reward=reward, fallen=fallen, state_f32=s32)
    assert np.array_equal(out["tau"], base["tau"])
    assert np.array_equal(reward, ref_reward)
    assert np.array_equal(s32, base["state"].astype(np.float32), equal_nan=True)
    assert reward[3] == 0

    # HOST pointers, f64 state + kin pose requested next to the reward
    out = {"tau": np.empty((E, n)), "state": np.empty((E, F)), "kin_pose": np.empty((E, n))}
    reward2 = np.empty(E)
    batch.step_fused_ex(dt, x0, out, reward=reward2, fallen=fallen)
    assert np.array_equal(out["state"], base["state"], equal_nan=True)
    assert np.array_equal(out["kin_pose"], base["kin_pose"])
    assert np.array_equal(reward2, ref_reward)

    # observation skipped (a substep between two policy queries): tau + reward only
    out = {"tau": np.empty((E, n))}
    reward3 = np.empty(E)
    batch.step_fused_ex(dt, x0, out, reward=reward3, fallen=fallen)
    assert np.array_equal(out["tau"], base["tau"]) and np.array_equal(reward3, ref_reward)

    # DEVICE pointers
    dev_in = {k: torch.from_numpy(v).cuda() for k, v in x0.items()}
    dout = {"tau": torch.empty((E, n), dtype=torch.float64, device="cuda")}
    drew = torch.empty(E, dtype=torch.float64, device="cuda")
    ds32 = torch.empty((E, F), dtype=torch.float32, device="cuda")
    batch.use_torch_stream()
    batch.step_fused_ex(dt, dev_in, dout, reward=drew, fallen=torch.from_numpy(fallen).cuda(), state_f32=ds32)
    batch.sync()
    torch.cuda.synchronize()
    assert np.array_equal(dout["tau"].cpu().numpy(), base["tau"])
    assert np.array_equal(drew.cpu().numpy(), ref_reward)
    assert np.array_equal(ds32.cpu().numpy(), base["state"].astype(np.float32), equal_nan=True)
    batch.set_stream(None)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_exp_pd_control_force(trl, oracle, name):
    """cExpPDController::UpdateControlForce (explicit PD incl. world-coordinate targets, inactive
    joints, gain overrides, accumulation into tau) vs the oracle; dt <= 0 is a no-op."""
    E = 77
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    dt = 1.0 / 600
    rng = np.random.default_rng(11)
    tau0 = rng.normal(size=(E, om.n))
    tar_vel = rng.normal(size=(E, om.n))
    kp = rng.uniform(50, 500, size=(E, om.N))
    kd = rng.uniform(5, 50, size=(E, om.N))
    for gains in (False, True):
        tau = batch.exp_pd_update_control_force(dt, x["pose"], x["vel"], x["tar_pose"], tar_vel,
                                                joint_active=x["joint_active"], kp=kp if gains else None,
                                                kd=kd if gains else None, tau=tau0.copy())
        ref = np.zeros_like(tau)
        for e in range(E):
            ref[e] = om.exp_pd(dt, x["pose"][e], x["vel"][e], x["tar_pose"][e], tar_vel[e], x["joint_active"][e],
                               kp=kp[e] if gains else None, kd=kd[e] if gains else None, tau0=tau0[e])
        assert_close(tau, ref, "explicit PD torque")
        assert np.array_equal(tau[:, :7], tau0[:, :7])  # no controller on the root
    same = batch.exp_pd_update_control_force(0.0, x["pose"], x["vel"], x["tar_pose"], tar_vel, tau=tau0.copy())
    assert np.array_equal(same, tau0)
    # tar_vel == NULL is the reference's default zero target velocity
    a = batch.exp_pd_update_control_force(dt, x["pose"], x["vel"], x["tar_pose"], None, tau=tau0.copy())
    b = batch.exp_pd_update_control_force(dt, x["pose"], x["vel"], x["tar_pose"], np.zeros_like(tar_vel), tau=tau0.copy())
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_rbd_end_effector_jacobian(trl, oracle, name):
    """SURVEY.md 8f rank 2 remainder: cRBDUtil::BuildEndEffectorJacobian, one joint for every env and one joint per env."""
    E = 70
    om, pm, batch, x, _ = _setup(trl, oracle, name, E, with_ground=False)
    for jid in sorted({om.N - 1, 1, om.N // 2}):
        J = batch.rbd_end_effector_jacobian(x["pose"], int(jid))
        ref = np.stack([om.end_eff_jacobian(x["pose"][e], jid) for e in range(E)])
        assert_close(J.reshape(E, -1), ref.reshape(E, -1), "end-effector Jacobian (joint %d)" % jid)
    ids = (np.arange(E) % om.N).astype(np.int32)
    J = batch.rbd_end_effector_jacobian(x["pose"], ids)
    ref = np.stack([om.end_eff_jacobian(x["pose"][e], int(ids[e])) for e in range(E)])
    assert_close(J.reshape(E, -1), ref.reshape(E, -1), "end-effector Jacobian (joint per env)")
    with pytest.raises(trl.TrlError):
        batch.rbd_end_effector_jacobian(x["pose"], om.N)


def test_terrainrl_sim_wrapper_batched(trl, oracle):
    """simAdapter/terrainRLSim.py's TerrainRLSimWrapper over the batched adapter: reset / step / spaces, [E]-shaped."""
    from terrainrlsim_b200.envs import TerrainRLSimWrapper
    from terrainrlsim_b200.sim_adapter import BatchedSimAdapter
    E = 40
    sim = BatchedSimAdapter("dog", E)
    sim.init()
    env = TerrainRLSimWrapper(sim, config={"time_limit": 256})
    ob = env.reset()
    A = env.getActionSpace().getMinimum().shape[0]
    assert ob.shape == (E, sim.getObservationSpaceSize()) and A == sim.getActionSpaceSize()
    assert env.observation_space.low.shape == (ob.shape[1],)
    pose, _, _, _ = sim.world.state()
    ob2, reward, done, info = env.step(pose[:, 7:])
    assert ob2.shape == ob.shape and reward.shape == (E,) and done.shape == (E,) and done.dtype == bool and info is None
    assert np.all(np.isfinite(ob2)) and np.all((reward >= 0) & (reward <= 1))
    assert env.endOfEpoch() is done
    env.finish()
