"""CPU-only tests of the host side: the C-ABI library loads and exports every declared symbol, the model /
arg-file loaders follow the reference's data formats, and the product fails loudly without a GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import CONFIG_NAMES, ROOT
from terrainrlsim_b200 import _capi, synth
import terrainrlsim_b200 as trl

REF = "/root/reference"


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "trl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(trl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _capi.lib()
    bound = {s[0] for s in _capi.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert hasattr(lib, name)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = trl.CharModel.from_name("hopper")
    with pytest.raises(trl.TrlError) as ei:
        trl.Batch(m, 4, 0)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(trl.TrlError):
        trl.Batch(m, 4, -1)  # the survey's "device -1 = CPU oracle" idea is deliberately not offered


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_model_from_matrices_matches_fixture_and_oracle_layout(oracle, name):
    d = synth.load_model_dict(name)
    m = trl.CharModel.from_name(name)
    om = oracle.Model(name)
    assert (m.num_joints, m.num_dof) == (d["num_joints"], d["num_dof"]) == (om.N, om.n)
    jm, bd, pd = m.matrices()
    assert np.array_equal(jm, np.array(d["joint_mat"], dtype=float))
    assert np.array_equal(bd, np.array(d["body_defs"], dtype=float))
    assert np.array_equal(pd, np.array(d["pd_params"], dtype=float))
    fr, loop = m.motion()
    assert loop == om.loop
    assert np.array_equal(fr, om.frames)  # cumulative start times (cMotion::PostProcessFrames)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_json_loader_on_reference_files_equals_fixture(name):
    d = synth.load_model_dict(name)
    src = d["source"]
    m = trl.CharModel.load(os.path.join(REF, src["character"]), os.path.join(REF, src["motion"]))
    jm, bd, pd = m.matrices()
    assert np.array_equal(jm, np.array(d["joint_mat"], dtype=float))
    assert np.array_equal(bd, np.array(d["body_defs"], dtype=float))
    assert np.array_equal(pd, np.array(d["pd_params"], dtype=float))
    fr, loop = m.motion()
    assert loop == d["motion"]["loop"]
    assert fr.shape == (len(d["motion"]["frames"]), d["num_dof"] + 1)


CHAR_JSON = """
{
  // comments are tolerated, like the reference's jsoncpp reader
  "Skeleton": { "Joints": [
     {"Name": "root", "Type": "planar", "Parent": -1, "AttachX": 5, "AttachY": 6, "AttachZ": 7, "TorqueLim": 0},
     {"Name": "a", "Type": "revolute", "Parent": 0, "AttachX": 0.1, "AttachThetaZ": 0.5, "LimLow0": -1, "LimHigh0": 2,
      "TorqueLim": 300, "IsEndEffector": 1},
     {"Name": "b", "Type": "spherical", "Parent": 1, "AttachY": -0.3, "DiffWeight": 0.25},
     {"Name": "c", "Type": "prismatic", "Parent": 0, "ForceLim": 10}
  ]},
  "BodyDefs": [
     {"Name": "root", "Shape": "box", "Mass": 8.0, "ColGroup": 1, "Param0": 0.7, "Param1": 0.2, "Param2": 0.2},
     {"Name": "a", "Shape": "capsule", "Mass": 0.5, "AttachY": -0.1, "Param0": 0.03, "Param1": 0.6},
     {"Name": "b", "Shape": "box", "Mass": 1, "Param0": 0.1, "Param1": 0.1, "Param2": 0.1, "ColorA": 0.5},
     {"Name": "c", "Shape": "null"}
  ],
  "PDControllers": [
     {"Name": "root", "Kp": 0, "Kd": 0},
     {"Name": "a", "Kp": 200.0, "Kd": 20.0, "TargetTheta0": 0.3, "UseWorldCoord": 1},
     {"Name": "b", "Kp": 100, "Kd": 10},
     {"Name": "c", "Kp": 1000, "Kd": 20}
  ]
}
"""


def test_json_loader_defaults_and_layout(tmp_path):
    p = tmp_path / "char.txt"
    p.write_text(CHAR_JSON)
    m = trl.CharModel.load(str(p))
    assert (m.num_joints, m.num_dof, m.num_frames) == (4, 7 + 1 + 4 + 1, 0)
    jm, bd, pd = m.matrices()
    inf = float("inf")
    # root: type planar(1), param offset 0, attach point forced to 0 (PostProcessJointMat), defaults elsewhere
    assert list(jm[0]) == [1, -1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, inf, 0, 1, 0]
    assert list(jm[1]) == [0, 0, 0.1, 0, 0, 0, 0, 0.5, -1, 1, 1, 2, 0, 0, 300, inf, 1, 1, 7]
    assert jm[2][0] == 4 and jm[2][18] == 8 and jm[2][17] == 0.25
    assert jm[3][0] == 2 and jm[3][18] == 12 and jm[3][15] == 10
    assert list(bd[0][:4]) == [0, 8.0, 1, 0] and bd[0][16] == 1
    assert bd[1][0] == 1 and bd[1][2] == -1 and bd[1][5] == -0.1
    assert bd[2][16] == 0.5 and bd[3][0] == 3
    assert list(pd[1][:4]) == [1, 200.0, 20.0, 0.3] and pd[1][17] == 1 and pd[3][0] == 3


def test_json_loader_errors(tmp_path):
    with pytest.raises(trl.TrlError):
        trl.CharModel.load(str(tmp_path / "missing.txt"))
    bad = tmp_path / "bad.txt"
    bad.write_text(CHAR_JSON.replace('"revolute"', '"hinge"'))
    with pytest.raises(trl.TrlError) as ei:
        trl.CharModel.load(str(bad))
    assert "Unsupported joint type" in str(ei.value)
    bad.write_text(CHAR_JSON.replace('"Parent": 1', '"Parent": 3'))
    with pytest.raises(trl.TrlError) as ei:
        trl.CharModel.load(str(bad))
    assert "Parent id must be < child id" in str(ei.value)
    bad.write_text("{ not json")
    with pytest.raises(trl.TrlError):
        trl.CharModel.load(str(bad))
    # motion with the wrong number of dofs (cKinCharacter::LoadMotion's check)
    good = tmp_path / "char.txt"
    good.write_text(CHAR_JSON)
    mot = tmp_path / "motion.txt"
    mot.write_text(json.dumps({"Loop": True, "Frames": [[0.1] + [0] * 5, [0.1] + [0] * 5]}))
    with pytest.raises(trl.TrlError) as ei:
        trl.CharModel.load(str(good), str(mot))
    assert "DOF mismatch" in str(ei.value)
    mot.write_text(json.dumps({"Loop": False, "Frames": [[0.1] + [0] * 13, [0.2] + [1] * 13, [0.0] + [2] * 13]}))
    m = trl.CharModel.load(str(good), str(mot))
    fr, loop = m.motion()
    assert not loop and np.allclose(fr[:, 0], [0, 0.1, 0.3])


def test_arg_file_parser(tmp_path):
    p = tmp_path / "args.txt"
    p.write_text("-scenario= imitate_eval\n// -character_file= wrong.txt\n-character_file= data/characters/x.txt\n"
                 "-num_update_steps= 20 // trailing\n-cam_position= 0 1 -2.5\n")
    buf = C.create_string_buffer(256)
    lib = _capi.lib()
    assert lib.trl_arg_file_get(str(p).encode(), b"character_file", buf, 256) == 0
    assert buf.value == b"data/characters/x.txt"
    assert lib.trl_arg_file_get(str(p).encode(), b"num_update_steps", buf, 256) == 0 and buf.value == b"20"
    assert lib.trl_arg_file_get(str(p).encode(), b"cam_position", buf, 256) == 0 and buf.value == b"0 1 -2.5"
    assert lib.trl_arg_file_get(str(p).encode(), b"missing", buf, 256) != 0
    assert lib.trl_arg_file_get(b"/nonexistent/file", b"x", buf, 256) != 0


def test_synthetic_inputs_are_deterministic_and_well_formed():
    for name in CONFIG_NAMES:
        a = synth.make_inputs(name, 64)
        b = synth.make_inputs(name, 64)
        mi = synth.ModelInfo(synth.load_model_dict(name))
        for k in a:
            assert np.array_equal(a[k], b[k])
        assert a["pose"].shape == (64, mi.n)
        assert np.allclose(np.linalg.norm(a["pose"][:, 3:7], axis=1), 1)
        assert np.all(a["vel"][:, 6] == 0)
        c = synth.make_inputs(name, 64, rank=1)
        assert not np.array_equal(a["pose"], c["pose"])
    kind, f = synth.make_grounds("dog", 3)
    assert kind == 1 and len(f) == 3 and len(f[0]) == 2 and f[0][0]["res"][0] >= 801
    kind, f = synth.make_grounds("biped3d", 2)
    assert kind == 2 and len(f[0]) == 4 and f[0][0]["res"] == (201, 201)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_sim_adapter_reads_reference_arg_files():
    """BatchedSimAdapter.init() resolves -character_file / -motion_file / -num_update_steps from an arg file; without
    a GPU it must stop at batch creation with the loud no-fallback error, after the model has loaded."""
    import torch
    from terrainrlsim_b200.sim_adapter import BatchedSimAdapter
    cwd = os.getcwd()
    os.chdir(REF)  # the reference's arg files use paths relative to its root
    try:
        sim = BatchedSimAdapter(["-arg_file=", "args/test_dog_args.txt"], 4)
        if torch.cuda.is_available():
            sim.init()
            assert sim.getActionSpaceSize() == 20 and sim.num_update_steps == 20
            sim.finish()
        else:
            with pytest.raises(trl.TrlError) as ei:
                sim.init()
            assert "no CPU fallback" in str(ei.value)
            assert sim.model.num_joints == 21 and sim.num_update_steps == 20
    finally:
        os.chdir(cwd)


def test_state_file_round_trip(tmp_path):
    """cCharacter::ReadState / WriteState: quaternions normalised on read, missing keys keep the caller's values, the
    writer's number format is the reference's std::to_string ("%f")."""
    import json
    import terrainrlsim_b200 as trl
    m = trl.CharModel.from_name("biped3d")
    d = synth.load_model_dict("biped3d")
    pose0 = np.array(d["state"]["pose"])
    vel0 = np.array(d["state"]["vel"])
    p = tmp_path / "state.txt"
    scaled = pose0.copy()
    scaled[3:7] *= 3.0
    p.write_text(json.dumps({"Pose": scaled.tolist(), "Vel": vel0.tolist()}))
    pose, vel = m.read_state(str(p))
    assert np.allclose(np.linalg.norm(pose[3:7]), 1.0) and np.allclose(pose[:3], pose0[:3])
    assert np.array_equal(vel, vel0)
    p.write_text(json.dumps({"Pose": pose0.tolist()}))
    _, vel = m.read_state(str(p), vel=np.full(m.num_dof, 7.0))
    assert np.all(vel == 7.0)
    out = tmp_path / "out.txt"
    m.write_state(str(out), pose0, vel0)
    txt = out.read_text()
    assert txt.startswith('{\n"Pose":[') and ("%f" % pose0[1]) in txt
    pose2, vel2 = m.read_state(str(out))
    assert np.allclose(pose2, pose0 / 1.0, atol=1e-6) or np.allclose(pose2[7:], pose0[7:], atol=1e-6)
    p.write_text(json.dumps({"Pose": [1.0, 2.0]}))
    with pytest.raises(trl.TrlError):
        m.read_state(str(p))
    if os.path.isdir(REF):
        ref_pose, ref_vel = m.read_state(os.path.join(REF, "data/states/biped3D/biped3d_stand_state.txt"))
        assert ref_pose.shape == (25,) and abs(np.linalg.norm(ref_pose[3:7]) - 1) < 1e-12


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_env_registry_lists_the_reference_envs():
    """simAdapter/terrainRLSim.py getEnvsList / getEnv over args/envs.json (the adapter itself needs a GPU)."""
    import torch
    from terrainrlsim_b200 import envs
    lst = envs.getEnvsList(REF)
    assert "PD_Biped2D_Flat_Terrain-v0" in lst and lst["PD_Biped2D_Flat_Terrain-v0"]["config_file"].endswith(".txt")
    assert envs.getEnv("no-such-env", root=REF) is None
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            envs.getEnv("PD_Biped2D_Flat_Terrain-v0", root=REF, num_envs=4)
