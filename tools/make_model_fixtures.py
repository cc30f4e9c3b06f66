#!/usr/bin/env python3
"""Build the compact character-model fixtures shipped in terrainrlsim_b200/data/.

Run in the build container (where /root/reference exists):

    python tools/make_model_fixtures.py

For each BASELINE.json config it reads the reference's *input* data files
(character JSON, motion JSON, state JSON; data/characters, data/motions,
data/states) and re-encodes them as the flat numeric matrices the reference
builds from them at load time:

  joint_mat [N][19]  cKinTree::Load / ParseJoint / PostProcessJointMat
                     (anim/KinTree.cpp:451-499,980-1041; column enum KinTree.h:24-46)
  body_defs [N][17]  cKinTree::LoadBodyDefs / ParseBodyDef
                     (anim/KinTree.cpp:127-199; column enum KinTree.h:57-77)
  pd_params [N][18]  cPDController::LoadParams / ParsePDParams
                     (sim/PDController.cpp:29-91; column enum PDController.h:12-32)
  motion frames [F][1+n] with col 0 = frame duration (as stored on disk; the
                     cumulative-time conversion of cMotion::PostProcessFrames,
                     anim/Motion.cpp:207-217, is done by the loader), and Loop
  state Pose/Vel [n] cCharacter::ReadState (anim/Character.cpp:277-342)

The output is what `trl_model_create()` consumes (it mirrors
cRBDModel::Init(joint_mat, body_defs, gravity), sim/RBDModel.cpp:16-37), so the
GPU box (which has no /root/reference) can build every model. The product's own
JSON loader `trl_model_load()` parses the reference file format directly and is
tested against these matrices.
"""
import json
import math
import os
import sys

REF = os.environ.get("TRL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "terrainrlsim_b200", "data")

JOINT_TYPES = ["revolute", "planar", "prismatic", "fixed", "spherical", "none"]
JOINT_KEYS = ["Type", "Parent", "AttachX", "AttachY", "AttachZ", "AttachThetaX", "AttachThetaY",
              "AttachThetaZ", "LimLow0", "LimLow1", "LimLow2", "LimHigh0", "LimHigh1", "LimHigh2",
              "TorqueLim", "ForceLim", "IsEndEffector", "DiffWeight", "Offset"]
BODY_KEYS = ["Shape", "Mass", "ColGroup", "EnableFallContact", "AttachX", "AttachY", "AttachZ",
             "AttachThetaX", "AttachThetaY", "AttachThetaZ", "Param0", "Param1", "Param2",
             "ColorR", "ColorG", "ColorB", "ColorA"]
PD_KEYS = ["JointID", "Kp", "Kd"] + ["TargetTheta%d" % i for i in range(7)] + \
          ["TargetVel%d" % i for i in range(7)] + ["UseWorldCoord"]
SHAPES = {"box": 0, "capsule": 1, "null": 3}
PARAM_SIZE = {0: 1, 1: 3, 2: 1, 3: 0, 4: 4}

INF = float("inf")


def is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def joint_row(j):
    row = [0.0, -1.0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, INF, INF, 0, 1, 0]
    row = [float(x) for x in row]
    row[0] = float(JOINT_TYPES.index(j["Type"])) if "Type" in j else 5.0
    for i, k in enumerate(JOINT_KEYS):
        if i != 0 and k in j and j[k] is not None:
            row[i] = float(j[k])
    return row


def body_row(b):
    row = [0.0] * 17
    row[2] = -1.0
    row[16] = 1.0
    row[0] = float(SHAPES[b["Shape"]])
    for i, k in enumerate(BODY_KEYS):
        if k in b and is_num(b[k]):
            row[i] = float(b[k])
    return row


def pd_row(p, idx):
    row = [0.0] * 18
    for i, k in enumerate(PD_KEYS):
        if k in p and is_num(p[k]):
            row[i] = float(p[k])
    row[0] = float(idx)
    return row


def build(name, char_file, motion_file, state_file, extra):
    c = json.load(open(os.path.join(REF, char_file)))
    joints = [joint_row(j) for j in c["Skeleton"]["Joints"]]
    off = 0
    for j, row in enumerate(joints):
        size = 7 if int(row[1]) == -1 else PARAM_SIZE[int(row[0])]
        row[18] = float(off)
        off += size
    joints[0][2] = joints[0][3] = joints[0][4] = 0.0
    bodies = [body_row(b) for b in c["BodyDefs"]]
    pds = [pd_row(p, i) for i, p in enumerate(c["PDControllers"])]
    out = {"name": name, "source": {"character": char_file, "motion": motion_file, "state": state_file},
           "num_joints": len(joints), "num_dof": off,
           "joint_names": [j.get("Name", "") for j in c["Skeleton"]["Joints"]],
           "joint_mat": joints, "body_defs": bodies, "pd_params": pds}
    if motion_file:
        m = json.load(open(os.path.join(REF, motion_file)))
        out["motion"] = {"loop": bool(m.get("Loop", False)), "frames": m["Frames"]}
        assert len(m["Frames"][0]) == off + 1
    if state_file:
        s = json.load(open(os.path.join(REF, state_file)))
        out["state"] = {"pose": s["Pose"], "vel": s["Vel"]}
        assert len(s["Pose"]) == off
    out.update(extra)
    if extra.get("terrain_file"):
        # the terrain parameter file as the product's own loader reads it (cGroundFactory::ParseParamsJson), blend 0
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
        import terrainrlsim_b200 as trl
        tc = trl.terrain_load(os.path.join(REF, extra["terrain_file"]))
        out["terrain_cfg"] = {"type": int(tc.type), "ground_width": tc.ground_width, "spacing_x": tc.spacing_x,
                              "spacing_z": tc.spacing_z, "params": [float(v) for v in tc.params]}

    def enc(o):
        if isinstance(o, float) and math.isinf(o):
            return "Infinity"
        raise TypeError

    txt = json.dumps(out, indent=None, separators=(",", ":"))
    # json.dumps writes Infinity literally; keep it (python json reads it back)
    path = os.path.join(OUT, name + ".json")
    with open(path, "w") as f:
        f.write(txt + "\n")
    print("wrote", path, "N=%d n=%d" % (len(joints), off))


CONFIGS = [
    # name, character, motion, state, extras (observation config per SURVEY.md section 8 table)
    ("biped2d", "data/characters/biped2D_full.txt", "data/motions/biped2D_full_walk.txt",
     "data/states/biped2D_full_walk_state.txt",
     {"ctrl": "ct_pd_phase", "obs": "ct", "ground_samples": 0, "terrain": "var2d_flat", "terrain_file": "data/terrain/flat.txt",
      "num_update_steps": 20, "envs_per_gpu": 4096}),
    ("hopper", "data/characters/monoped_hopper0.txt", "data/motions/monoped_hopper.txt",
     "data/states/monoped_hopper.txt",
     {"ctrl": "monoped_hopper", "obs": "terrain_rl_hopper", "ground_samples": 200, "terrain": "var2d_mixed",
      "terrain_file": "data/terrain/mixed_hopper.txt",
      "num_update_steps": 80, "envs_per_gpu": 4096}),
    ("dog", "data/characters/dog.txt", "data/motions/dog_bound.txt", "data/states/dog_bound_state3.txt",
     {"ctrl": "dog", "obs": "terrain_rl", "ground_samples": 200, "terrain": "var2d_mixed", "terrain_file": "data/terrain/mixed.txt",
      "num_update_steps": 20, "envs_per_gpu": 4096}),
    ("raptor", "data/characters/raptor.txt", "data/motions/raptor_run.txt", "data/states/raptor_run_state.txt",
     {"ctrl": "raptor", "obs": "terrain_rl", "ground_samples": 200, "terrain": "var2d_flat", "terrain_file": "data/terrain/flat.txt",
      "num_update_steps": 20, "envs_per_gpu": 8192}),
    ("biped3d", "data/characters/biped_3D.txt", "data/motions/biped_3D_walk.txt",
     "data/states/biped3D/biped3d_stand_state.txt",
     {"ctrl": "biped3D_cacla", "obs": "terrain_rl_grid", "ground_samples": 256, "terrain": "var3d_flat",
      "terrain_file": "data/terrain/flat3d.txt",
      "num_update_steps": 20, "envs_per_gpu": 16384}),
]

if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    for cfg in CONFIGS:
        build(*cfg)
