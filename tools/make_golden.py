#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container (where /root/reference exists):

    make -C oracle ref && python tools/make_golden.py

Every output array in the fixtures is computed by the reference's own, unmodified translation units
(oracle/_ref/libtrl_ref.so, see oracle/Makefile and oracle/ref_wrap.cpp) on seeded synthetic inputs;
nothing here calls the C oracle or the product.  tests/test_golden.py then checks the oracle (CPU) and
the CUDA path (GPU, through the C-ABI) against these files -- on the GPU box too, where the reference
tree does not exist.

Files: tests/golden/ref_<config>.npz (dynamics / kinematics / stable PD / kin character),
       tests/golden/ref_synthetic.npz (planar + fixed joints, null body),
       tests/golden/ref_ground_<terrain>.npz (reference-generated terrain pieces + sampled heights).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_py  # noqa: E402
from terrainrlsim_b200 import synth  # noqa: E402
from helpers import synthetic_character, synthetic_state  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = ref_py.REF_ROOT
CONFIGS = ["biped2d", "hopper", "dog", "raptor", "biped3d"]
E = 8


def dynamics_block(r, N, n, pose, vel, tar_pose, tar_vel, active, dts, kp=None, kd=None):
    out = {"M": [], "C": [], "tau": [], "J": [], "Jcom": [], "g": [], "joint_world": [], "com": [], "com_vel": [],
           "pose_diff": [], "pose_err": []}
    for e in range(pose.shape[0]):
        tau, M, C = r.imp_pd(dts[e], pose[e], vel[e], tar_pose[e], tar_vel[e], active[e],
                             None if kp is None else kp[e], None if kd is None else kd[e])
        out["M"].append(M)
        out["C"].append(C)
        out["tau"].append(tau)
        J, Jc, g = r.rbd_extras(pose[e])
        out["J"].append(J)
        out["Jcom"].append(Jc)
        out["g"].append(g)
        out["joint_world"].append(np.stack([r.joint_world_trans(pose[e], j) for j in range(N)]))
        com, cv = r.calc_com(pose[e], vel[e])
        out["com"].append(com)
        out["com_vel"].append(cv)
        out["pose_diff"].append(r.vel_to_pose_diff(pose[e], vel[e]))
        out["pose_err"].append(r.calc_vel(pose[e], pose[(e + 1) % pose.shape[0]], 1.0))
    return {k: np.array(v) for k, v in out.items() if v}


def make_config(name):
    d = synth.load_model_dict(name)
    jm, bd, pd = ref_py.load_character(os.path.join(REF, d["source"]["character"]))
    x = synth.make_inputs(name, E)
    rng = np.random.default_rng(20261020)
    fx = {"joint_mat": jm, "body_defs": bd, "pd_params": pd}
    for k in ("pose", "vel", "tar_pose", "tar_vel", "joint_active"):
        fx[k] = x[k]
    fx["dt"] = np.array([1.0 / 600, 1.0 / 2400, 1.0 / 30, 1.0 / 600] * (E // 4))
    fx["kp"] = rng.uniform(10, 500, (E, jm.shape[0]))
    fx["kd"] = rng.uniform(1, 50, (E, jm.shape[0]))
    for world in (0, 1):
        pdw = pd.copy()
        if not world:
            pdw[:, 17] = 0
        r = ref_py.RefModel(jm, bd, pdw)
        blk = dynamics_block(r, r.N, r.n, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"], x["joint_active"], fx["dt"])
        if world:
            fx["tau_world_targets"] = blk["tau"]
        else:
            fx.update(blk)
            fx["tau_gains"] = np.array([r.imp_pd(fx["dt"][e], x["pose"][e], x["vel"][e], x["tar_pose"][e],
                                                 x["tar_vel"][e], x["joint_active"][e], fx["kp"][e], fx["kd"][e])[0]
                                        for e in range(E)])
            fx["exp_tau"] = np.array([r.exp_pd(fx["dt"][e], x["pose"][e], x["vel"][e], x["tar_pose"][e],
                                               x["tar_vel"][e], x["joint_active"][e]) for e in range(E)])
            fx["end_eff_J"] = np.array([[r.end_eff_jacobian(x["pose"][e], j) for j in range(r.N)] for e in range(2)])
            fx["body_pos"] = np.array([[r.body_part_pos(x["pose"][e], j) if bd[j, 0] != 3 else np.zeros(3)
                                        for j in range(r.N)] for e in range(E)])
            fx["lerp"] = np.array([r.lerp_poses(x["pose"][e], x["pose"][(e + 3) % E], 0.1 + 0.1 * e) for e in range(E)])
            fx["heading"] = np.array([r.calc_heading(x["pose"][e]) for e in range(E)])
            fx["origin_trans"] = np.array([r.build_origin_trans(x["pose"][e]) for e in range(E)])
    # kinematic character on the reference's own character / motion / state files
    kc = ref_py.RefKinChar(os.path.join(REF, d["source"]["character"]), os.path.join(REF, d["source"]["motion"]))
    fx["frames"] = kc.frames()
    fx["loop"] = np.array(int(kc.loop))
    dur = kc.duration
    times = np.concatenate([rng.uniform(0, 10 * dur, 10), [0.0, -0.3 * dur, -2.2 * dur, dur, 3 * dur, 0.5 * dur]])
    origin_pos = rng.normal(size=(times.size, 3))
    ang = rng.uniform(-3, 3, times.size)
    origin_rot = np.stack([np.cos(ang / 2), 0 * ang, np.sin(ang / 2), 0 * ang], axis=1)
    kp_, kv_, idx_, ph_ = [], [], [], []
    for i, t in enumerate(times):
        p, v = kc.pose(t, origin_pos[i], origin_rot[i])
        kp_.append(p)
        kv_.append(v)
        a, b = kc.index_phase(t)
        idx_.append(a)
        ph_.append(b)
    fx.update({"kin_time": times, "kin_origin_pos": origin_pos, "kin_origin_rot": origin_rot,
               "kin_pose": np.array(kp_), "kin_vel": np.array(kv_), "kin_idx": np.array(idx_), "kin_phase": np.array(ph_)})
    sp, sv = kc.read_state(os.path.join(REF, d["source"]["state"]))
    fx["state_pose"], fx["state_vel"] = sp, sv
    np.savez_compressed(os.path.join(OUT, "ref_%s.npz" % name), **fx)
    return fx


def make_synthetic():
    for null_body in (0, 1):
        jm, bd, pd = synthetic_character(with_null_body=bool(null_body))
        r = ref_py.RefModel(jm, bd, pd)
        pose, vel, tar, tar_vel = synthetic_state(jm, E)
        active = np.ones((E, jm.shape[0]), dtype=np.uint8)
        active[::3, 2] = 0
        dts = np.full(E, 1.0 / 600)
        fx = {"joint_mat": jm, "body_defs": bd, "pd_params": pd, "pose": pose, "vel": vel, "tar_pose": tar,
              "tar_vel": tar_vel, "joint_active": active, "dt": dts}
        fx.update(dynamics_block(r, r.N, r.n, pose, vel, tar, tar_vel, active, dts))
        fx["body_pos"] = np.array([[r.body_part_pos(pose[e], j) if bd[j, 0] != 3 else np.zeros(3)
                                    for j in range(r.N)] for e in range(E)])
        fx["end_eff_J"] = np.array([[r.end_eff_jacobian(pose[e], j) for j in range(r.N)] for e in range(2)])
        np.savez_compressed(os.path.join(OUT, "ref_synthetic%d.npz" % null_body), **fx)


def make_ground(terrain, scale, seed, tag):
    rg = ref_py.RefGround(os.path.join(REF, "data", "terrain", terrain), world_scale=scale, seed=seed)
    rng = np.random.default_rng(seed)
    fx = {"kind": np.array(rg.kind), "world_scale": np.array(scale), "seed": np.array(seed)}
    # generator inputs, parsed by the reference's own loader (cGroundFactory::ParseParamsJson as restated in
    # oracle/ref_wrap.cpp + cTerrainGen2D::LoadParams), so the oracle / product generators can be run on them
    import json
    with open(os.path.join(REF, "data", "terrain", terrain)) as f:
        meta = json.load(f)
    fx["type_name"] = np.array(meta["Type"])
    fx["ground_width"] = np.array(float(meta.get("GroundWidth", 20)))
    if rg.kind == 1:
        try:
            fx["params"] = ref_py.terrain_params_2d(os.path.join(REF, "data", "terrain", terrain), 0)
        except RuntimeError:
            fx["params"] = ref_py.terrain_params_2d_default()
    for step in range(2):
        pieces = rg.pieces()
        lo = min(p["bounds"][0] for p in pieces) - 1.0
        hi = max(p["bounds"][1] for p in pieces) + 1.0
        pos = np.zeros((3000, 3))
        pos[:, 0] = rng.uniform(lo, hi, 3000)
        pos[:, 1] = rng.normal(size=3000)
        zlo = min(max(p["bounds"][2], -30) for p in pieces) - 1.0
        zhi = max(min(p["bounds"][3], 30) for p in pieces) + 1.0
        pos[:, 2] = rng.uniform(zlo, zhi, 3000)
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
        pos = np.concatenate([pos, np.array(extra)], axis=0)
        h, valid = rg.sample_height(pos)
        fx["s%d_num_pieces" % step] = np.array(len(pieces))
        for s, p in enumerate(pieces):
            for k in ("data", "origin", "scale", "bounds"):
                fx["s%d_p%d_%s" % (step, s, k)] = p[k]
            fx["s%d_p%d_res" % (step, s)] = np.array(p["res"])
        fx["s%d_pos" % step], fx["s%d_h" % step], fx["s%d_valid" % step] = pos, h, valid
        xs = max(p["bounds"][1] for p in pieces)
        fx["s%d_update_min" % step] = np.array([xs - 1.0, 0, -1.0])
        fx["s%d_update_max" % step] = np.array([xs + 1.0, 0, 1.0])
        rg.update([xs - 1.0, 0, -1.0], [xs + 1.0, 0, 1.0])
    np.savez_compressed(os.path.join(OUT, "ref_ground_%s.npz" % tag), **fx)


def main():
    if not ref_py.have_reference_tree():
        sys.exit("the reference tree is needed to generate golden fixtures")
    os.makedirs(OUT, exist_ok=True)
    for name in CONFIGS:
        make_config(name)
    make_synthetic()
    make_ground("mixed.txt", 1.0, 1234, "mixed")
    make_ground("mixed_hopper.txt", 4.0, 99, "mixed_hopper_ws4")
    make_ground("flat.txt", 1.0, 5, "flat")
    make_ground("flat3d.txt", 1.0, 7, "flat3d")
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote %d files, %.1f KB" % (len(os.listdir(OUT)), total / 1024.0))


if __name__ == "__main__":
    main()
