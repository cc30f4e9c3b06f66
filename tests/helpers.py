"""Shared helpers for the parity tests: oracle <-> product plumbing and the tolerance definition."""
import numpy as np

from terrainrlsim_b200 import synth

RTOL = 1e-9  # north_star: "within 1e-9 relative on torques and features"


def rel_err(a, b, axis=-1):
    """max |a-b| per vector, relative to that vector's max magnitude in the oracle (floor 1: features and
    torques are O(1..1e3); a pure element-wise relative error is meaningless for entries that cancel to ~0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    same_nan = np.isnan(a) & np.isnan(b)
    d = np.where(same_inf | same_nan, 0.0, np.abs(a - b))
    d = np.where(np.isnan(d), np.inf, d)
    fin = np.where(np.isfinite(b), np.abs(b), 0.0)
    scale = np.maximum(fin.max(axis=axis, keepdims=True), 1.0)
    return float((d / scale).max()) if d.size else 0.0


def assert_close(a, b, what, rtol=RTOL):
    e = rel_err(a, b)
    assert e <= rtol, "%s: max relative error %.3e > %.1e" % (what, e, rtol)
    return e


def oracle_cfg(oracle, name):
    c = synth.obs_config(name)
    return oracle.obs_cfg(**c)


def product_cfg(trl, name):
    c = synth.obs_config(name)
    return trl.make_obs_cfg(**c)


def oracle_grounds(oracle, kind, fields):
    return [oracle.Ground(kind, pieces) for pieces in fields]


def synthetic_character(seed=0, with_null_body=False):
    """A character that exercises every joint type on non-root joints (the bundled characters only have revolute,
    prismatic and spherical ones): root -> revolute -> PLANAR -> FIXED -> prismatic, root -> spherical -> planar,
    and a fixed leaf. Flat tables in the reference's column layout (anim/KinTree.h:24-77, sim/PDController.h:12-32)."""
    rng = np.random.default_rng(seed)
    REV, PLA, PRI, FIX, SPH = 0, 1, 2, 3, 4
    types = [REV, REV, PLA, FIX, PRI, SPH, PLA, FIX]
    parents = [-1, 0, 1, 2, 3, 0, 5, 6]
    sizes = {REV: 1, PLA: 3, PRI: 1, FIX: 0, SPH: 4}
    N = len(types)
    jm = np.zeros((N, 19))
    bd = np.zeros((N, 17))
    pd = np.zeros((N, 18))
    off = 0
    for j in range(N):
        size = 7 if parents[j] < 0 else sizes[types[j]]
        jm[j, 0] = types[j]
        jm[j, 1] = parents[j]
        if parents[j] >= 0:
            jm[j, 2:5] = rng.uniform(-0.4, 0.4, 3)
            jm[j, 5:8] = rng.uniform(-0.8, 0.8, 3)
        jm[j, 8:11] = 1
        jm[j, 11:14] = 0
        jm[j, 14] = rng.uniform(20, 200)   # torque limit
        jm[j, 15] = rng.uniform(20, 200)   # force limit
        jm[j, 16] = 1 if j in (4, 7) else 0
        jm[j, 17] = rng.uniform(0.5, 1.5)
        jm[j, 18] = off
        off += size
        capsule = (j % 3 == 1)
        bd[j, 0] = 1 if capsule else 0
        bd[j, 1] = rng.uniform(0.5, 5.0)
        bd[j, 2] = -1
        bd[j, 4:7] = rng.uniform(-0.2, 0.2, 3)
        bd[j, 7:10] = rng.uniform(-0.5, 0.5, 3)
        bd[j, 10:13] = rng.uniform(0.05, 0.4, 3)
        bd[j, 16] = 1
        pd[j, 0] = j
        if parents[j] >= 0:
            pd[j, 1] = rng.uniform(50, 400)
            pd[j, 2] = rng.uniform(5, 40)
    if with_null_body:
        bd[7, 0] = 3   # shape "null": an invalid body (leaf)
        bd[7, 1] = 0
    return jm, bd, pd


def synthetic_state(jm, E, seed=1):
    """random pose / vel / targets for a flat joint table (unit quaternions, dead velocity slots zero)"""
    rng = np.random.default_rng(seed)
    N = jm.shape[0]
    sizes = {0: 1, 1: 3, 2: 1, 3: 0, 4: 4}
    n = int(jm[-1, 18] + (7 if jm[-1, 1] < 0 else sizes[int(jm[-1, 0])]))
    pose = rng.normal(0, 0.4, (E, n))
    vel = rng.normal(0, 1.0, (E, n))
    tar = pose + rng.normal(0, 0.2, (E, n))
    quats = [3] + [int(jm[j, 18]) for j in range(1, N) if int(jm[j, 0]) == 4]
    for q in quats:
        for a in (pose, tar):
            a[:, q:q + 4] = rng.normal(size=(E, 4))
            a[:, q:q + 4] /= np.linalg.norm(a[:, q:q + 4], axis=1, keepdims=True)
    vel[:, 6] = 0
    for q in quats[1:]:
        vel[:, q + 3] = 0
    tar[:, :7] = 0
    return pose, vel, tar, np.zeros((E, n))
