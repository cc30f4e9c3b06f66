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
