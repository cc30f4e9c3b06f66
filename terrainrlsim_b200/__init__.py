"""terrainrlsim_b200 -- B200-native batched control-and-observation step of TerrainRLSim.

Only what the hot path needs: csrc/ (CUDA kernels + C ABI, built in-tree into lib/libtrl_b200.so),
api.py (host mirror of the reference interface over the C ABI), synth.py (deterministic synthetic inputs)
and data/ (flat model matrices of the five BASELINE.json configs).
"""
from .api import (Batch, CharModel, TrlError, make_obs_cfg, terrain_cfg, terrain_load, GROUND_NONE, GROUND_PLANE, GROUND_VAR2D, GROUND_VAR3D,
                  PTR_DEVICE, PTR_HOST)

__all__ = ["Batch", "CharModel", "TrlError", "make_obs_cfg", "terrain_cfg", "terrain_load", "GROUND_NONE", "GROUND_PLANE", "GROUND_VAR2D",
           "GROUND_VAR3D", "PTR_DEVICE", "PTR_HOST"]
