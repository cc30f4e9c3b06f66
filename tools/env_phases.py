#!/usr/bin/env python3
"""Phase timing of one block of k_obs_env (diagnostic build with -DTRL_PD_CLK, loaded through TRL_LIB)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from terrainrlsim_b200 import _capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "biped3d"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    wl = bench.Workload(name, E, 0, 0)
    wl.run_steps(6)  # fused steps: k_obs_env runs first on its stream
    torch.cuda.synchronize()
buf = np.zeros(16, dtype=np.int64)
f = _capi.lib().trl_debug_env_clk
f.argtypes = [C.c_void_p]
assert f(buf.ctypes.data) == 0
names = ["start", "model idx staged", "heading / origin", "ground height", "rec built + stored", "pose features", "vel features"]
for i, nme in enumerate(names):
    print("%-20s t=%7d (+%6d)" % (nme, buf[i] - buf[0], buf[i] - buf[max(i - 1, 0)]))
