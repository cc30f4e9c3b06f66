#!/usr/bin/env python3
"""Phase timing of one block of k_pd_aba (diagnostic build with -DTRL_PD_CLK, loaded through TRL_LIB).
build: python -c "from terrainrlsim_b200 import build; build.build_diagnostic()"
usage: TRL_LIB=terrainrlsim_b200/lib/libtrl_b200_clk.so tools/pd_phases.py [config] [envs]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import terrainrlsim_b200 as trl  # noqa: E402
from terrainrlsim_b200 import synth, _capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "biped3d"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
model = trl.CharModel.from_name(name)
batch = trl.Batch(model, E, 0)
x = synth.make_inputs(name, E)
dev = {k: torch.as_tensor(v).cuda() for k, v in x.items() if k in ("pose", "vel", "tar_pose", "tar_vel", "joint_active")}
tau = torch.zeros((E, model.num_dof), dtype=torch.float64, device="cuda")
batch.use_torch_stream()
for _ in range(5):
    batch.imp_pd_update_control_force(1.0 / 600, dev["pose"], dev["vel"], dev["tar_pose"], dev["tar_vel"],
                                      joint_active=dev["joint_active"], tau=tau)
torch.cuda.synchronize()
buf = np.zeros((8, 32), dtype=np.int64)
f = _capi.lib().trl_debug_pd_clk
f.argtypes = [C.c_void_p]
assert f(buf.ctypes.data) == 0
names = {0: "start", 1: "staged", 2: "root down", 3: "barrier", 16: "walk1 end", 17: "barrier", 18: "root solve", 19: "barrier",
         10: "  hip: local trans", 11: "  hip: T V A", 12: "  hip: body f", 14: "  hip: world tar", 13: "  hip: quat_vel_rel", 28: "walk2 end", 29: "barrier", 30: "end"}
t0 = buf[0, 0]
for w in range(8):
    if buf[w, 0] == 0:
        continue
    prev = buf[w, 0]
    print("warp %d" % w)
    for i in range(32):
        if buf[w, i] == 0:
            continue
        print("  %2d %-12s t=%7d cycles (+%6d)" % (i, names.get(i, "event" if i < 16 else "joint"), buf[w, i] - t0, buf[w, i] - prev))
        prev = buf[w, i]
