#!/usr/bin/env python3
"""Run the stable-PD entry point alone (device pointers) -- the target of the ncu captures of k_pd_block.
usage: tools/run_pd.py [config] [envs] [iters]"""
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import terrainrlsim_b200 as trl  # noqa: E402
from terrainrlsim_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "biped3d"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
model = trl.CharModel.from_name(name)
batch = trl.Batch(model, E, 0)
x = synth.make_inputs(name, E)
dev = {k: torch.as_tensor(v).cuda() for k, v in x.items() if k in ("pose", "vel", "tar_pose", "tar_vel", "joint_active")}
tau = torch.zeros((E, model.num_dof), dtype=torch.float64, device="cuda")
batch.use_torch_stream()
for _ in range(3):
    batch.imp_pd_update_control_force(1.0 / 600, dev["pose"], dev["vel"], dev["tar_pose"], dev["tar_vel"],
                                      joint_active=dev["joint_active"], tau=tau)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    batch.imp_pd_update_control_force(1.0 / 600, dev["pose"], dev["vel"], dev["tar_pose"], dev["tar_vel"],
                                      joint_active=dev["joint_active"], tau=tau)
b.record()
torch.cuda.synchronize()
print("%s E=%d: %.2f us per stable-PD call" % (name, E, a.elapsed_time(b) / iters * 1e3))
