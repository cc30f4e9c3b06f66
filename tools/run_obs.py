#!/usr/bin/env python3
"""Run the observation entry point (trl_build_poli_state: k_obs_env + k_obs_samples) alone, device pointers -- the
target of the ncu captures of k_obs_samples.
usage: tools/run_obs.py [config] [envs] [iters] [fields]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import terrainrlsim_b200 as trl  # noqa: E402
from terrainrlsim_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "biped3d"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
fields = int(sys.argv[4]) if len(sys.argv) > 4 else 256
model = trl.CharModel.from_name(name)
batch = trl.Batch(model, E, 0, trl.make_obs_cfg(**synth.obs_config(name)))
kind, fl = synth.make_grounds(name, min(E, fields))
batch.ground_set_heightfields(kind, fl)
x = synth.make_inputs(name, E)
dev = {k: torch.as_tensor(v).cuda() for k, v in x.items()}
batch.use_torch_stream()


def call():
    return batch.build_poli_state(dev["pose"], dev["body_pos"], dev["body_vel"], vel=dev["vel"], phase=dev.get("phase"),
                                  flip=dev.get("flip"))


for _ in range(3):
    call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    call()
b.record()
torch.cuda.synchronize()
print("%s E=%d: %.2f us per observation call" % (name, E, a.elapsed_time(b) / iters * 1e3))
