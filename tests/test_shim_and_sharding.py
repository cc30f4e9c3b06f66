"""C++ shim compiles/links against the C ABI; env sharding across ranks (gloo, world_size 2) matches bench.py's."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from terrainrlsim_b200 import build, synth


def _compile_shim(tmp_path):
    build.build()
    exe = str(tmp_path / "shim_smoke")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "shim_smoke.cpp"), "-o", exe,
                           "-L", build.LIB_DIR, "-ltrl_b200", "-Wl,-rpath," + build.LIB_DIR])
    return exe


CHAR = {"Skeleton": {"Joints": [{"Type": "planar", "Parent": -1},
                                {"Type": "revolute", "Parent": 0, "AttachY": -0.2, "TorqueLim": 100}]},
        "BodyDefs": [{"Shape": "box", "Mass": 5, "Param0": 0.4, "Param1": 0.2, "Param2": 0.2},
                     {"Shape": "box", "Mass": 1, "AttachY": -0.2, "Param0": 0.1, "Param1": 0.4, "Param2": 0.1}],
        "PDControllers": [{"Kp": 0, "Kd": 0}, {"Kp": 100, "Kd": 10}]}


def test_cpp_shim_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _compile_shim(tmp_path)
    f = tmp_path / "char.txt"
    f.write_text(json.dumps(CHAR))
    r = subprocess.run([exe, str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "joints=2 dofs=8" in r.stdout
    if not torch.cuda.is_available():
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_shim_runs_on_gpu(tmp_path):
    exe = _compile_shim(tmp_path)
    f = tmp_path / "char.txt"
    f.write_text(json.dumps(CHAR))
    r = subprocess.run([exe, str(f), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok sum|tau|" in r.stdout


def _adapter_files(tmp_path):
    """a two-joint character, a flat terrain file and the arg file that points at them (paths relative to tmp_path)"""
    exe = str(tmp_path / "adapter_smoke")
    build.build()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "adapter_smoke.cpp"), "-o", exe,
                           "-L", build.LIB_DIR, "-ltrl_b200", "-Wl,-rpath," + build.LIB_DIR])
    os.makedirs(tmp_path / "data", exist_ok=True)
    (tmp_path / "data" / "char.txt").write_text(json.dumps(CHAR))
    (tmp_path / "data" / "terrain.txt").write_text(json.dumps({"Type": "var2d_gaps", "Params": [{"GapWMax": 1.0}]}))
    (tmp_path / "args.txt").write_text("-character_file= data/char.txt\n-terrain_file= data/terrain.txt\n"
                                       "-char_ctrl= dog\n-num_update_steps= 20\n-world_scale= 4\n")
    return exe, str(tmp_path / "args.txt"), str(tmp_path) + "/"


def test_cpp_batched_sim_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe, args, rel = _adapter_files(tmp_path)
    r = subprocess.run([exe, args, rel, "8", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_batched_sim_adapter_runs_on_gpu(tmp_path):
    exe, args, rel = _adapter_files(tmp_path)
    r = subprocess.run([exe, args, rel, "8", "2", "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok obs=207" in r.stdout  # 200 ground samples + pose (2 N - 1 = 3) + vel (2 N = 4) for N = 2


WORKER = r"""
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from terrainrlsim_b200 import synth
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
E = 64                                   # envs per rank: weak scaling, disjoint seeds per rank, no data-path collective
x = synth.make_inputs("biped3d", E, rank=rank)
csum = torch.tensor([float(np.abs(x["pose"]).sum()), float(E)], dtype=torch.float64)
lst = [torch.zeros_like(csum) for _ in range(world)]
dist.all_gather(lst, csum)               # the only collective: a per-rank statistics record, outside the hot path
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX) # bench.py's max-over-ranks timing reduction
if rank == 0:
    print(json.dumps({"sums": [float(v[0]) for v in lst], "envs": sum(float(v[1]) for v in lst), "tmax": float(t)}))
dist.destroy_process_group()
"""


def test_env_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["envs"] == 128 and d["tmax"] == 2.0
    # each rank generated its own shard (different seeds), identical to a direct call
    for rank in range(2):
        x = synth.make_inputs("biped3d", 64, rank=rank)
        assert d["sums"][rank] == pytest.approx(float(np.abs(x["pose"]).sum()), rel=0, abs=0)
    assert d["sums"][0] != d["sums"][1]
