#!/usr/bin/env python3
"""Kernel timeline of one fused step (trl_debug_trace_*): start / end of every kernel relative to the first start.
usage: python tools/trace_step.py [config] [envs] [--graph]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from terrainrlsim_b200 import _capi  # noqa: E402

NAMES = ["-", "k_pd_aba/block", "k_obs_feat", "-", "k_kin_tile", "k_obs_rec", "k_obs_samples/gather"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "biped3d"
    E = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 16384
    use_graph = "--graph" in sys.argv
    drop = [a[len("--drop="):] for a in sys.argv if a.startswith("--drop=")]  # pd | kin | obs: leave a chain out
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        wl = bench.Workload(name, E, 0, 0, sets=1 if use_graph else None)
        for d in drop:
            keys = {"pd": ["tau"], "kin": ["kin_pose", "kin_vel", "kin_body_pos"], "obs": ["state"]}[d]
            for i in range(wl.sets):
                out = {k: v for k, v in wl.outputs[i].items() if k not in keys}
                wl.outputs[i] = out
                wl.args[i] = wl.batch.make_step_args(wl.dt, wl.inputs[i], out)
        if use_graph:
            wl.capture()
        wl.run_steps(8)
        torch.cuda.synchronize()
        lib = _capi.lib()
        assert lib.trl_debug_trace_enable(1) == 0
        buf = (C.c_ulonglong * 32)()
        host = "--host" in sys.argv  # HOST pointer mode: pinned host buffers, sliced pipeline; spans cover all slices
        if host:
            import time
            hin = {k: torch.from_numpy(v).pin_memory() for k, v in wl.host_inputs.items()}
            n, N, F = wl.model.num_dof, wl.model.num_joints, wl.batch.F
            hout = {"tau": torch.empty((E, n), dtype=torch.float64).pin_memory(),
                    "kin_pose": torch.empty((E, n), dtype=torch.float64).pin_memory(),
                    "kin_vel": torch.empty((E, n), dtype=torch.float64).pin_memory(),
                    "kin_body_pos": torch.empty((E, N, 3), dtype=torch.float64).pin_memory(),
                    "state": torch.empty((E, F), dtype=torch.float64).pin_memory()}
            hargs = wl.batch.make_step_args(wl.dt, {k: t.numpy() for k, t in hin.items()}, {k: t.numpy() for k, t in hout.items()})
            wl.batch.set_stream(None)
            for _ in range(3):
                wl.batch.step_fused_args(hargs)
        for rep in range(4):
            if host:
                t0w = time.perf_counter()
                wl.batch.step_fused_args(hargs)
                print("host-mode step wall time %.1f us" % ((time.perf_counter() - t0w) * 1e6))
            elif use_graph:
                wl.graph.replay()
            else:
                wl.step()
            torch.cuda.synchronize()
            assert lib.trl_debug_trace_read(buf) == 0
            rows = [(NAMES[i], buf[2 * i], buf[2 * i + 1]) for i in range(len(NAMES)) if buf[2 * i + 1] != 0]
            t0 = min(r[1] for r in rows)
            print("step %d%s" % (rep, " (graph of %d steps: spans cover all of them)" % wl.sets if use_graph else ""))
            for nm, a, b in sorted(rows, key=lambda r: r[1]):
                print("  %-14s start %8.1f us   end %8.1f us   (%.1f us)" % (nm, (a - t0) / 1e3, (b - t0) / 1e3, (b - a) / 1e3))
        lib.trl_debug_trace_enable(0)


if __name__ == "__main__":
    main()
