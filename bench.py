#!/usr/bin/env python3
"""bench.py -- env-steps/sec of the batched stable-PD + FK/mocap + terrain + feature-pack step.

    python bench.py --gpus N --steps K --warmup W [--config biped3d] [--impl reference]

One "step" = one pass of the fused hot path (trl_step_fused) over the E environments of one GPU:
stable-PD torque, kin-char pose+vel, FK of the kin pose, G terrain samples, F packed features.
Default workload: biped3D, 16384 envs/GPU on 4096 terrains generated on the device from the config's own terrain
file -- the configuration BASELINE.json's north_star target is quoted on; the other configs run in the same
invocation on every rank and are reported under "other_workloads".

value  : whole-job env-steps/s with every input already resident in HBM (DEVICE pointer mode, CUDA graph replay,
         CUDA events, max over ranks). L2 is defeated by rotating over R input/output sets (> 2.5x L2 in total).
e2e    : same metric through the HOST-pointer C ABI (trl_step_fused_ex, pinned host buffers, synchronous calls; H2D of
         the step's inputs and D2H of torque, f64 state and reward inside the timed region), see measure_e2e();
         e2e.variants / e2e.pipelined report other transfer contracts beside it.
roofline: dominant kernel (stable PD = k_pd_aba, CUDA-event-timed alone on the same rotating sets) -- algorithmic FP64
         flops per launch (SURVEY.md 8d table x E) over its time, against the FP64 DFMA peak measured in this run
         (MEASURED_PEAKS.json has no FP64 entry); `traffic` = its DRAM bytes from the committed ncu capture
         (profiles/r02_pd_traffic.json); roofline_step_hbm does the same for the whole step's algorithmic bytes against
         the measured HBM copy peak.
cpu_baseline / --impl reference: the CPU oracle (a restatement of the reference's algorithms pinned to the reference's
         own translation units, tests/test_ref_pin.py) on all host cores, on the same workload: one step = one pass
         over a GPU's worth of envs on the same number of terrains, built from the same seeds.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stable-PD+FK+terrain env-steps/sec"
UNIT = "env-steps/s"
L2_BYTES = 126 * 1024 * 1024

# SURVEY.md section 8(d) cost model (structure-exploiting count), per env-step
ROOFLINE_MODEL = {
    # name: (PD+RBD flops, total flops, algorithmic bytes)
    "biped2d": (12.6e3, 14.2e3, 1.56e3),
    "hopper": (3.3e3, 7.8e3, 3.7e3),
    "dog": (27.9e3, 34.7e3, 5.7e3),
    "raptor": (23.0e3, 29.5e3, 5.5e3),
    "biped3d": (16.3e3, 33.0e3, 6.5e3),
}
WORKLOAD_NAMES = {
    "biped2d": "genBiped2D imitation (biped2D_full.txt, ct_pd_phase)",
    "hopper": "monopedHopper (monoped_hopper0.txt) on var2d_mixed",
    "dog": "dog 2D (dog.txt) on var2d_mixed",
    "raptor": "raptor 2D (raptor.txt) on var2d_flat",
    "biped3d": "biped3D (biped_3D.txt) 3D stable-PD on var3d_flat",
}


def roofline_model(name):
    pd, total, byt = ROOFLINE_MODEL[name]
    return {"pd_flops_per_env_step": pd, "total_flops_per_env_step": total, "bytes_per_env_step": byt}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def default_fields(name, E, override=0):
    """Distinct terrains per GPU (one per env up to 4096, env % 4096 above)."""
    if override:
        return min(E, override)
    return min(E, 4096)


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (test infrastructure used here only as the measured CPU baseline)
# ----------------------------------------------------------------------------------------------
class CpuArm:
    """The CPU oracle's fused step over `sample_envs` envs, one thread per env range (like cScenarioTrain::Run)."""

    def __init__(self, name, sample_envs, threads, fields=None):
        from oracle import oracle_py
        from terrainrlsim_b200 import synth
        self.oracle = oracle_py
        self.threads = threads
        self.sample = sample_envs
        self.om = oracle_py.Model(name)
        self.cfg = oracle_py.obs_cfg(**synth.obs_config(name))
        self.num_fields = min(sample_envs, fields if fields else 32)
        # the config's own terrain file and generator seeds, as on the GPU arm (the oracle's generator and the device one
        # build bit-identical terrains: tests/test_terrain.py)
        t = synth.terrain_setup(name, self.num_fields)
        if t["type"] == 15:
            self.grounds = [oracle_py.Ground(2, oracle_py.GroundVar3D(t["spacing_x"], t["spacing_z"], t["world_scale"],
                                                                      t["ground_width"]).pieces()) for _ in t["seeds"]]
        else:
            self.grounds = [oracle_py.Ground(1, oracle_py.GroundVar2D(t["type"], t["params"], t["ground_width"], t["world_scale"],
                                                                      int(sd)).pieces()) for sd in t["seeds"]]
        d = synth.load_model_dict(name)
        self.dt = (1.0 / 30) / d["num_update_steps"]
        self.arrays = dict(synth.derive_body_state(name, synth.make_inputs(name, sample_envs)))
        oracle_py.step_range(self.om, self.cfg, self.arrays, self.grounds, self.dt, 0, min(8, sample_envs))  # allocate outputs
        self.bounds = np.linspace(0, sample_envs, threads + 1).astype(int)

    def run(self, passes):
        """`passes` passes over the sample; returns seconds."""
        def work(lo, hi):
            for _ in range(passes):
                self.oracle.step_range(self.om, self.cfg, self.arrays, self.grounds, self.dt, int(lo), int(hi))

        ts = [threading.Thread(target=work, args=(self.bounds[i], self.bounds[i + 1])) for i in range(self.threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0


def cpu_baseline(name, E, fields, target_seconds=8.0):
    """The oracle on all host cores over the GPU arm's own workload (E envs, the same number of terrains), ~10 s."""
    threads = os.cpu_count() or 1
    arm = CpuArm(name, E, threads, fields=fields)
    t = arm.run(1)
    passes = max(1, int(target_seconds / max(t, 1e-6)))
    t = arm.run(passes)
    val = E * passes / t
    return {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs on %d terrains x %d passes of the C oracle (faithful restatement; faster than the Eigen "
                      "reference: no heap temporaries), one thread per env range, %d threads" % (E, arm.num_fields, passes, threads)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port on all host cores
    (the reference itself cannot be built here: Eigen / Bullet are not vendored). Each step is a bounded sample
    of the workload; under torchrun only rank 0 works and prints."""
    if rank != 0:
        return
    from terrainrlsim_b200 import synth
    name = args.config
    threads = os.cpu_count() or 1
    # the same workload as the GPU arm: one step = one pass over a GPU's worth of envs on the same number of terrains
    sample = args.envs or synth.load_model_dict(name)["envs_per_gpu"]
    arm = CpuArm(name, sample, threads, fields=default_fields(name, sample, args.terrains))
    arm.run(max(args.warmup, 1))
    t = arm.run(args.steps)  # K passes inside persistent worker threads (no per-step thread start in the timed region)
    val = sample * args.steps / t
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD_NAMES[name], "envs_per_gpu": sample, "terrains_per_gpu": arm.num_fields,
                       "sample": "one GPU's worth of envs per step on the host cores"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "each step = one oracle pass over %d envs (bounded sample of the workload), "
                                       "%d threads" % (sample, threads)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s in sm if s >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


class Workload:
    """One config on one GPU: model, batch, terrain, R rotating device-resident input/output sets."""

    def __init__(self, name, E, device, rank, sets=None, fields=None):
        import torch
        import terrainrlsim_b200 as trl
        from terrainrlsim_b200 import synth
        self.name, self.E, self.device = name, E, device
        self.d = synth.load_model_dict(name)
        self.dt = (1.0 / 30) / self.d["num_update_steps"]
        self.model = trl.CharModel.from_name(name)
        self.batch = trl.Batch(self.model, E, device, trl.make_obs_cfg(**synth.obs_config(name)))
        if fields is None:
            fields = default_fields(name, E)
        self.num_fields = fields
        # one procedural terrain per field, generated on the device from the config's terrain parameter file
        t = synth.terrain_setup(name, fields)
        self.batch.ground_generate(trl.terrain_cfg(t["type"], t["params"], t["ground_width"], t["world_scale"], t["spacing_x"],
                                                   t["spacing_z"]), fields, seeds=t["seeds"])
        n, N, F = self.model.num_dof, self.model.num_joints, self.batch.F
        self.in_bytes = E * (4 * n * 8 + N + 8 + 3 * 8 + 4 * 8 + 2 * N * 3 * 8 + 8 + 1)
        self.out_bytes = E * (3 * n * 8 + N * 3 * 8 + F * 8)
        if sets is None:
            sets = max(2, int(np.ceil(2.5 * L2_BYTES / (self.in_bytes + self.out_bytes))))
        self.sets = sets
        dev = torch.device("cuda", device)
        self.host_inputs = synth.derive_body_state(name, synth.make_inputs(name, E, rank=rank))
        self.inputs, self.outputs, self.args = [], [], []
        self.batch.use_torch_stream()
        for s in range(sets):
            x = self.host_inputs if s == 0 else synth.derive_body_state(name, synth.make_inputs(name, E, seed=0xB200 + 7919 * s + rank))
            xin = {k: torch.from_numpy(v).to(dev) for k, v in x.items()}
            out = {"tau": torch.empty((E, n), dtype=torch.float64, device=dev),
                   "kin_pose": torch.empty((E, n), dtype=torch.float64, device=dev),
                   "kin_vel": torch.empty((E, n), dtype=torch.float64, device=dev),
                   "kin_body_pos": torch.empty((E, N, 3), dtype=torch.float64, device=dev),
                   "state": torch.empty((E, F), dtype=torch.float64, device=dev)}
            self.inputs.append(xin)
            self.outputs.append(out)
            self.args.append(self.batch.make_step_args(self.dt, xin, out))
        self.i = 0
        self.graph = None
        self.pd_graph = None
        self.launches_per_step = 0

    def capture(self):
        """Capture one pass over all R sets (R fused steps) into a CUDA graph: the hot loop is launch-bound from
        Python otherwise. Also captures the stable-PD entry point alone for the roofline measurement."""
        import torch
        torch.cuda.synchronize()
        l0 = self.batch.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.batch.use_torch_stream()
            for _ in range(self.sets):
                self.step()
        self.launches_per_step = (self.batch.launch_count() - l0) // self.sets
        self.pd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.pd_graph):
            self.batch.use_torch_stream()
            for _ in range(self.sets):
                self.pd_only()
        self.batch.use_torch_stream()
        torch.cuda.synchronize()

    def run_steps(self, k):
        """exactly k fused steps: whole graph replays (R steps each) + the remainder launched directly"""
        if self.graph is None:
            for _ in range(k):
                self.step()
            return
        for _ in range(k // self.sets):
            self.graph.replay()
        for _ in range(k % self.sets):
            self.step()

    def run_pd(self, k):
        for _ in range(k // self.sets):
            self.pd_graph.replay()

    def step(self):
        rc = self.batch.step_fused_args(self.args[self.i % self.sets])
        self.i += 1
        if rc != 0:
            from terrainrlsim_b200 import _capi
            raise RuntimeError("trl_step_fused failed: " + _capi.last_error())

    def breakdown(self, reps=20):
        """CUDA-event time of each entry point alone (graph of one pass over the R sets, replayed): ms per launch."""
        import torch
        res = {}
        b = self.batch

        def f_pd(x, o):
            b.imp_pd_update_control_force(self.dt, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"],
                                          joint_active=x["joint_active"], tau=o["tau"])

        def f_kin(x, o):
            from terrainrlsim_b200 import _capi
            p = b._prep([(x["kin_time"], np.float64), (x["origin_pos"], np.float64), (x["origin_rot"], np.float64),
                         (o["kin_pose"], np.float64), (o["kin_vel"], np.float64)])
            _capi.lib().trl_kin_char_pose(b._h, *p)

        def f_fk(x, o):
            from terrainrlsim_b200 import _capi
            p = b._prep([(o["kin_pose"], np.float64), (None, np.float64), (o["kin_body_pos"], np.float64)])
            _capi.lib().trl_kin_tree_fk(b._h, *p)

        def f_obs(x, o):
            from terrainrlsim_b200 import _capi
            p = b._prep([(x["pose"], np.float64), (x["vel"], np.float64), (x["body_pos"], np.float64),
                         (x["body_vel"], np.float64), (None, np.float64), (None, np.float64), (None, np.float64),
                         (x["phase"], np.float64), (x["flip"], np.uint8), (o["state"], np.float64), (None, np.int32)])
            _capi.lib().trl_build_poli_state(b._h, *p)

        kin_args = [b.make_step_args(self.dt, self.inputs[s], {k: v for k, v in self.outputs[s].items() if k.startswith("kin")})
                    for s in range(self.sets)]
        idx = {id(self.inputs[s]): s for s in range(self.sets)}

        def f_kin_fused(x, o):
            b.step_fused_args(kin_args[idx[id(x)]])

        rew = [torch.empty((self.E,), dtype=torch.float64, device=self.inputs[0]["pose"].device) for _ in range(self.sets)]

        def f_reward(x, o):  # SURVEY.md 8f rank-1 "next" row, measured beside the step's kernels
            from terrainrlsim_b200 import _capi
            p = b._prep([(x["pose"], np.float64), (x["vel"], np.float64), (o["kin_pose"], np.float64),
                         (o["kin_vel"], np.float64), (x["body_vel"], np.float64), (None, np.uint8),
                         (rew[idx[id(x)]], np.float64), (None, np.float64)])
            _capi.lib().trl_imitate_calc_reward(b._h, *p)

        n_dof = self.model.num_dof
        jac = [torch.empty((self.E, 6, n_dof), dtype=torch.float64, device=self.inputs[0]["pose"].device) for _ in range(2)]
        gfo = torch.empty((self.E, n_dof), dtype=torch.float64, device=self.inputs[0]["pose"].device)

        def f_extras(x, o):  # world Jacobian, CoM Jacobian, gravity force
            from terrainrlsim_b200 import _capi
            p = b._prep([(x["pose"], np.float64), (jac[0], np.float64), (jac[1], np.float64), (gfo, np.float64)])
            _capi.lib().trl_rbd_extras(b._h, *p)

        for name, fn in (("imp_pd", f_pd), ("kin_char", f_kin), ("fk", f_fk), ("kin_fused", f_kin_fused),
                         ("poli_state", f_obs), ("imitate_reward", f_reward), ("rbd_extras", f_extras)):
            for s in range(self.sets):
                fn(self.inputs[s], self.outputs[s])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                b.use_torch_stream()
                for s in range(self.sets):
                    fn(self.inputs[s], self.outputs[s])
            b.use_torch_stream()
            g.replay()
            ms = timed_loop(lambda k: [g.replay() for _ in range(k)], reps)
            res[name] = ms / (reps * self.sets)
        return res

    def pd_only(self):
        s = self.i % self.sets
        self.i += 1
        x, o = self.inputs[s], self.outputs[s]
        self.batch.imp_pd_update_control_force(self.dt, x["pose"], x["vel"], x["tar_pose"], x["tar_vel"],
                                               joint_active=x["joint_active"], tau=o["tau"])


def timed_loop(fn, steps, barrier=None):
    """fn(steps) bracketed by (barrier +) synchronize on both sides, CUDA events on the current stream (the stream
    the library launches on). Returns ms total."""
    import torch
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(steps)
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1)


def bind_to_gpu_numa_node(local_rank):
    """Restrict this rank's threads to the CPUs of the NUMA node its GPU hangs off, so pinned staging buffers (first touch)
    land in that node's memory: with 8 ranks the host-coupled path is bound by the host's PCIe / DRAM fabric."""
    info = {"node": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        info["all_cpus"] = sorted(os.sched_getaffinity(0))
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(info["all_cpus"])
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update({"node": node, "cpus": len(cpus)})
    except Exception as exc:  # not fatal: the bench runs unbound
        info["error"] = str(exc)[:120]
    return info


def measure_e2e(w, name, E, world, rank, local_rank, args, barrier, max_over_ranks):
    """env-steps/s through the C ABI in HOST pointer mode: pinned host buffers, every step's H2D and D2H inside the timed
    region, synchronous calls (the call returns when the results are in host memory).

    What crosses PCIe is what SURVEY.md section 8d counts as the step's algorithmic bytes: pose, vel, tar_pose, the kin
    clock and the body states up (tar_vel is the reference's default zero: NULL); tau and the [E][F] double state down.
    The kinematic reference pose / vel are computed every step as well and consumed on the device by the imitation-reward
    epilogue (trl_step_fused_ex), whose [E] reward comes down too -- extra work the CPU arm does not do.
    Variants reported beside it: the previous round's contract (every kin output downloaded), the state as fp32, the
    state every 20th substep only (the reference's own policy-query rate, sim/CtController.cpp:300-325), and the headline
    contract double-buffered over two batches."""
    import torch
    from terrainrlsim_b200 import _capi
    n, N, F = w.model.num_dof, w.model.num_joints, w.batch.F
    pin = []

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pin.append(t)
        return t.numpy()

    def empty(shape, dtype=torch.float64):
        t = torch.empty(shape, dtype=dtype).pin_memory()
        pin.append(t)
        return t.numpy()

    def host_io(wl):
        hin = {k: pinned(v) for k, v in wl.host_inputs.items() if k != "tar_vel"}
        full = dict(hin)
        full["tar_vel"] = pinned(wl.host_inputs["tar_vel"])
        return hin, full

    hin, hin_full = host_io(w)
    tau, state, reward = empty((E, n)), empty((E, F)), empty((E,))
    state32 = empty((E, F), torch.float32)
    kin = {"kin_pose": empty((E, n)), "kin_vel": empty((E, n)), "kin_body_pos": empty((E, N, 3))}
    b = w.batch
    b.set_stream(None)
    nb = lambda d: int(sum(v.nbytes for v in d.values()))  # noqa: E731

    # (args, opts, host->device bytes, device->host bytes) per contract; pointers resolved once, like a C++ caller
    a_head = b.make_step_args(w.dt, hin, {"tau": tau, "state": state})
    o_head = b.make_step_opts(reward=reward)
    a_r01 = b.make_step_args(w.dt, hin_full, dict(kin, tau=tau, state=state))
    a_f32 = b.make_step_args(w.dt, hin, {"tau": tau})
    o_f32 = b.make_step_opts(reward=reward, state_f32=state32)
    a_sub = b.make_step_args(w.dt, hin, {"tau": tau})
    o_sub = b.make_step_opts(reward=reward)
    h2d = nb(hin)
    steps = max(5, min(args.steps, 50))

    def timed(fn, k):
        for _ in range(3):
            fn(0)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(k):
            rc = fn(i)
            if rc != 0:
                raise RuntimeError("trl_step_fused_ex (host mode) failed: " + _capi.last_error())
        t1 = time.perf_counter()
        barrier()
        return max_over_ranks((t1 - t0) * 1e3)

    def rate(ms, k):
        return world * E * k / (ms * 1e-3)

    ms = timed(lambda i: b.step_fused_args(a_head, o_head), steps)
    d2h = tau.nbytes + state.nbytes + reward.nbytes
    e2e = {"value": rate(ms, steps), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h), "steps": steps,
           "ms_per_step": ms / steps, "pcie_gbs_per_rank": (h2d + d2h) * steps / (ms * 1e-3) / 1e9,
           "timer": "host perf_counter around synchronous calls",
           "contract": "up: pose, vel, tar_pose, joint_active, kin clock / origin, body pos / vel, phase, flip; down: tau, "
                       "[E][F] f64 state, [E] imitation reward (kin pose / vel / FK computed every step, consumed on the device)"}
    variants = {}
    ms = timed(lambda i: b.step_fused_args(a_r01), steps)
    variants["all_outputs_r01"] = {"value": rate(ms, steps), "h2d_bytes_per_step": nb(hin_full),
                                   "d2h_bytes_per_step": int(tau.nbytes + state.nbytes + nb(kin)),
                                   "note": "round-1 contract: tar_vel uploaded, kin pose / vel / body positions downloaded"}
    ms = timed(lambda i: b.step_fused_args(a_f32, o_f32), steps)
    variants["state_f32"] = {"value": rate(ms, steps), "h2d_bytes_per_step": h2d,
                             "d2h_bytes_per_step": int(tau.nbytes + state32.nbytes + reward.nbytes),
                             "note": "state downloaded as float (what the reference feeds Caffe); f64 on the device"}
    k20 = 20 * max(1, steps // 20)
    ms = timed(lambda i: b.step_fused_args(a_head, o_head) if i % 20 == 0 else b.step_fused_args(a_sub, o_sub), k20)
    variants["state_every_20th_substep"] = {
        "value": rate(ms, k20), "steps": k20, "h2d_bytes_per_step": h2d,
        "d2h_bytes_per_step": int(tau.nbytes + reward.nbytes + state.nbytes / 20),
        "note": "observation computed and downloaded on 1 substep in 20 (cCtController::UpdateCalcTau's policy-query rate); "
                "torque, kin pose and reward every substep"}
    # the headline contract, double-buffered: two batches (each E envs, own pinned buffers) driven alternately through the
    # asynchronous HOST-pointer mode, so one batch's download overlaps the other's upload and kernels
    try:
        w2 = Workload(name, E, local_rank, rank + 1000, sets=1, fields=min(w.num_fields, 256))
        hin2, _ = host_io(w2)
        tau2, state2, reward2 = empty((E, n)), empty((E, F)), empty((E,))
        w2.batch.set_stream(None)
        a2 = w2.batch.make_step_args(w2.dt, hin2, {"tau": tau2, "state": state2})
        o2 = w2.batch.make_step_opts(reward=reward2)
        pair = [(b, a_head, o_head), (w2.batch, a2, o2)]
        for bt, aa, oo in pair:
            bt.step_fused_args(aa, oo)  # warm-up (synchronous)
            bt.set_host_async(True)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for bt, aa, oo in pair:
            bt.step_fused_args(aa, oo)
        for _ in range(steps // 2):
            for bt, aa, oo in pair:
                bt.sync()
                bt.step_fused_args(aa, oo)
        for bt, aa, oo in pair:
            bt.sync()
        t1 = time.perf_counter()
        barrier()
        nsteps = 2 + 2 * (steps // 2)
        pms = max_over_ranks((t1 - t0) * 1e3)
        e2e["pipelined"] = {"value": rate(pms, nsteps), "unit": UNIT, "steps": nsteps,
                            "mode": "two batches of E envs in flight (trl_batch_set_host_async), every step's H2D and D2H "
                                    "inside the timed region"}
        for bt, aa, oo in pair:
            bt.set_host_async(False)
        del w2
    except Exception as exc:  # the synchronous number above stands on its own
        e2e["pipelined"] = {"error": str(exc)[:200]}
    e2e["variants"] = variants
    return e2e


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; terrainrlsim_b200 has no CPU fallback (use --impl reference)")
    from terrainrlsim_b200 import _capi, synth
    import ctypes as C

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)  # pinned host buffers are allocated on the GPU's own NUMA node
    use_dist = world > 1
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if use_dist:
            dist.barrier()

    def max_over_ranks(v):
        if not use_dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    name = args.config
    E = args.envs or synth.load_model_dict(name)["envs_per_gpu"]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    side = torch.cuda.Stream(device=local_rank)  # a real (non-legacy) stream: events and kernels share it
    torch.cuda.set_stream(side)
    w = Workload(name, E, local_rank, rank, fields=default_fields(name, E, args.terrains))
    for _ in range(3):
        w.step()
    if not args.no_graph:
        w.capture()

    # warm-up: at least W steps and at least ~0.3 s so the clock sampler sees the GPU under load
    t_end = time.perf_counter() + 0.3
    k = 0
    while k < max(args.warmup, 3) or time.perf_counter() < t_end:
        w.run_steps(w.sets)
        k += w.sets
        torch.cuda.synchronize()

    l0 = w.batch.launch_count()
    ms = timed_loop(w.run_steps, args.steps, barrier)
    if w.graph is not None:
        launches = (args.steps // w.sets) * w.sets * w.launches_per_step + (w.batch.launch_count() - l0)
    else:
        launches = w.batch.launch_count() - l0
    ms = max_over_ranks(ms)
    value = world * E * args.steps / (ms * 1e-3)

    # dominant kernel (stable PD) timed alone on the same rotating sets
    pd_steps = max(4, min(args.steps, 400) // w.sets) * w.sets
    if w.pd_graph is not None:
        w.run_pd(w.sets)
        ms_pd = timed_loop(w.run_pd, pd_steps) / pd_steps
    else:
        def _pd(k):
            for _ in range(k):
                w.pd_only()
        _pd(5)
        ms_pd = timed_loop(_pd, pd_steps) / pd_steps

    # end-to-end through the public HOST-pointer API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(w, name, E, world, rank, local_rank, args, barrier, max_over_ranks)
        w.batch.use_torch_stream()

    kernel_ms = w.breakdown() if not args.no_graph else None

    # FP64 peak probe (register-resident DFMA chains) for the roofline denominator
    tf = C.c_double(0)
    fp64_peak = None
    if _capi.lib().trl_device_fp64_probe(local_rank, C.byref(tf)) == 0:
        fp64_peak = tf.value
    clocks = sampler.stop() if sampler else None

    # optional episode-statistics all-gather over NCCL (outside the timed region, 64-byte record per GPU)
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    stats[0] = float(args.steps * E)
    stats[1] = float(w.outputs[0]["tau"].abs().max().item())
    stats[2] = float((~torch.isfinite(w.outputs[0]["state"])).sum().item())
    gathered = None
    if use_dist:
        lst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(lst, stats)
        gathered = [x.cpu().tolist() for x in lst]

    # the other BASELINE.json configs at their stated envs per GPU, on every rank (weak scaling, max over ranks)
    others = {}
    if not args.no_others:
        del w.inputs[1:], w.outputs[1:], w.args[1:]
        for oname in ROOFLINE_MODEL:
            if oname == name:
                continue
            oE = synth.load_model_dict(oname)["envs_per_gpu"]
            ow = Workload(oname, oE, local_rank, rank)
            for _ in range(3):
                ow.step()
            if not args.no_graph:
                ow.capture()
            ok = 20 * ow.sets
            ow.run_steps(ok)
            oms = max_over_ranks(timed_loop(ow.run_steps, ok, barrier))
            opd = max_over_ranks(timed_loop(ow.run_pd, ok) / ok) if ow.pd_graph is not None else None
            pdf, totf, byt = ROOFLINE_MODEL[oname]
            others[oname] = {"envs_per_gpu": oE, "n_gpus": world, "value": world * oE * ok / (oms * 1e-3), "unit": UNIT,
                             "ms_per_step": oms / ok, "ms_per_pd_launch": opd,
                             "fp64_tflops_per_gpu": totf * oE * ok / (oms * 1e-3) / 1e12,
                             "hbm_gbs_per_gpu": byt * oE * ok / (oms * 1e-3) / 1e9}
            del ow

    if rank == 0:
        pd_flops, total_flops, bytes_step = ROOFLINE_MODEL[name]
        hbm_peak, hbm_src = measured_peaks()
        ach_tf = pd_flops * E / (ms_pd * 1e-3) / 1e12
        nominal = 37.2
        peak_tf = fp64_peak if fp64_peak else nominal
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_pd_traffic.json")))
            if tj.get("config") == name and tj.get("envs") == E:
                traffic = tj["pd_entry_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        roof = {"bound": "fp64", "kernel": "stable-PD entry: k_pd_aba", "achieved": ach_tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
                "traffic_note": "DRAM bytes of one k_pd_aba launch inside the fused step under ncu --set full (profiles/r02_pd_traffic.json)",
                "peak_source": "DFMA probe measured in this run" if fp64_peak else "nominal 148 SM x 64 x 2 x 1.965 GHz",
                "ms_per_launch": ms_pd, "algorithmic_flops_per_launch": pd_flops * E}
        ms_step = ms / args.steps
        ach_gbs = bytes_step * E / (ms_step * 1e-3) / 1e9
        roof_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "peak_source": hbm_src, "whole_step_fp64_tflops": total_flops * E / (ms_step * 1e-3) / 1e12}
        # the step's longest kernel chain since round 2 is the observation (k_obs_rec + k_obs_feat + k_obs_gather): HBM-bound
        # gathers. Algorithmic bytes per env: pose / vel and body rows in, the vertex floats of every ground sample in, the
        # [F] double state out.
        roof_obs = None
        if kernel_ms and kernel_ms.get("poli_state"):
            G = w.batch.cfg.num_samples
            verts = 3 if w.d.get("terrain", "").startswith("var3d") else 2  # vertex floats a sample interpolates
            obs_bytes = 2 * w.model.num_dof * 8 + 2 * w.model.num_joints * 24 + 9 + G * verts * 4 + w.batch.F * 8
            ach = obs_bytes * E / (kernel_ms["poli_state"] * 1e-3) / 1e9
            obs_traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r02_obs_traffic.json")))
                if tj.get("config") == name and tj.get("envs") == E:
                    obs_traffic = tj["obs_entry_bytes_per_launch"]
            except (OSError, ValueError, KeyError):
                pass
            roof_obs = {"bound": "hbm", "kernel": "observation entry: k_obs_rec + k_obs_feat + k_obs_gather / k_obs_samples",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": obs_traffic,
                        "ms_per_launch": kernel_ms["poli_state"], "algorithmic_bytes_per_launch": obs_bytes * E,
                        "peak_source": hbm_src}
        cpu = None
        if not args.no_cpu_baseline:
            if numa.get("all_cpus"):
                os.sched_setaffinity(0, numa["all_cpus"])  # the CPU baseline uses every host core
            cpu = cpu_baseline(name, E, w.num_fields)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAMES[name], "envs_per_gpu": E, "joints": w.model.num_joints,
                           "dofs": w.model.num_dof, "ground_samples": w.batch.cfg.num_samples, "features": w.batch.F,
                           "dt": w.dt, "terrains_per_gpu": w.num_fields,
                           "terrain": "%s generated on the device (trl_ground_generate)" % w.d.get("terrain_file", w.d["terrain"]),
                           "l2": "rotating %d device-resident input/output sets, %.0f MB total > L2 (126 MB)" %
                                 (w.sets, w.sets * (w.in_bytes + w.out_bytes) / 1e6),
                           "sharding": "envs partitioned by rank, no hot-path collective"},
                "e2e": e2e, "gpu_launches": int(launches), "launch": "direct" if args.no_graph else
                "CUDA graph of %d fused steps (%d kernels each) replayed" % (w.sets, w.launches_per_step),
                "clocks": clocks, "roofline": roof,
                "kernel_ms": kernel_ms, "roofline_observation": roof_obs, "roofline_step_hbm": roof_hbm, "roofline_model": roofline_model(name), "cpu_baseline": cpu,
                "host_cores": os.cpu_count(), "numa": {k: v for k, v in numa.items() if k != "all_cpus"},
                "other_workloads": others,
                "episode_stats_allgather": gathered}
        print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="biped3d", choices=sorted(ROOFLINE_MODEL))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the config's BASELINE.json size)")
    ap.add_argument("--terrains", type=int, default=0, help="distinct terrains per GPU (default: min(envs, 4096))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
