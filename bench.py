#!/usr/bin/env python
"""bench.py -- batched world-steps/sec of the stepTheWorld hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU path on the host cores

Workload (config D of BASELINE.json / SURVEY.md section 8d): 65536 worlds PER GPU of a wheeled rover --
chassis box + 4 wheels on hinge2 joints (steer axis locked by stops, drive axis motorised) --
driving on one heightfield (256x256 @ 0.1 m) shared by all worlds; dt 1 ms, 20 SOR iterations,
ODE row order with LCG shuffles.  One "step" = device collision (broadphase + narrowphase, which
regenerates the contact stream) + the fused assemble / SOR-LCP / integrate kernel over all worlds.
Worlds shard over GPUs with no collective on the step path (weak scaling: per-GPU work fixed).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the
C ABI with host buffers (pinned H2D of the per-world wheel commands and D2H of the chassis state
inside the timed region).  The per-step working set (row scratch ~0.4 GB per 65536 worlds) is
larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.
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
sys.path.insert(0, ROOT)

H = 0.001
WORLDS_PER_GPU = 65536
SETTLE_STEPS = 600  # untimed: the rovers are dropped from 0.12 m and must be driving on the terrain
METRIC = "batched world-steps/sec"
UNIT = "world-steps/s"
WORKLOAD = "D: hinge2 rover (5 bodies, 4 hinge2, sphere wheels) on shared 256x256 heightfield, dt=1ms, 20 SOR iters"


def build_rover_worlds(be, num_worlds, world_offset=0, torch_free=True, wheel_shape="sphere"):
    """Config D scene on any backend; per-world placement / wheel speed from splitmix64(0xD ^ id)."""
    from mars_ode_physics_b200 import capi, scenes
    sc = scenes.rover(be, with_geoms=True, wheel_shape=wheel_shape,
                      max_contacts=8 if wheel_shape == "sphere" else 20)
    heights, x0, y0 = scenes.terrain_heights()
    be.geom_create_heightfield(heights, 0.1, x0, y0,
                               surface={"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5})
    ids = np.arange(world_offset, world_offset + num_worlds, dtype=np.uint64)
    ux = np.array([scenes.uniform01(0xD, int(i), 1) for i in ids]) if num_worlds <= 4096 else \
        scenes.uniform01_array(0xD, world_offset + num_worlds, 1)[world_offset:]
    uy = np.array([scenes.uniform01(0xD, int(i), 2) for i in ids]) if num_worlds <= 4096 else \
        scenes.uniform01_array(0xD, world_offset + num_worlds, 2)[world_offset:]
    uv = np.array([scenes.uniform01(0xD, int(i), 3) for i in ids]) if num_worlds <= 4096 else \
        scenes.uniform01_array(0xD, world_offset + num_worlds, 3)[world_offset:]
    dx, dy = 16.0 * ux - 8.0, 16.0 * uy - 8.0
    vel = (2.0 + 4.0 * uv).astype(np.float32)
    nb = len(sc["bodies"])
    st = be.state_download(nb)
    st = np.array(st, dtype=st.dtype)
    st[:, 0, :] += dx.astype(st.dtype)
    st[:, 1, :] += dy.astype(st.dtype)
    st[:, 2, :] += 0.12
    if hasattr(be, "state_upload"):
        be.state_upload(st)
    else:
        for w in range(num_worlds):
            for b in range(nb):
                be.body_set_state(b, st[b, :, w], w)
    for j in sc["joints"]:
        be.joint_set_param_batch(j, capi.PARAM_VEL | capi.PARAM_GROUP2, vel)
    sc["vel"] = vel
    return sc


def build_box_stack_worlds(be, num_worlds, world_offset=0):
    """Config B: 10-box stack on a plane, per-world jitter from splitmix64(0xB200 ^ id)."""
    from mars_ode_physics_b200 import scenes
    sc = scenes.box_stack(be, n=10, max_contacts=48)
    nb = len(sc["bodies"])
    st = np.array(be.state_download(nb))
    n = world_offset + num_worlds
    for i in range(nb):
        x = -0.01 + 0.02 * scenes.uniform01_array(0xB200, n, 3 * i + 0)[world_offset:]
        y = -0.01 + 0.02 * scenes.uniform01_array(0xB200, n, 3 * i + 1)[world_offset:]
        yaw = -0.1 + 0.2 * scenes.uniform01_array(0xB200, n, 3 * i + 2)[world_offset:]
        st[i, 0, :] = x
        st[i, 1, :] = y
        st[i, 3, :] = np.cos(0.5 * yaw)
        st[i, 6, :] = np.sin(0.5 * yaw)
    be.state_upload(st) if hasattr(be, "state_upload") else [
        be.body_set_state(b, st[b, :, w], w) for w in range(num_worlds) for b in range(nb)]
    sc["vel"] = None
    sc["chassis"] = sc["bodies"][-1]
    return sc


def build_quadruped_worlds(be, num_worlds, world_offset=0):
    """Config C: 12-hinge legged robot, foot spheres on a plane, constant per-world joint speeds."""
    from mars_ode_physics_b200 import capi, scenes
    sc = scenes.quadruped(be, max_contacts=8)
    s = {"mu": 0.8, "mu2": 0.8, "soft_erp": 0.2, "soft_cfm": 1e-5}
    be.geom_create(capi.GEOM_PLANE, -1, (0, 0, 1, 0), surface=s)
    for f in sc["feet"]:
        be.geom_create(capi.GEOM_SPHERE, f, (sc["foot_radius"],), surface=s)
    be.geom_create(capi.GEOM_BOX, sc["torso"], (0.4, 0.2, 0.05), surface=s)
    n = world_offset + num_worlds
    vel = (-0.5 + scenes.uniform01_array(0xC, n, 1)[world_offset:]).astype(np.float32)
    for j in sc["hip"] + sc["knee"]:
        be.joint_set_param_batch(j, capi.PARAM_VEL, vel)
    sc["vel"] = vel
    sc["chassis"] = sc["torso"]
    return sc


def build_rover_cylinder_worlds(be, num_worlds, world_offset=0):
    return build_rover_worlds(be, num_worlds, world_offset, wheel_shape="cylinder")


BUILDERS = {"D": None, "B": build_box_stack_worlds, "C": build_quadruped_worlds, "Dcyl": build_rover_cylinder_worlds}
WORKLOADS = {
    "Dcyl": "D with cylinder wheels (r 0.1, width 0.05; up to 4 contacts per wheel) on the shared heightfield",
    "B": "B: 10-box stack on a plane (box-box + box-plane contacts), dt=1ms, 20 SOR iters",
    "C": "C: 12-hinge legged robot (13 bodies), foot spheres on a plane, dt=1ms, 20 SOR iters",
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.  ONE long-running `nvidia-smi -lms 100`
    process started well before the timed region (no fork, no Python thread and no GIL traffic while the launch
    loop is being timed); its time-stamped lines are parsed afterwards and cut to the timed windows."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpus = ",".join(str(g) for g in gpu_indices)  # one nvidia-smi process covers all GPUs of the job
        self.proc, self.out = None, None
        self.window = None  # (t0, t1) time.time() bounds; samples outside are dropped when any lie inside
        self.samples = []

    def start(self):
        import tempfile
        try:
            self.out = tempfile.TemporaryFile(mode="w+")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpus, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            import datetime
            self.out.seek(0)
            for line in self.out.read().strip().splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    float(parts[1])
                except Exception:
                    continue
                self.samples.append(parts[1:] + [ts])
        except Exception:
            pass
        self.proc = None

    def summary(self):
        samples = self.samples
        if self.window is not None:
            inside = [x for x in samples if self.window[0] - 0.05 <= x[-1] <= self.window[1] + 0.05]
            samples = inside or samples
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(x[0]) for x in samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(x[3 + i].lower().startswith("active") for x in samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": reasons,
                "samples": len(samples), "power_w_max": max(float(x[2]) for x in samples)}


def pin_to_gpu_numa_node(local_rank, world_size):
    """Keep this rank's launch thread on the cores next to its GPU (NUMA node of the GPU's PCI device, shared
    evenly between the ranks on that node); falls back to an even split of the allowed cores."""
    try:
        import torch
        allowed = sorted(os.sched_getaffinity(0))
        cores = None
        try:
            bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
            dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
            dev = torch.cuda.get_device_properties(local_rank).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
            node = int(open(path).read().strip())
            if node >= 0:
                txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
                node_cores = []
                for part in txt.split(","):
                    lo, _, hi = part.partition("-")
                    node_cores += list(range(int(lo), int(hi or lo) + 1))
                node_cores = [c for c in node_cores if c in allowed]
                # ranks sharing the node: assume GPUs are spread evenly over the nodes in index order
                nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
                per_node = max(1, (world_size + nodes - 1) // nodes)
                k = local_rank % per_node
                chunk = max(1, len(node_cores) // per_node)
                cores = node_cores[k * chunk:(k + 1) * chunk] or node_cores
        except Exception:
            cores = None
        if not cores:
            chunk = max(1, len(allowed) // max(1, world_size))
            cores = allowed[local_rank * chunk:(local_rank + 1) * chunk] or allowed
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


class CpuBaseline:
    """The oracle (CPU restatement of ODE quickstep + the collision spec) on the host cores:
    `sample_worlds` worlds of the same workload, settled, then timed over bounded samples."""

    def __init__(self, sample_worlds, threads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc  # test infrastructure; used here only as the reported CPU baseline
        from concurrent.futures import ThreadPoolExecutor
        self.n, self.threads = sample_worlds, threads
        self.oc = orc.OracleBatch(sample_worlds)
        build_rover_worlds(self.oc, sample_worlds)
        self.oc.set_order_mode(orc.ORDER_ODE_ISLANDS)  # ODE's island order, as the device arm runs
        per = (sample_worlds + threads - 1) // threads
        self.chunks = [(i, min(i + per, sample_worlds)) for i in range(0, sample_worlds, per)]
        self.ex = ThreadPoolExecutor(max_workers=threads)
        self.run(SETTLE_STEPS)

    def run(self, steps):
        t0 = time.perf_counter()
        list(self.ex.map(lambda c: self.oc.bench_run(c[0], c[1], steps, H, 1), self.chunks))
        dt = time.perf_counter() - t0
        return self.n * steps / dt, dt

    def run_for(self, seconds):
        """Time about `seconds` of CPU work; returns (world-steps/s, steps, seconds)."""
        rate, _ = self.run(20)
        steps = max(20, int(seconds * rate / self.n))
        rate, dt = self.run(steps)
        return rate, steps, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path (libode is not available, so this is the
    oracle port of ODE-0.16 quickstep) on all host threads, same config / metric, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 64 * cores
    cb = CpuBaseline(sample, cores)
    rate, _ = cb.run(20)
    # each bench "step" is a bounded sample: `steps_per` steps of `sample` worlds; the whole run
    # (warmup + steps) is sized to stay within ~2 minutes
    budget_s = min(120.0, 1.0 * (args.steps + args.warmup))
    steps_per = max(5, int(budget_s * rate / sample / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cb.run(steps_per)
    vals, t_total = [], 0.0
    for _ in range(args.steps):
        v, dt = cb.run(steps_per)
        vals.append(v)
        t_total += dt
    value = float(sample * steps_per * args.steps / t_total)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "worlds_per_sample": sample, "world_steps_per_step": sample * steps_per,
                   "settle_steps": SETTLE_STEPS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} settled worlds x {steps_per} steps per bench step, oracle f64 "
                                   f"(CPU restatement of ODE-0.16 quickstep, not libode), {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def measure_workload(Arena, torch, stream, name, builder, W, world_offset, settle, steps, peak, order_mode=1, dense=False):
    """A secondary configuration (BASELINE configs 2, 3, 5): device-resident collide + step over `steps` steps,
    CUDA events on the launch stream, plus the per-kernel split.  Returns (dict for the JSON line, ms).
    dense: the worlds settle under dWorldQuickStep, the timed steps run dWorldStep (fast_step == false)."""
    ar = Arena(W, device=torch.cuda.current_device())
    ar.set_stream(stream.cuda_stream)
    builder(ar, W, world_offset=world_offset)
    ar.set_order_mode(order_mode)
    for _ in range(settle):
        ar.collide_step(H)
    if dense:
        ar.set_stepper(1)
        for _ in range(3):
            ar.collide_step(H)
    ar.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ar.collide_step(H)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(min(steps, 20))]
    for e in ev:
        e[0].record(stream)
        ar.collide()
        e[1].record(stream)
        ar.step(H)
        e[2].record(stream)
    torch.cuda.synchronize()
    collide_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / len(ev)
    step_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / len(ev)
    ar.debug_kernel_times(True)
    for _ in range(min(steps, 20)):
        ar.collide()
        ar.step(H)
    kt = ar.debug_kernel_times(False)
    balg = ar.algorithmic_bytes_per_world_step()
    rows = float(np.mean([ar.last_num_rows(w) for w in range(0, W, max(1, W // 128))]))
    achieved = balg * W / (step_ms * 1e-3) / 1e9
    out = {"workload": name, "worlds_per_gpu": W, "settle_steps": settle, "steps": steps, "ms_per_step": ms / steps,
           "mean_rows_per_world": rows, "contacts_dropped": ar.contacts_dropped(),
           "kernel_ms": {"collide": collide_ms, "step": step_ms, "assemble": float(kt[0]), "sor": float(kt[1]),
                         "integrate": float(kt[2])},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_world_step": balg}}
    if dense:
        out["stepper"] = "dWorldStep (fast_step == false): dense LCP, Dantzig's method, FP64, one warp per world in shared memory; kernel_ms.sor is b2p_dense_warp_kernel"
        out["lcp_failures"] = ar.dense_failures()
        out["max_lcp_residual"] = float(ar.dense_residual().max())
    ar.close()
    return out, ms


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mars_ode_physics_b200 import Arena, capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    pinned_cores = pin_to_gpu_numa_node(local_rank, world_size) if world_size > 1 else 0
    # stdout carries exactly one line (the JSON result of rank 0): anything a library prints there meanwhile
    # (e.g. NCCL's version banner) goes to stderr instead
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = args.worlds
    ar = Arena(W, device=local_rank)
    stream = torch.cuda.current_stream()
    ar.set_stream(stream.cuda_stream)
    if args.workload == "D":
        sc = build_rover_worlds(ar, W, world_offset=rank * W)
    else:
        sc = BUILDERS[args.workload](ar, W, world_offset=rank * W)
    if args.resident_rows:
        ar.set_resident_rows(args.resident_rows)
    if args.register_rows >= 0:
        ar.set_register_rows(args.register_rows)
    ar.set_sor_kernel({"auto": -1, "paired": 0, "tmem": 1}[args.sor_kernel])
    ar.set_sor_warps(args.sor_warps)
    if args.no_graph:
        ar.set_graph_launch(False)
    ar.set_order_mode(capi.ORDER_ODE_ISLANDS if args.row_order == "islands" else capi.ORDER_WORLD_BLOCK)
    if args.reorder == "none":
        ar.set_reorder_method(capi.REORDER_NONE)
    nb, nj = len(sc["bodies"]), len(sc["joints"])
    cmd_joints = sc["joints"] if sc["vel"] is not None else []

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident loop ----------------
    # rank 0 samples the clocks of every GPU of the job (the other ranks spawn nothing: eight nvidia-smi loops
    # next to eight launch threads cost host time the launches need).  The sampler runs from before the settle
    # phase, so that no process is forked at the start of the timed window; only samples inside the timed windows
    # are reported.
    sampler = ClockSampler(range(world_size) if world_size > 1 else [local_rank])
    if rank == 0:
        sampler.start()
    for _ in range(args.settle):
        ar.collide_step(H)
    ar.sync()
    for _ in range(max(args.warmup, 3)):
        ar.collide_step(H)
    K = args.steps
    launches0 = ar.kernel_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    t_epoch0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for k in range(K):  # the timed region: nothing but the product call (one CUDA graph launch: 4 kernels)
        ar.collide_step(H)
    end.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms = start.elapsed_time(end)
    launches = ar.kernel_launch_count() - launches0
    # collide / step split: a separate pass with events around the two calls
    KE = min(K, 50)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(KE)]
    for k in range(KE):
        ev[k][0].record(stream)
        ar.collide()
        ev[k][1].record(stream)
        ar.step(H)
        ev[k][2].record(stream)
    torch.cuda.synchronize()
    collide_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / KE
    step_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / KE
    # per-kernel durations of the step (CUDA events on the arena's stream between its three launches;
    # a separate pass because every step then synchronises)
    ar.debug_kernel_times(True)
    for _ in range(K):
        ar.collide()
        ar.step(H)
    kt = ar.debug_kernel_times(False)
    asm_ms, sor_ms, int_ms = float(kt[0]), float(kt[1]), float(kt[2])

    # ---------------- end-to-end loop through the C ABI with host buffers ----------------
    # Every step: this step's wheel commands host -> device from pinned memory, collide + step, chassis state
    # device -> host.  The loop is a pipelined control loop of depth 2: the observation of step k is read while
    # step k+1 runs (its D2H goes over a copy stream), i.e. commands are computed from the observation of the
    # step before last.  Command buffers are prepared outside the timed region (4 rotating pinned buffers).
    NB = 4
    cmds = []
    for i in range(NB):
        c = torch.empty((max(nj, 1), W), dtype=torch.float32).pin_memory()
        c[:] = (torch.from_numpy(sc["vel"])[None, :] + 0.001 * i) if sc["vel"] is not None else 0.0
        cmds.append(c)
    obs = [torch.empty((13, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    jobs = [torch.empty((max(nj, 1), 4, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    param = (capi.PARAM_VEL | capi.PARAM_GROUP2) if args.workload in ("D", "Dcyl") else capi.PARAM_VEL

    def e2e_step(k, pending):
        cmd = cmds[k % NB]
        if cmd_joints and cmd_joints == list(range(len(cmd_joints))):
            ar.joints_set_param_batch_ptr(param, cmd.data_ptr(), len(cmd_joints))  # one transfer
        else:
            for ji, j in enumerate(cmd_joints):
                ar.joint_set_param_batch_ptr(j, param, cmd[ji].data_ptr())
        ar.collide_step(H)
        # observation: chassis state + angle / rate of every joint (what a controller reads each tick)
        t = ar.observe_begin(sc["chassis"], obs[k & 1].data_ptr(), jobs[k & 1].data_ptr() if nj else None)
        if pending is not None:
            ar.io_wait(pending)  # the observation of the previous step is complete: the caller may read it
        return t

    pending = None
    for k in range(3):
        pending = e2e_step(k, pending)
    ar.io_wait(pending)
    ar.sync()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_e2e0 = time.perf_counter()
    pending = None
    for k in range(K):
        pending = e2e_step(k, pending)
    ar.io_wait(pending)
    e1.record(stream)
    barrier()
    t_e2e = time.perf_counter() - t_e2e0
    sampler.window = (t_epoch0, time.time())  # device-resident loop .. end of the end-to-end loop
    if rank == 0:
        sampler.stop()
    # the last D2H ends on the copy stream: take the larger of the stream time and the host wall clock
    ms_e2e = max(e0.elapsed_time(e1), t_e2e * 1e3)
    assert np.isfinite(obs[0].numpy()).all() and np.isfinite(obs[1].numpy()).all(), "non-finite chassis state"
    assert np.isfinite(jobs[0].numpy()).all() and np.isfinite(jobs[1].numpy()).all(), "non-finite joint observation"

    # ---------------- reduce over ranks (max time) ----------------
    t = torch.tensor([ms, ms_e2e, step_ms, collide_ms, asm_ms, sor_ms, int_ms], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, step_ms, collide_ms, asm_ms, sor_ms, int_ms = [float(x) for x in t.cpu()]
    total_worlds = W * world_size
    value = total_worlds * K / (ms * 1e-3)
    e2e_value = total_worlds * K / (ms_e2e * 1e-3)

    balg = ar.algorithmic_bytes_per_world_step()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---------------- BASELINE config 5 (strong scaling): 524288 rover worlds split over the GPUs of the job ---------
    extra = {}
    if args.workload == "D" and not args.no_extra:
        total_strong = 524288
        Ws = total_strong // world_size
        barrier()
        res, ms_s = measure_workload(Arena, torch, stream, WORKLOAD, build_rover_worlds, Ws, rank * Ws,
                                     args.settle, max(10, min(K, 20)), peak,
                                     capi.ORDER_ODE_ISLANDS if args.row_order == "islands" else capi.ORDER_WORLD_BLOCK)
        t2 = torch.tensor([ms_s], dtype=torch.float64, device="cuda")
        if world_size > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_s = float(t2.cpu()[0])
        res["total_worlds"] = total_strong
        res["ms_per_step"] = ms_s / res["steps"]
        res["value"] = total_strong * res["steps"] / (ms_s * 1e-3)
        res["unit"] = UNIT
        res["scaling"] = "strong (total worlds fixed at 524288; same per-world seeds for every GPU count)"
        extra["strong_scaling_524288"] = res
        # ---------------- BASELINE configs 2 and 3 at their sizes (one GPU each; rank 0 reports) ----------------
        if rank == 0:
            for key, name, builder, Wx, settle_x in (("config_B_4096", WORKLOADS["B"], build_box_stack_worlds, 4096, 150),
                                                     ("config_C_16384", WORKLOADS["C"], build_quadruped_worlds, 16384, 300)):
                r, ms_x = measure_workload(Arena, torch, stream, name, builder, Wx, 0, settle_x, 100, peak,
                                           capi.ORDER_ODE_ISLANDS if args.row_order == "islands" else capi.ORDER_WORLD_BLOCK)
                r["value"] = Wx * r["steps"] / (ms_x * 1e-3)
                r["unit"] = UNIT
                extra[key] = r
            # the exact stepper (dWorldStep, the reference's fast_step == false) on config D
            Wd = W
            r, ms_x = measure_workload(Arena, torch, stream, WORKLOAD, build_rover_worlds, Wd, 0, args.settle, 10, peak,
                                       capi.ORDER_ODE_ISLANDS if args.row_order == "islands" else capi.ORDER_WORLD_BLOCK,
                                       dense=True)
            r["value"] = Wd * r["steps"] / (ms_x * 1e-3)
            r["unit"] = UNIT
            extra["dense_stepper_D_%d" % Wd] = r
        barrier()
    achieved = balg * W / (step_ms * 1e-3) / 1e9
    mean_rows = float(np.mean([ar.last_num_rows(w) for w in range(0, W, max(1, W // 256))]))
    # compulsory traffic of the SOR kernel per world: rows in (64 B), lambda out (4 B), fch out, counters
    sor_bytes = mean_rows * 68.0 + nb * 24.0 + 12.0
    sor_name = {capi.SOR_TMEM: "b2p_sor_tmem_kernel (rows in registers + Tensor Memory, one thread per world)",
                capi.SOR_PAIRED: "b2p_sor_kernel (rows in registers + shared memory, two lanes per world)"}[ar.sor_kernel_used()]
    # measured DRAM traffic per launch (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum) of the same
    # command, committed under profiles/ by the round's profiling pass
    traffic, per_kernel_traffic = None, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.workload == "D" and W == WORLDS_PER_GPU:
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("step_kernel_dram_bytes_per_launch")
            per_kernel_traffic = {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in tj.get("kernels", {}).items()}
        except Exception:
            traffic = None
    clocks = sampler.summary() if rank == 0 else {}
    # FP32 side of the roofline (SURVEY 8d asks for it beside the HBM figure): 27 FMA per row-iteration against
    # 148 SMs x 128 FP32 lanes at the SM clock the run held
    sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    fma_peak = 148 * 128 * sm_mhz * 1e6
    row_it_s = mean_rows * 20 * W / (sor_ms * 1e-3) if sor_ms > 0 else 0.0
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.workload == "D" else WORKLOADS[args.workload],
                       "worlds_per_gpu": W, "total_worlds": total_worlds,
                       "sharding": f"worlds/{world_size} GPUs, no collective on the step path",
                       "l2": "per-step working set (row scratch) > L2; no flush needed",
                       "settle_steps": args.settle, "row_reordering": args.reorder, "row_order": args.row_order,
                       "launch": "one CUDA graph (collide + assemble + SOR + integrate) per step; in the end-to-end loop two, split "
                                 "behind the assemble kernel, so that the next command upload can wait for an event "
                                 "recorded between them",
                       "host_cores_pinned": pinned_cores,
                       "collision_semantics": "repo-defined (oracle/orc_collide.c), not ODE's dCollide: one contact "
                                              "per sphere-heightfield pair, <= 4 deepest vertices per box-plane pair",
                       "mean_rows_per_world": mean_rows,
                       "kernel_ms": {"collide": collide_ms, "step": step_ms, "assemble": asm_ms, "sor": sor_ms,
                                     "integrate": int_ms}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(len(cmd_joints) * W * 4),
                    "d2h_bytes_per_step": int((13 + 4 * nj) * W * 4), "ms_per_step": ms_e2e / K,
                    "observation": "chassis state (13 floats) + angle1 rate1 angle2 rate2 of every joint, per world",
                    "mode": "pipelined control loop, depth 2: on a copy stream the D2H of step k overlaps step k+1 and "
                            "the H2D of step k+1's commands follows step k's assemble kernel"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         # the same with the collide kernel's time in the denominator (the algorithmic bytes count
                         # the geoms and the device-produced contact stream)
                         "frac_incl_collide": balg * W / ((step_ms + collide_ms) * 1e-3) / 1e9 / peak,
                         "kernel": "one dWorldQuickStep = b2p_assemble_kernel + SOR-LCP kernel + b2p_integrate_kernel "
                                   "(three launches; achieved = algorithmic bytes of the step / their summed duration)",
                         "algorithmic_bytes_per_world_step": balg,
                         "sor_kernel": {"name": sor_name, "ms": sor_ms, "share_of_step": sor_ms / step_ms,
                                        "io_bytes_per_world": sor_bytes,
                                        "io_gbs": sor_bytes * W / (sor_ms * 1e-3) / 1e9 if sor_ms > 0 else None,
                                        "measured_dram_gbs": (per_kernel_traffic["sor"] / (sor_ms * 1e-3) / 1e9
                                                              if "sor" in per_kernel_traffic and sor_ms > 0 else None),
                                        "row_iterations_per_s": row_it_s or None,
                                        "fp32_fma_frac": 27.0 * row_it_s / fma_peak if row_it_s else None,
                                        "fp32_fma_peak": f"148 SMs x 128 lanes x {sm_mhz:.0f} MHz",
                                        "bound": "on-chip: dependent smem round trip + FP32 issue per row-iteration "
                                                 "(DESIGN.md section 4); its HBM traffic is one pass over the rows"},
                         "measured_dram_gbs": {k: (per_kernel_traffic[k] / (ms_k * 1e-3) / 1e9)
                                               for k, ms_k in (("assemble", asm_ms), ("sor", sor_ms),
                                                               ("integrate", int_ms), ("collide", collide_ms))
                                               if k in per_kernel_traffic and ms_k > 0},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
            "clocks": clocks,
            "extra_configs": extra,
            "wall_s": t_wall,
        }
        if world_size == 1 and args.workload == "D" and not args.no_facade:
            # the same workload built and stepped through the reference-facing plugin API (C++ facade:
            # WorldPhysicsLoader / createFrame / createObject / createJoint / stepTheWorld, batched I/O), its own
            # process and arena; commands and observations travel through pinned host buffers every step
            exe = os.path.join(ROOT, "mars_ode_physics_b200", "lib", "facade_bench")
            try:
                ar.close()
                out = subprocess.run([exe, str(W), str(max(K, 20)), "5", str(SETTLE_STEPS)], capture_output=True,
                                     text=True, timeout=600,
                                     env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank))))
                fl = json.loads(out.stdout.strip().splitlines()[-1])
                line["e2e_facade"] = {"value": fl["e2e_facade"], "unit": UNIT, "ms_per_step": fl["ms_per_step"],
                                      "h2d_bytes_per_step": fl["h2d_bytes_per_step"],
                                      "d2h_bytes_per_step": fl["d2h_bytes_per_step"], "ok": fl["ok"],
                                      "path": "facade_bench: plugin loader -> PhysicsInterface::stepTheWorld "
                                              "(device collision + step + post-step bookkeeping), wall clock"}
            except Exception as e:  # the number is additional evidence, never a reason to lose the bench line
                line["e2e_facade"] = {"value": None, "error": str(e)[:200]}
        if not args.no_cpu_baseline and world_size == 1 and args.workload == "D":
            cores = os.cpu_count() or 1
            sample = 128 * cores  # enough worlds that ~12 s of CPU work stays within a few thousand steps
            v, nsteps, dt = CpuBaseline(sample, cores).run_for(12.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} settled worlds x {nsteps} steps of the same workload, "
                                              f"oracle f64 (CPU restatement of ODE-0.16 quickstep, not libode), "
                                              f"{cores} threads, {dt:.1f} s"}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=WORLDS_PER_GPU, help="worlds per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-facade", action="store_true", help="skip the C++ plugin-API leg (e2e_facade)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the secondary configurations (strong scaling at 524288 worlds, configs B and C)")
    ap.add_argument("--no-graph", action="store_true",
                    help="profiling aid: launch the four kernels directly instead of replaying the CUDA graph")
    ap.add_argument("--settle", type=int, default=SETTLE_STEPS, help="profiling aid: untimed settle steps")
    ap.add_argument("--resident-rows", type=int, default=0,
                    help="tuning: constraint rows per world kept in shared memory by the SOR kernel (0 = auto)")
    ap.add_argument("--register-rows", type=int, default=-1,
                    help="tuning: sweep positions per world kept in registers by the SOR kernel (-1 = auto)")
    ap.add_argument("--sor-kernel", default="auto", choices=["auto", "paired", "tmem"],
                    help="SOR-LCP kernel: rows in registers + Tensor Memory (tmem) or registers + shared memory (paired)")
    ap.add_argument("--sor-warps", type=int, default=0, choices=[0, 4, 8],
                    help="Tensor-Memory SOR kernel: warps (x32 worlds) per CTA; 0 = auto")
    ap.add_argument("--reorder", default="random", choices=["random", "none"],
                    help="SOR row reordering: ODE 0.16 default (random) or REORDERING_METHOD__DONT_REORDER")
    ap.add_argument("--row-order", default="islands", choices=["islands", "block"],
                    help="constraint row order: ODE's island order (dxProcessIslands; default) or one block per world")
    ap.add_argument("--workload", default="D", choices=["B", "C", "D", "Dcyl"],
                    help="D (default) is the headline configuration; B and C are extra data points")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
