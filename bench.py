#!/usr/bin/env python
"""bench.py - agent-updates/s and ms/tick of the per-tick agent update (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3_1m] [--agents A] [--impl ours|reference]

One JSON line on stdout (rank 0).  Workload at N=1: C3, 1M agents in the 2025-block city lattice
(SURVEY.md §8d); a "step" is one Simulator::Update tick over the whole crowd.

  value        whole-job agent-updates/s with all state resident in HBM (CUDA events on the
               simulator's stream, max over ranks), timed on the CONGESTED crowd: the tick gets dearer
               for the first ~400 ticks (more constraints per agent, LP3D), so the K timed ticks start at
               tick --steady-tick (600); the window is repeated --repeats times, `ms_per_step` is the
               median window and `windows_ms_per_step` lists all of them.  `from_rest` is the same K
               ticks right after the warm-up (the cheap phase), reported beside it
  e2e          same metric through the C ABI with HOST buffers (ecmgpu_update_io): every tick uploads
               positions and velocities from pinned memory, runs the tick and downloads positions,
               velocities and active flags; transfers of consecutive ticks overlap with compute.
               With strips (N > 1) every rank moves the (slot, position, velocity) records of the
               agents it owns (ecmgpu_update_io_owned), so the bytes per tick do not grow with N
  roofline     dominant kernel (by measured phase time) against the measured HBM copy bandwidth
  cpu_baseline the unmodified reference (oracle/_ref) on one host core, bounded sample

--impl reference times the reference's own CPU Simulator::Update (single-threaded, as shipped) on a sample of the
same workload sized to --cpu-budget seconds; `cpu_baseline.replica_farm` adds what one such process per host core
delivers together (a labelled upper bound: the reference cannot spread ONE crowd over cores).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ecmgenerator_b200 import scenarios as S  # noqa: E402
from ecmgenerator_b200 import host  # noqa: E402

LAST_SETUP = {}  # filled by build_workload: who planned the routes and how long it took (reported, not timed as the metric)
METRIC = "agent_updates_per_s"
UNIT = "agent-updates/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
def _plan(w, start, goal, radius, threads: int, planner: str, device: int):
    """planner "host": csrc/host/planner.cpp on `threads` cores; "device": ecmgpu_plan_paths on this rank's GPU
    (same polylines bit for bit, tests/test_zz4_gpu_planner.py)."""
    if planner == "device":
        from ecmgenerator_b200 import gpu

        sim = gpu.GpuSim(w, 8, float(S.DT), device=device)
        try:
            return sim.plan_paths(start, goal, radius)
        finally:
            sim.close()
    return host.plan_paths(w, start, goal, radius, threads=threads)


def _plan_sharded(w, c, rank: int, world: int, planner: str = "host", device: int = 0):
    """Every rank plans 1/world of the paths (host threads are shared by the ranks), then all-gather."""
    import torch.distributed as dist

    n = c.n
    lo, hi = n * rank // world, n * (rank + 1) // world
    threads = max(1, (os.cpu_count() or 8) // world)
    off, pxy, _ = _plan(w, c.pos[lo:hi], c.goal[lo:hi], c.radius[lo:hi], threads, planner, device)
    parts = [None] * world
    dist.all_gather_object(parts, (np.diff(off).astype(np.int32), pxy))
    lens = np.concatenate([p[0] for p in parts])
    pts = np.concatenate([p[1] for p in parts])
    out_off = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=out_off[1:])
    return out_off, pts


def _workload_stamp() -> str:
    import hashlib

    h = hashlib.sha1()
    base = os.path.join(ROOT, "ecmgenerator_b200")
    for rel in ("scenarios.py", "csrc/host/lattice_world.cpp", "csrc/host/planner.cpp", "csrc/host/planner.h"):
        with open(os.path.join(base, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:10]


def build_workload(config: str, agents: int | None, rank: int = 0, world: int = 1, planner: str = "host", device: int = 0):
    world_fn, crowd_fn = S.CONFIGS[config]
    w = world_fn()
    t = time.time()
    # ECM_WORKLOAD_CACHE=<dir>: measurement sessions that build the same workload several times (tools/gpu_session.sh)
    # keep the sampled crowd and its planned routes on disk; worlds and crowds are seeded, so the content is the same
    # The sampled crowd and its planned routes are kept on disk (seeded, so the content is what a fresh build gives; the
    # file name carries a hash of the sources that define it): ECM_WORKLOAD_CACHE=<dir>, default ./workloads if it
    # exists (git-ignored, travels to the GPU box), ECM_WORKLOAD_CACHE= (empty) to switch it off.  Set-up only: the
    # planning time is outside every timed region either way.
    cache = os.environ.get("ECM_WORKLOAD_CACHE", os.path.join(ROOT, "workloads") if os.path.isdir(os.path.join(ROOT, "workloads")) else "")
    cache_file = os.path.join(cache, f"{config}_{agents or 0}_{_workload_stamp()}.npz") if cache else None  # every rank reads it; rank 0 writes it
    if cache_file and os.path.exists(cache_file):
        z = np.load(cache_file)
        c = S.Crowd(z["pos"], z["goal"], z["radius"], z["speed"])
        log(f"[bench] {config}: {c.n} agents and their routes from {cache_file} in {time.time() - t:.1f}s")
        return w, c, z["off"], z["pxy"]
    c = crowd_fn(w, n=agents) if agents else crowd_fn(w)
    t1 = time.time()
    if world > 1:
        off, pxy = _plan_sharded(w, c, rank, world, planner, device)
    else:
        off, pxy, _ = _plan(w, c.pos, c.goal, c.radius, 0, planner, device)
    lens = np.diff(off)
    good = lens >= 2
    if not good.all():  # drop agents the planner could not serve (start or goal level with a cell corner)
        keep = np.nonzero(good)[0]
        c = c.take(keep)
        new_off = np.zeros(len(keep) + 1, np.int32)
        np.cumsum(lens[keep], out=new_off[1:])
        pxy = np.concatenate([pxy[off[i]:off[i + 1]] for i in keep]) if len(keep) < 200_000 else _gather(pxy, off, keep)
        off = new_off
    LAST_SETUP.update(planner=planner, plan_seconds=round(time.time() - t1, 2), queries=int(len(lens)))
    log(f"[bench] {config}: {c.n} agents sampled in {t1 - t:.1f}s, paths planned on the {planner} in {time.time() - t1:.1f}s "
        f"(mean {np.diff(off).mean():.1f} points, {int((~good).sum())} dropped)")
    if cache_file and rank == 0:
        os.makedirs(cache, exist_ok=True)
        np.savez(cache_file, pos=c.pos, goal=c.goal, radius=c.radius, speed=c.speed, off=off, pxy=pxy)
    return w, c, off, pxy


def _gather(pxy, off, keep):
    lens = (off[1:] - off[:-1])[keep]
    idx = np.repeat(off[:-1][keep], lens) + (np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens))
    return pxy[idx]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class _silence_stdout:
    """Redirects file descriptor 1 to /dev/null (C++ iostream output of the reference included)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)
        return False


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed ncu --set full
    capture (profiles/*_traffic.json); bench.py itself never runs under a profiler."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        k = d["kernels"][kernel]
        mb = k["dram_read_mbytes"] + k["dram_write_mbytes"]
        return mb * 1e6, os.path.basename(files[-1])
    except Exception:
        return None, None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def _cpu_sample(w, c, off, pxy, n):
    """Spatially compact sample: the n agents closest to the crowd's centroid keep the local density."""
    ctr = c.pos.mean(axis=0)
    idx = np.argsort(((c.pos - ctr) ** 2).sum(axis=1), kind="stable")[:n]
    idx.sort()
    sub = c.take(idx)
    sub_off = np.zeros(n + 1, np.int32)
    lens = (off[1:] - off[:-1])[idx]
    np.cumsum(lens, out=sub_off[1:])
    return sub, sub_off, _gather(pxy, off, idx)


def _cpu_run(w, sub, sub_off, sub_xy, ticks, warm):
    """(seconds for `ticks` ticks, kind): the unmodified reference (oracle/_ref) if it was built, else the C port."""
    from oracle import pyref

    n = sub.n
    with _silence_stdout():  # the reference prints banners and timing tables to stdout
        if pyref.available("ref-kdtree"):
            kind = "reference"
            sim = pyref.RefSim(w, n + 8, float(S.DT), "ref-kdtree")
            sim.bulk_load(sub.pos, None, sub.radius, sub.speed, sub_off, sub_xy)
        else:
            from oracle.pyoracle import OracleSim

            kind = "port"
            sim = OracleSim(w, n + 8, float(S.DT), "ref-kdtree")
            sim.bulk_load(sub.pos, sub.radius, sub.speed, sub_off, sub_xy)
        for _ in range(warm):
            sim.step(1)
        t0 = time.perf_counter()
        for _ in range(ticks):
            sim.step(1)
        dt = time.perf_counter() - t0
        sim.close()
    return dt, kind


def _cpu_replica(args):
    return _cpu_run(*args)[0]


def cpu_reference_run(w, c, off, pxy, sample_agents: int, ticks: int, warm: int, budget_s: float | None = None, replicas: int = 0):
    """The reference's Simulator::Update on a window of the same workload, one host core (it has no threading).

    budget_s: shrink the sample so that warm + ticks ticks take about that long (the reference scans all ECM cells
    and all obstacle segments per agent: its cost per agent-update is flat in the sample size).
    replicas > 0: additionally run that many independent copies of the same sample side by side, one per core -
    what a farm of reference processes would deliver on this box (labelled upper bound)."""
    n = min(sample_agents, c.n) if sample_agents > 0 else c.n
    if budget_s is not None:
        probe = _cpu_sample(w, c, off, pxy, min(256, n))
        per_agent = _cpu_run(w, *probe, 1, 0)[0] / probe[0].n
        n = int(max(64, min(n, budget_s / ((ticks + warm) * per_agent))))
    sub, sub_off, sub_xy = _cpu_sample(w, c, off, pxy, n)
    dt, kind = _cpu_run(w, sub, sub_off, sub_xy, ticks, warm)
    res = {"value": n * ticks / dt, "unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count(), "same_config": bool(n == c.n),
           "sample": f"{n} agents nearest the crowd centroid of the same world ({w.n_cells} ECM cells, {w.n_obst_vertices} obstacle segments), "
                     f"{ticks} ticks after {warm} warm-up", "ms_per_tick_sample": 1e3 * dt / ticks}
    if replicas > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(replicas) as pool:  # no CUDA context in this process (reference arm only)
            t0 = time.perf_counter()
            dts = pool.map(_cpu_replica, [(w, sub, sub_off, sub_xy, ticks, warm)] * replicas)
            wall = time.perf_counter() - t0
        res["replica_farm"] = {"replicas": replicas, "value": replicas * n * ticks / max(dts), "unit": UNIT, "wall_s": wall,
                               "note": "independent single-threaded copies of the sample, one per core: an upper bound for this box, "
                                       "not something the reference can do for ONE crowd"}
    return res


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, c, off, pxy = build_workload(args.config, args.agents)
    # sized so that the whole run stays near a minute whatever --steps asks for; then the same once more on every core
    res = cpu_reference_run(w, c, off, pxy, args.cpu_sample, max(1, args.steps), max(0, min(args.warmup, 2)), budget_s=args.cpu_budget,
                            replicas=min(os.cpu_count() or 1, 128))
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_tick_sample"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.config}: {c.n} agents, {w.n_obstacles} blocks, {w.n_cells} ECM cells; each step = one tick of a bounded sample",
                       "agents": c.n, "dt": float(S.DT)},
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def strips_parity(sim, w, c, off, pxy, ticks, rank, device):
    """Multi-GPU parity, visible in the bench line: `ticks` ticks on the strips and on ONE GPU (rank 0, a fresh single
    simulator) from the same global state, compared bit for bit.  Collective: every rank calls it."""
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200.multigpu import REBALANCE_STATE, owner_of

    sim.sync()
    state = {k: sim.gather(k)[0] for k in REBALANCE_STATE}
    owners = sim.gather(gpu.ACTIVE)[1]
    own0 = owner_of(state[gpu.POS][:, 0], sim.bounds)
    miss0 = sim.global_stats(("halo_misses",))["halo_misses"]
    sim.update(ticks)
    sim.sync()
    pos1, owners1 = sim.gather(gpu.POS)
    vel1, _ = sim.gather(gpu.VEL)
    miss1 = sim.global_stats(("halo_misses",))["halo_misses"]
    if rank != 0:
        return None
    n = c.n
    one = gpu.GpuSim(w, n, float(S.DT), device=device, record_neighbors=False, neighbor_cell=float(sim.stats()["neighbor_cell"]),
                     path_pool_points=int(off[-1] * 1.25) + 4096)
    one.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    for k, a in state.items():
        one.write(k, a)
    one.write(gpu.ACTIVE, owners)
    one.update(ticks)
    one.sync()
    p, v, a = one.read(gpu.POS, 0, n), one.read(gpu.VEL, 0, n), one.read(gpu.ACTIVE, 0, n) > 0
    one.close()
    same_owner = bool(np.array_equal(owners1 > 0, a)) and int(owners1.max()) <= 1
    same_pos = bool(np.array_equal(pos1[a].view(np.uint32), p[a].view(np.uint32)))
    same_vel = bool(np.array_equal(vel1[a].view(np.uint32), v[a].view(np.uint32)))
    own1 = owner_of(pos1[:, 0], sim.bounds)
    inner = np.asarray(sim.bounds[1:-1], np.float32)
    near = int((np.abs(pos1[a][:, :1] - inner[None, :]).min(axis=1) < 1.0).sum()) if len(inner) else 0
    dx = np.abs(pos1[a][:, 0] - state[gpu.POS][a][:, 0])
    return {"bitwise_vs_1gpu": same_owner and same_pos and same_vel, "ticks": int(ticks), "migrations": int(((own0 != own1) & a).sum()),
            "agents_within_1m_of_a_border": near, "mean_abs_dx_m": float(dx.mean()),
            "agents_compared": int(a.sum()), "halo_misses_during": int(miss1 - miss0),
            "rows_differing": int((pos1[a].view(np.uint32) != p[a].view(np.uint32)).any(axis=1).sum()),
            "how": "global state gathered from the strips, loaded into a single simulator on rank 0; both advanced, positions / velocities / ownership compared bit for bit"}


def run_ours(args):
    import torch

    from ecmgenerator_b200 import gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, c, off, pxy = build_workload(args.config, args.agents, rank, world, args.planner, local)
    n = c.n
    if world > 1:
        from ecmgenerator_b200.multigpu import StripSim

        sim = StripSim(w, c, off, pxy, rank, world, local)
    else:
        sim = gpu.GpuSim(w, n, float(S.DT), device=local, record_neighbors=False, neighbor_cell=args.cell, static_bin=args.bin,
                         path_pool_points=int(off[-1] * 1.25) + 4096)
        sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
        if args.neighbors == "kdtree":
            sim.set_neighbor_mode(gpu.NEIGHBORS_KDTREE)
    if args.neighbors == "kdtree" and world > 1:
        raise SystemExit("--neighbors kdtree runs on one GPU")
    mean_p = float(np.diff(off).mean())

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        sim.sync()

    # ---- warm-up (also builds bins / grid)
    for _ in range(max(args.warmup, 3)):
        sim.update(1)
    sim.sync()
    raw0 = sim if world == 1 else sim.sim

    def timed_window(steps):
        """K ticks between two stream marks, barrier + synchronize on both sides; max over ranks."""
        barrier()
        sim.mark(0)
        for _ in range(steps):
            sim.update(1)
        sim.mark(1)
        barrier()
        t_ms = sim.elapsed_ms(0, 1)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([t_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms

    def active_now():
        return float(sim.stats()["n_active"]) if world == 1 else float(sim.global_active())

    def phase_pass(reps):
        """Per-phase CUDA-event times, averaged over `reps` ticks of the current state.  The events are event-record nodes
        inside the captured tick (csrc/ecmgpu.cu: ecmgpu_set_profiling), i.e. the phases of the same graph the timed windows
        replay - not of a launch-by-launch tick with its host gaps.  (Strips over NCCL and the KD-tree mode are launch by
        launch either way.)"""
        sim.set_profiling(True)
        acc_ = {"grid": 0.0, "attract": 0.0, "orca": 0.0, "fallback": 0.0, "tick": 0.0}  # orca = k_orca alone, fallback = k_fallback
        for _ in range(reps):
            sim.update(3)  # back to back like the timed windows; the events of the last tick are read (each tick overwrites them)
            t_ms = sim.last_tick_phases()
            for k_ in acc_:
                acc_[k_] += t_ms[k_] / reps
        sim.set_profiling(False)
        return acc_

    # ---- disclosure: K ticks from the crowd at rest (the cheap phase of the run)
    clocks = ClockSampler(local)
    clocks.start()
    done = max(args.warmup, 3)
    from_rest = None
    if args.steady_tick > done:
        a_r = active_now()
        st_a = sim.stats()
        ms_r = timed_window(args.steps)
        st_b = sim.stats()
        done += args.steps
        rest_phases = None
        if world == 1:
            rest_phases = phase_pass(10)
            done += 30
        from_rest = {"from_tick": int(done - args.steps - (30 if world == 1 else 0)), "ms_per_step": ms_r / args.steps, "value": a_r * args.steps / (ms_r * 1e-3), "unit": UNIT,
                     "phase_ms": rest_phases,
                     "lp3d_runs_per_tick_rank0": (st_b["lp3d_runs"] - st_a["lp3d_runs"]) / args.steps,
                     "note": "the same K ticks right after the warm-up, crowd at rest: cheaper than the congested crowd the headline is timed on"}
        # ---- let the crowd congest (untimed): the tick cost levels off after ~400 ticks (profiles/r01_experiments.md)
        while done < args.steady_tick:
            sim.update(min(50, args.steady_tick - done))
            done += min(50, args.steady_tick - done)
        sim.sync()
    st0 = sim.stats()
    active0 = st0["n_active"]
    # the state the timed ticks start from: the end-to-end loop below replays it, so that `value` and `e2e` time the
    # same phase of the simulation
    pos_w, vel_w, act_w = raw0.read(gpu.POS, 0, n), raw0.read(gpu.VEL, 0, n), raw0.read(gpu.ACTIVE, 0, n) > 0

    # ---- timed: K ticks, state resident in HBM, repeated; the median window is the headline
    windows, win_active = [], []
    for _ in range(max(1, args.repeats)):
        win_active.append(active_now())
        windows.append(timed_window(args.steps))
    order = sorted(range(len(windows)), key=lambda i: windows[i])
    mid = order[len(order) // 2]
    ms = windows[mid]
    st1 = sim.stats()
    launches = (st1["kernel_launches"] - st0["kernel_launches"]) // max(1, args.repeats)
    updates = win_active[mid] * args.steps
    value = updates / (ms * 1e-3)
    lp3d_per_tick = (st1["lp3d_runs"] - st0["lp3d_runs"]) / (args.steps * max(1, args.repeats))
    timed_from = int(done)
    done += args.steps * max(1, args.repeats)

    # ---- per-phase times for the roofline of the dominant kernel (separate pass, same state)
    roof = None
    if world == 1:
        acc = phase_pass(max(3, min(args.steps, 20)))
        peak, peak_src = measured_hbm_peak()
        # algorithmic bytes per agent-update (DESIGN.md "Roofline"): whole tick 176 + 8 P;
        # k_attract 40 + 8 P (pos, speed, path header + polyline; attraction + prefvel out),
        # k_orca 136 (own vel/radius, 5 x (pos,vel,radius) neighbours; pos, vel, force out)
        alg = {"attract": 40.0 + 8.0 * mean_p, "orca": 136.0, "tick": 176.0 + 8.0 * mean_p}
        dom = "orca" if acc["orca"] >= acc["attract"] else "attract"
        ach = alg[dom] * active0 / (acc[dom] * 1e-3) / 1e9
        ach_tick = alg["tick"] * active0 / ((ms / args.steps) * 1e-3) / 1e9  # whole tick: the timed region itself (graph launch)
        kname = "k_" + dom
        traffic, traffic_src = ncu_traffic(kname)
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_agent": alg[dom],
                "kernel_ms": acc[dom], "phase_ms": acc,
                "whole_tick": {"achieved": ach_tick, "frac": ach_tick / peak, "algorithmic_bytes_per_agent": alg["tick"]},
                "state": f"congested crowd (from tick {timed_from}), like the headline",
                # the five event-record nodes drain the GPU between the phases: the profiled tick runs this much longer than
                # the timed windows' tick, so kernel_ms errs on the slow side (frac on the low side) by at most this
                "phase_events_overhead_ms": acc["tick"] - ms / args.steps}
        if from_rest and from_rest.get("phase_ms"):  # the same kernel on the crowd at rest (what round 1's line reported)
            k_ms = from_rest["phase_ms"][dom]
            roof["from_rest"] = {"kernel_ms": k_ms, "achieved": alg[dom] * active0 / (k_ms * 1e-3) / 1e9,
                                 "frac": alg[dom] * active0 / (k_ms * 1e-3) / 1e9 / peak}
    # ---- end to end through the C ABI with host buffers.  One GPU: whole slot arrays (ecmgpu_update_io);
    # strips: every rank moves the records of the agents it owns (ecmgpu_update_io_owned)
    e2e = None
    DEPTH = 3
    raw = sim if world == 1 else sim.sim
    k = max(3, min(args.steps, 50))
    if world == 1:
        # DEPTH generations of pinned host buffers: call i uses set i % DEPTH while the sets of i-1 and i-2 are still in flight
        hp = [gpu.PinnedArray((n, 2), np.float32) for _ in range(DEPTH)]
        hv = [gpu.PinnedArray((n, 2), np.float32) for _ in range(DEPTH)]
        op = [gpu.PinnedArray((n, 2), np.float32) for _ in range(DEPTH)]
        ov = [gpu.PinnedArray((n, 2), np.float32) for _ in range(DEPTH)]
        oa = [gpu.PinnedArray((n,), np.uint8) for _ in range(DEPTH)]
        bufs = hp + hv + op + ov + oa
        for g in range(DEPTH):
            hp[g].array[:] = pos_w
            hv[g].array[:] = vel_w
        moved = [0, 0]

        def io_call(i):
            g = i % DEPTH
            moved[0] += 16 * n
            moved[1] += 17 * n
            return raw.update_io(n, hp[g], hv[g], op[g], ov[g], oa[g])  # H2D pos+vel | tick | D2H pos+vel+active

        def io_result(i):
            # the host reads the whole result (one flag per slot).  np.count_nonzero, not `(a > 0).sum()`: the latter
            # builds a 1 MB temporary and sums it as int64 - 0.08 ms of host time per tick, which made the host thread that
            # also issues the calls the bottleneck of the loop (tools/e2e_probe.py: io_both 0.437 vs io_both_consume 0.521)
            return int(np.count_nonzero(oa[i % DEPTH].array))
    else:
        act_now = raw.read(gpu.ACTIVE, 0, n) > 0
        ids = np.flatnonzero(act_now)
        p0, v0 = raw.read(gpu.POS, 0, n), raw.read(gpu.VEL, 0, n)
        keep_w = act_now & act_w  # owned at the start of the timed ticks too: their state of that moment is replayed
        p0[keep_w], v0[keep_w] = pos_w[keep_w], vel_w[keep_w]
        cap = n
        rin = [gpu.PinnedArray((cap,), gpu.AGENT_REC) for _ in range(DEPTH)]
        rout = [gpu.PinnedArray((cap,), gpu.AGENT_REC) for _ in range(DEPTH)]
        cnt = [gpu.PinnedArray((1,), np.int32) for _ in range(DEPTH)]
        bufs = rin + rout + cnt
        for g in range(DEPTH):
            r = rin[g].array
            r["slot"][: len(ids)] = ids
            r["x"][: len(ids)], r["y"][: len(ids)] = p0[ids, 0], p0[ids, 1]
            r["vx"][: len(ids)], r["vy"][: len(ids)] = v0[ids, 0], v0[ids, 1]
        moved = [0, 0]
        rec_b = gpu.AGENT_REC.itemsize

        def io_call(i):
            g = i % DEPTH
            moved[0] += rec_b * len(ids)
            return raw.update_io_owned(len(ids), rin[g], rout[g], cnt[g])  # H2D owned records | tick | D2H owned records

        def io_result(i):
            m = int(cnt[i % DEPTH].array[0])
            moved[1] += rec_b * m + 4  # lower bound: the copy is sized from the last confirmed count plus migrant room
            return m

    def e2e_run(steps):
        # the host keeps DEPTH - 1 calls in flight: it issues call i, then consumes the results of call i - (DEPTH - 1)
        tickets, res = [], 0
        for i in range(steps):
            tickets.append(io_call(i))
            j = i - (DEPTH - 1)
            if j >= 0:
                raw.io_wait(tickets[j])
                res = io_result(j)
        for j in range(max(0, steps - (DEPTH - 1)), steps):
            raw.io_wait(tickets[j])
            res = io_result(j)
        return res

    e2e_run(4)
    moved[0] = moved[1] = 0
    barrier()
    t0 = time.perf_counter()
    act = e2e_run(k)
    barrier()
    dt = time.perf_counter() - t0
    h2d, d2h = moved[0] / k, moved[1] / k
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([dt, float(act), h2d, d2h], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, act, h2d, d2h = float(tmax[0].item()), int(t[1].item()), float(t[2].item()), float(t[3].item())
    probe = None
    if world > 1 and os.environ.get("ECM_E2E_PROBE"):  # where the strips' end-to-end tick goes: the same loop without the upload
        full_call = io_call

        def io_call(i):  # noqa: F811
            return raw.update_io_owned(0, None, rout[i % DEPTH], cnt[i % DEPTH])

        e2e_run(4)
        barrier()
        t1 = time.perf_counter()
        e2e_run(k)
        barrier()
        probe = {"no_upload_ms_per_step": 1e3 * (time.perf_counter() - t1) / k}
        io_call = full_call
    e2e = {"value": act * k / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": 1e3 * dt / k, "steps": k, "calls_in_flight": DEPTH - 1,
           "api": "ecmgpu_update_io (whole slot arrays)" if world == 1 else "ecmgpu_update_io_owned (records of the agents each rank owns)"}
    if probe:
        e2e["probe"] = probe
    for a in bufs:
        a.free()
    # the host link as this box gives it (pinned memory, one direction at a time): context for the e2e number
    lk = gpu.PinnedArray((n, 2), np.float32)
    raw.read_async(gpu.POS, lk, 0, n)
    raw.sync()
    raw.mark(0)
    for _ in range(4):
        raw.read_async(gpu.POS, lk, 0, n)
    raw.mark(1)
    for _ in range(4):
        raw.write_async(gpu.POS, lk, 0, n)  # the values just read: the state does not change
    raw.mark(2)
    raw.sync()
    e2e["link"] = {"d2h_GBps": 4 * 8 * n / (raw.elapsed_ms(0, 1) * 1e6), "h2d_GBps": 4 * 8 * n / (raw.elapsed_ms(1, 2) * 1e6)}
    lk.free()
    # sampled over all timed regions (resident ticks, per-phase pass, end-to-end ticks)
    clk = clocks.stop()

    global_active = int(active0)
    parity = None
    if world > 1:
        import torch.distributed as dist

        if args.parity_ticks > 0:
            parity = strips_parity(sim, w, c, off, pxy, args.parity_ticks, rank, local)
        global_active = sim.global_active()
        # whole-job counters; zero halo misses = every owned agent's 5-NN ball stayed inside what its rank sees,
        # i.e. the strips computed exactly what one GPU computes (DESIGN.md "Multi-GPU")
        job_counters = sim.global_stats(("halo_misses", "knn_fallbacks", "obstacle_overflows", "lp3d_runs", "location_failures", "replans"))
        if job_counters["halo_misses"] > 0:  # loud: from the first miss on the strips do not compute what one GPU computes
            log(f"[bench] WARNING: {job_counters['halo_misses']} halo misses - the multi-GPU result is NOT the single-GPU result; widen the halo")
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        t = time.time()
        cpu = cpu_reference_run(w, c, off, pxy, args.cpu_sample, args.cpu_ticks, 1, budget_s=args.cpu_inline_budget)
        log(f"[bench] cpu baseline in {time.time() - t:.1f}s")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {n} agents, {w.n_obstacles} blocks, {w.n_cells} ECM cells, dt=1/60, mean path points {mean_p:.1f}",
                       "agents": n, "active": global_active, "dt": float(S.DT), "parallelism": f"strips{world}" if world > 1 else "1gpu",
                       "l2": "working set (agent state + path pool + snapshot) exceeds the 126 MB L2", "neighbor_cell": st1["neighbor_cell"],
                       "static_bin": st1["static_bin"]},
            "gpu_launches": int(launches), "clocks": clk,
            "counters": {k: int(st1[k]) for k in ("knn_fallbacks", "obstacle_overflows", "lp3d_runs", "location_failures", "replans")}}
    if LAST_SETUP:
        line["setup"] = dict(LAST_SETUP, note="route planning before the run; outside every timed region")
    if args.neighbors == "kdtree":
        line["config"]["neighbors"] = "kdtree (the reference's own lists; parity mode)"
        line["counters"].update({k: int(st1[k]) for k in ("kd_median_ties", "kd_small_ties")})
    if world > 1:
        line["counters"] = dict(job_counters, scope="all ranks, whole run")
        line["config"]["halo_m"] = float(sim.halo)
    if roof:
        line["roofline"] = roof
    if e2e:
        line["e2e"] = e2e
    line["windows_ms_per_step"] = [round(x / args.steps, 6) for x in windows]
    line["window"] = {"from_tick": timed_from, "repeats": len(windows), "reported": "median window", "min_ms_per_step": min(windows) / args.steps,
                      "max_ms_per_step": max(windows) / args.steps, "lp3d_runs_per_tick_rank0": lp3d_per_tick}
    if from_rest:
        line["from_rest"] = from_rest
    if parity:
        line["parity"] = parity
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu-budget", type=float, default=45.0, help="--impl reference: seconds of CPU time the timed sample may take")
    ap.add_argument("--steady-tick", type=int, default=600,
                    help="the timed ticks start at this tick, on the congested crowd (0: right after the warm-up, crowd at rest)")
    ap.add_argument("--repeats", type=int, default=5, help="how many times the K-tick window is timed (the median is reported)")
    ap.add_argument("--parity-ticks", type=int, default=40,
                    help="N > 1: after the timing, this many ticks on the strips AND on rank 0 alone from the same state, compared bit for bit (0 = skip)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3_1m", choices=sorted(S.CONFIGS))
    ap.add_argument("--agents", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="agents in the CPU baseline sample (0: the whole crowd if the time budget allows, else as many as fit)")
    ap.add_argument("--cpu-ticks", type=int, default=2)
    ap.add_argument("--cpu-inline-budget", type=float, default=20.0, help="seconds the cpu_baseline leg of our own line may take")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cell", type=float, default=0.0, help="neighbour grid cell (0 = from crowd density)")
    ap.add_argument("--bin", type=float, default=0.0, help="static bin edge (0 = from the ECM)")
    ap.add_argument("--planner", default="host", choices=["host", "device"],
                    help="who plans the agents' routes during set-up (outside the timed region): the host planner on all cores or "
                         "ecmgpu_plan_paths on the GPU - identical polylines")
    ap.add_argument("--neighbors", default="exact", choices=["exact", "kdtree"],
                    help="kdtree: the reference's own KD-tree lists (parity mode, single GPU; DESIGN.md 5a) - a cost figure, not the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
