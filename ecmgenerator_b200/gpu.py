"""ctypes binding of the CUDA C ABI (include/ecm_b200.h, libecmgpu.so).

`GpuSim` is a thin object wrapper with the same surface the parity tests use on the oracle side
(bulk_load / step / state / query_*), so tests read alike on both sides.  There is no CPU
fallback: constructing a GpuSim without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ECMGPU_LIB") or os.path.join(_HERE, "libecmgpu.so")  # ECMGPU_LIB: build variants for experiments
_lib = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)

# selectors (include/ecm_b200.h)
POS, VEL, PREFVEL, ATTRACTION, FORCE, RADIUS, SPEED, ACTIVE, CELL, NEIGHBORS, NEIGHBOR_COUNT, STATUS, REPLAN_PENDING = range(13)
NEIGHBORS_EXACT, NEIGHBORS_KDTREE = 0, 1
ST_NO_CELL, ST_REPLAN, ST_ARRIVING, ST_DESTROYED, ST_OBST_OVERFLOW, ST_KNN_FALLBACK, ST_LP3D, ST_HALO_MISS = (
    1, 2, 4, 8, 16, 32, 64, 128)

_DTYPES = {
    POS: (np.float32, 2), VEL: (np.float32, 2), PREFVEL: (np.float32, 2), ATTRACTION: (np.float32, 2),
    FORCE: (np.float32, 2), RADIUS: (np.float32, 1), SPEED: (np.float32, 1), ACTIVE: (np.uint8, 1),
    CELL: (np.int32, 1), NEIGHBORS: (np.int32, 5), NEIGHBOR_COUNT: (np.int32, 1), STATUS: (np.uint32, 1),
    REPLAN_PENDING: (np.uint8, 1),
}

EXPORTS = [
    "ecmgpu_create", "ecmgpu_destroy", "ecmgpu_last_error", "ecmgpu_set_ecm", "ecmgpu_set_obstacles", "ecmgpu_spawn",
    "ecmgpu_bulk_load", "ecmgpu_set_path", "ecmgpu_destroy_agent", "ecmgpu_update", "ecmgpu_sync", "ecmgpu_poll_events",
    "ecmgpu_read", "ecmgpu_write", "ecmgpu_read_async", "ecmgpu_write_async", "ecmgpu_alloc_pinned", "ecmgpu_free_pinned",
    "ecmgpu_locate", "ecmgpu_retract", "ecmgpu_find_neighbors", "ecmgpu_find_obstacles", "ecmgpu_get_stats",
    "ecmgpu_last_tick_ms", "ecmgpu_last_tick_phases", "ecmgpu_set_profiling", "ecmgpu_mark", "ecmgpu_mark_elapsed_ms", "ecmgpu_stream", "ecmgpu_comm_unique_id", "ecmgpu_comm_init",
    "ecmgpu_comm_set_strips", "ecmgpu_comm_init_local", "ecmgpu_update_phase", "ecmgpu_update_io", "ecmgpu_update_io_owned", "ecmgpu_io_wait", "ecmgpu_comm_p2p_export", "ecmgpu_comm_p2p_connect",
    "ecmgpu_set_neighbor_mode", "ecmgpu_valid_spawn_locations", "ecmgpu_draw_spawns", "ecmgpu_set_ecm_topology", "ecmgpu_plan_paths", "ecmgpu_plan_info",
    "ecmgpu_abi_sizes",
]


class Params(C.Structure):
    _fields_ = [("device", C.c_int), ("max_agents", C.c_int), ("step", C.c_float), ("neighbor_cell", C.c_float),
                ("static_bin", C.c_float), ("max_obstacle_range", C.c_float), ("path_pool_points", C.c_int),
                ("record_neighbors", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("n_slots", C.c_int), ("n_active", C.c_int), ("grid_w", C.c_int), ("grid_h", C.c_int),
                ("neighbor_cell", C.c_float), ("bins_w", C.c_int), ("bins_h", C.c_int), ("static_bin", C.c_float),
                ("max_cell_list", C.c_int), ("max_obstacle_list", C.c_int), ("ticks", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("knn_fallbacks", C.c_uint64), ("obstacle_overflows", C.c_uint64),
                ("lp3d_runs", C.c_uint64), ("location_failures", C.c_uint64), ("replans", C.c_uint64),
                ("halo_misses", C.c_uint64),
                ("kd_median_ties", C.c_uint64), ("kd_small_ties", C.c_uint64), ("event_overflows", C.c_uint64), ("nonfinite_agent_ticks", C.c_uint64)]


class EcmGpuError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libecmgpu.so (fails loudly if it was not built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EcmGpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ecmgpu_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
        L.ecmgpu_destroy.argtypes = [vp]
        L.ecmgpu_last_error.restype = C.c_char_p
        L.ecmgpu_last_error.argtypes = [vp]
        L.ecmgpu_set_ecm.argtypes = [vp, f32p, C.c_int, f32p, f32p, C.c_int, i32p, f32p]
        L.ecmgpu_set_obstacles.argtypes = [vp, C.c_int, f32p, i32p, i32p, u8p]
        L.ecmgpu_spawn.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_int]
        L.ecmgpu_bulk_load.argtypes = [vp, C.c_int, i32p, f32p, f32p, f32p, i32p, f32p]
        L.ecmgpu_set_path.argtypes = [vp, C.c_int, f32p, C.c_int]
        L.ecmgpu_destroy_agent.argtypes = [vp, C.c_int]
        L.ecmgpu_update.argtypes = [vp]
        L.ecmgpu_sync.argtypes = [vp]
        L.ecmgpu_poll_events.argtypes = [vp, i32p, C.c_int, i32p, i32p, C.c_int, i32p]
        for name in ("ecmgpu_read", "ecmgpu_write", "ecmgpu_read_async", "ecmgpu_write_async"):
            getattr(L, name).argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
        L.ecmgpu_alloc_pinned.restype = vp
        L.ecmgpu_alloc_pinned.argtypes = [C.c_uint64]
        L.ecmgpu_free_pinned.argtypes = [vp]
        L.ecmgpu_locate.argtypes = [vp, C.c_int, f32p, i32p]
        L.ecmgpu_retract.argtypes = [vp, C.c_int, f32p, u8p, f32p, i32p]
        L.ecmgpu_find_neighbors.argtypes = [vp, C.c_int, i32p, i32p]
        L.ecmgpu_find_obstacles.argtypes = [vp, C.c_int, i32p, C.c_int, i32p]
        L.ecmgpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.ecmgpu_last_tick_ms.argtypes = [vp, f32p]
        L.ecmgpu_last_tick_phases.argtypes = [vp, f32p]
        L.ecmgpu_set_profiling.argtypes = [vp, C.c_int]
        L.ecmgpu_mark.argtypes = [vp, C.c_int]
        L.ecmgpu_mark_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, f32p]
        L.ecmgpu_stream.restype = vp
        L.ecmgpu_stream.argtypes = [vp]
        L.ecmgpu_comm_unique_id.argtypes = [u8p]
        L.ecmgpu_comm_init.argtypes = [vp, u8p, C.c_int, C.c_int]
        L.ecmgpu_comm_set_strips.argtypes = [vp, f32p, C.c_float]
        L.ecmgpu_comm_init_local.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.ecmgpu_update_phase.argtypes = [vp, C.c_int]
        L.ecmgpu_comm_p2p_export.argtypes = [vp, u8p]
        L.ecmgpu_comm_p2p_connect.argtypes = [vp, u8p, u8p]
        L.ecmgpu_update_io.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.POINTER(C.c_uint64)]
        L.ecmgpu_io_wait.argtypes = [vp, C.c_uint64]
        L.ecmgpu_set_neighbor_mode.argtypes = [vp, C.c_int]
        L.ecmgpu_valid_spawn_locations.argtypes = [vp, C.c_int, f32p, f32p, u8p]
        L.ecmgpu_draw_spawns.argtypes = [vp, C.c_int, f32p, f32p, f32p, C.c_uint64, C.c_uint64, C.c_int, f32p, f32p, u8p]
        L.ecmgpu_set_ecm_topology.argtypes = [vp, i32p, i32p]
        L.ecmgpu_abi_sizes.argtypes = [i32p]
        L.ecmgpu_abi_sizes.restype = None
        sizes = (C.c_int32 * 4)()
        L.ecmgpu_abi_sizes(sizes)
        mine = (C.sizeof(Params), C.sizeof(Stats), AGENT_REC.itemsize, len(Stats._fields_))
        if tuple(sizes) != mine:  # a stale mirror would corrupt memory in ecmgpu_get_stats / ecmgpu_create
            raise EcmGpuError(f"struct layouts of {LIB_PATH} {tuple(sizes)} differ from this binding's {mine}")
        L.ecmgpu_plan_paths.argtypes = [vp, C.c_int, f32p, f32p, f32p, i32p, i32p, u8p, f32p, C.c_int, i32p]
        L.ecmgpu_update_io_owned.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# ecmgpu_agent_rec (include/ecm_b200.h): 20-byte record of one owned agent
AGENT_REC = np.dtype([("slot", np.int32), ("x", np.float32), ("y", np.float32), ("vx", np.float32), ("vy", np.float32)])


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (ecmgpu_alloc_pinned)."""

    def __init__(self, shape, dtype):
        self.L = lib()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.ecmgpu_alloc_pinned(self.nbytes)
        if not self.ptr:
            raise EcmGpuError("cudaHostAlloc failed")
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.ecmgpu_free_pinned(self.ptr)
            self.ptr = None


class GpuSim:
    """One simulator on one CUDA device (ecmgpu_sim)."""

    def __init__(self, world, max_agents: int, step: float, device: int = 0, neighbor_cell: float = 0.0,
                 static_bin: float = 0.0, max_obstacle_range: float = 0.0, path_pool_points: int = 0,
                 record_neighbors: bool = True):
        self.L = lib()
        self.max_agents = int(max_agents)
        prm = Params(int(device), self.max_agents, float(step), float(neighbor_cell), float(static_bin),
                     float(max_obstacle_range), int(path_pool_points), 1 if record_neighbors else 0)
        h = C.c_void_p()
        rc = self.L.ecmgpu_create(C.byref(prm), C.byref(h))
        if rc != 0:
            raise EcmGpuError(f"ecmgpu_create failed ({rc}): {self.L.ecmgpu_last_error(None).decode()}")
        self.h = h
        w = world
        a = [np.ascontiguousarray(x) for x in (w.bbox, w.vert_xy, w.vert_clear, w.edge_v, w.edge_cl)]
        self._ck(self.L.ecmgpu_set_ecm(self.h, _p(a[0], f32p), w.n_vertices, _p(a[1], f32p), _p(a[2], f32p), w.n_edges,
                                       _p(a[3], i32p), _p(a[4], f32p)))
        b = [np.ascontiguousarray(x) for x in (w.obst_xy, w.obst_next, w.obst_prev, w.obst_convex)]
        self._ck(self.L.ecmgpu_set_obstacles(self.h, w.n_obst_vertices, _p(b[0], f32p), _p(b[1], i32p), _p(b[2], i32p),
                                             _p(b[3], u8p)))
        if getattr(w, "vert_he", None) is not None and getattr(w, "he_next", None) is not None:  # half-edge rings: device planner
            t = [np.ascontiguousarray(x, np.int32) for x in (w.vert_he, w.he_next)]
            rc = self.L.ecmgpu_set_ecm_topology(self.h, _p(t[0], i32p), _p(t[1], i32p))
            # the tick does not need the rings: a graph they do not describe only makes plan_paths unavailable
            self.topology_error = None if rc == 0 else self.L.ecmgpu_last_error(self.h).decode()

    def _ck(self, rc):
        if rc != 0:
            raise EcmGpuError(f"ecmgpu error {rc}: {self.L.ecmgpu_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.ecmgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- agents
    def bulk_load(self, pos, radius, speed, path_off, path_xy, slots=None):
        pos = np.ascontiguousarray(pos, np.float32)
        radius = np.ascontiguousarray(radius, np.float32)
        speed = np.ascontiguousarray(speed, np.float32)
        path_off = np.ascontiguousarray(path_off, np.int32)
        path_xy = np.ascontiguousarray(path_xy, np.float32)
        sl = None if slots is None else np.ascontiguousarray(slots, np.int32)
        self._ck(self.L.ecmgpu_bulk_load(self.h, len(pos), _p(sl, i32p), _p(pos, f32p), _p(radius, f32p), _p(speed, f32p),
                                         _p(path_off, i32p), _p(path_xy, f32p)))
        return np.arange(len(pos), dtype=np.int32) if slots is None else sl

    def spawn(self, slot, pos, radius, speed, path_xy):
        path_xy = np.ascontiguousarray(path_xy, np.float32)
        self._ck(self.L.ecmgpu_spawn(self.h, int(slot), float(pos[0]), float(pos[1]), float(radius), float(speed),
                                     _p(path_xy, f32p), len(path_xy)))

    def set_path(self, slot, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        self._ck(self.L.ecmgpu_set_path(self.h, int(slot), _p(xy, f32p), len(xy)))

    def destroy_agent(self, slot):
        self._ck(self.L.ecmgpu_destroy_agent(self.h, int(slot)))

    # -- ticks
    def update(self, n: int = 1):
        for _ in range(int(n)):
            self._ck(self.L.ecmgpu_update(self.h))

    def update_io(self, count, in_pos, in_vel, out_pos, out_vel, out_active) -> int:
        """Pipelined tick with host I/O; arguments are PinnedArray (or None).  Returns the ticket."""
        ptr = lambda a: C.c_void_p(a.ptr) if a is not None else None  # noqa: E731
        t = C.c_uint64(0)
        self._ck(self.L.ecmgpu_update_io(self.h, int(count), ptr(in_pos), ptr(in_vel), ptr(out_pos), ptr(out_vel), ptr(out_active),
                                         C.byref(t)))
        return int(t.value)

    def update_io_owned(self, n_in, in_rec, out_rec, out_count) -> int:
        """Pipelined tick moving only the owned agents as AGENT_REC records (PinnedArray arguments; in_rec may be None)."""
        ptr = lambda a: C.c_void_p(a.ptr) if a is not None else None  # noqa: E731
        t = C.c_uint64(0)
        self._ck(self.L.ecmgpu_update_io_owned(self.h, int(n_in), ptr(in_rec), ptr(out_rec), int(out_rec.array.shape[0]), ptr(out_count),
                                               C.byref(t)))
        return int(t.value)

    def io_wait(self, ticket: int):
        self._ck(self.L.ecmgpu_io_wait(self.h, int(ticket)))

    def update_phase(self, phase: int):
        self._ck(self.L.ecmgpu_update_phase(self.h, int(phase)))

    def sync(self):
        self._ck(self.L.ecmgpu_sync(self.h))

    # -- multi-GPU strips
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = lib().ecmgpu_comm_unique_id(buf)
        if rc != 0:
            raise EcmGpuError(f"ecmgpu_comm_unique_id failed ({rc}): {lib().ecmgpu_last_error(None).decode()}")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self._ck(self.L.ecmgpu_comm_init(self.h, buf, int(rank), int(n_ranks)))

    def comm_init_local(self, rank: int, n_ranks: int, left: "GpuSim | None", right: "GpuSim | None"):
        self._ck(self.L.ecmgpu_comm_init_local(self.h, int(rank), int(n_ranks), left.h if left else None,
                                               right.h if right else None))

    def comm_p2p_export(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._ck(self.L.ecmgpu_comm_p2p_export(self.h, buf))
        return bytes(buf)

    def comm_p2p_connect(self, left: bytes | None, right: bytes | None):
        lb = (C.c_uint8 * 128).from_buffer_copy(left) if left is not None else None
        rb = (C.c_uint8 * 128).from_buffer_copy(right) if right is not None else None
        self._ck(self.L.ecmgpu_comm_p2p_connect(self.h, lb, rb))

    def comm_set_strips(self, bounds, halo: float):
        b = np.ascontiguousarray(bounds, np.float32)
        self._ck(self.L.ecmgpu_comm_set_strips(self.h, _p(b, f32p), float(halo)))

    def poll_events(self):
        """(replan_slots, destroyed_slots) since the last poll, ascending."""
        rp = np.zeros(self.max_agents, np.int32)
        ds = np.zeros(self.max_agents, np.int32)
        nr, nd = C.c_int(0), C.c_int(0)
        self._ck(self.L.ecmgpu_poll_events(self.h, _p(rp, i32p), len(rp), C.byref(nr), _p(ds, i32p), len(ds), C.byref(nd)))
        return rp[: nr.value].copy(), ds[: nd.value].copy()

    def step(self, n: int = 1):
        """n ticks, then the events of those ticks: same return shape as OracleSim.step."""
        self.update(n)
        return self.poll_events()

    # -- transfers
    def read(self, which, first=0, count=None):
        n = self.max_agents - first if count is None else int(count)
        dt, k = _DTYPES[which]
        out = np.zeros((n, k) if k > 1 else (n,), dt)
        self._ck(self.L.ecmgpu_read(self.h, which, out.ctypes.data_as(C.c_void_p), int(first), n))
        return out

    def write(self, which, arr, first=0):
        dt, k = _DTYPES[which]
        a = np.ascontiguousarray(arr, dt)
        n = a.shape[0]
        self._ck(self.L.ecmgpu_write(self.h, which, a.ctypes.data_as(C.c_void_p), int(first), n))

    def read_async(self, which, pinned: PinnedArray, first, count):
        self._ck(self.L.ecmgpu_read_async(self.h, which, C.c_void_p(pinned.ptr), int(first), int(count)))

    def write_async(self, which, pinned: PinnedArray, first, count):
        self._ck(self.L.ecmgpu_write_async(self.h, which, C.c_void_p(pinned.ptr), int(first), int(count)))

    def state(self, count=None):
        n = self.max_agents if count is None else int(count)
        out = {"pos": self.read(POS, 0, n), "vel": self.read(VEL, 0, n), "prefvel": self.read(PREFVEL, 0, n),
               "attraction": self.read(ATTRACTION, 0, n), "force": self.read(FORCE, 0, n), "active": self.read(ACTIVE, 0, n)}
        return out

    def set_kinematics(self, slot, pos, vel):
        self.write(POS, np.asarray([pos], np.float32), slot)
        self.write(VEL, np.asarray([vel], np.float32), slot)

    def set_attraction(self, slot, p):
        self.write(ATTRACTION, np.asarray([p], np.float32), slot)

    # -- queries
    def query_cells(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.int32)
        self._ck(self.L.ecmgpu_locate(self.h, len(xy), _p(xy, f32p), _p(out, i32p)))
        return out

    def retract(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        n = len(xy)
        ok = np.zeros(n, np.uint8)
        out = np.zeros((n, 2), np.float32)
        edge = np.zeros(n, np.int32)
        self._ck(self.L.ecmgpu_retract(self.h, n, _p(xy, f32p), _p(ok, u8p), _p(out, f32p), _p(edge, i32p)))
        return ok, out, edge

    def query_neighbors(self, count=None):
        n = self.max_agents if count is None else int(count)
        ids = np.full((n, 5), -1, np.int32)
        cnt = np.full(n, -1, np.int32)
        self._ck(self.L.ecmgpu_find_neighbors(self.h, n, _p(ids, i32p), _p(cnt, i32p)))
        return ids, cnt

    def plan_paths(self, start, goal, clearance, points_per_path: int = 48):
        """Batched ECMPathPlanner::FindPath on the device.  Same return shape as host.plan_paths:
        (path_off[n+1], path_xy[total, 2], n_ok), paths in query order; a failed query contributes zero points."""
        start = np.ascontiguousarray(start, np.float32).reshape(-1, 2)
        goal = np.ascontiguousarray(goal, np.float32).reshape(-1, 2)
        n = len(start)
        cl = np.ascontiguousarray(np.broadcast_to(np.asarray(clearance, np.float32), (n,)))
        off, ln, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        cap = max(64, int(points_per_path) * n)
        while True:
            pool = np.zeros((cap, 2), np.float32)
            used = C.c_int(0)
            self._ck(self.L.ecmgpu_plan_paths(self.h, n, _p(start, f32p), _p(goal, f32p), _p(cl, f32p), _p(off, i32p), _p(ln, i32p),
                                              _p(st, u8p), _p(pool, f32p), cap, C.byref(used)))
            if used.value <= cap:
                break
            cap = used.value + 64  # the pool was too small: now its size is known
        if (st == 2).any():
            raise EcmGpuError(f"ecmgpu_plan_paths: {int((st == 2).sum())} queries exceeded a planner capacity")
        out_off = np.zeros(n + 1, np.int32)
        np.cumsum(ln, out=out_off[1:])
        total = int(out_off[-1])
        idx = np.repeat(off.astype(np.int64), ln) + (np.arange(total) - np.repeat(out_off[:-1].astype(np.int64), ln))
        return out_off, pool[idx], int((ln > 0).sum())

    def plan_info(self):
        """(queries the last plan_paths kept in flight, device time of its kernels in ms, queries that needed the second
        pass) - ecmgpu_plan_info."""
        wk, ms, sp = C.c_int(0), C.c_float(0.0), C.c_int(0)
        self._ck(self.L.ecmgpu_plan_info(self.h, C.byref(wk), C.byref(ms), C.byref(sp)))
        return wk.value, ms.value, sp.value

    def valid_spawn_locations(self, xy, clearance):
        """Simulator::ValidSpawnLocation for a batch of points on the current positions (uint8 flags)."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        cl = np.ascontiguousarray(np.broadcast_to(np.asarray(clearance, np.float32), (len(xy),)))
        out = np.zeros(len(xy), np.uint8)
        self._ck(self.L.ecmgpu_valid_spawn_locations(self.h, len(xy), _p(xy, f32p), _p(cl, f32p), _p(out, u8p)))
        return out

    def draw_spawns(self, spawn_boxes, goal_boxes, clearance, seed: int, counter: int, max_attempts: int = 10):
        """Spawn draws with the counter-based generator on the device (ecmgpu_draw_spawns): (start[n,2], goal[n,2], ok[n])."""
        sb = np.ascontiguousarray(spawn_boxes, np.float32).reshape(-1, 4)
        gb = np.ascontiguousarray(goal_boxes, np.float32).reshape(-1, 4)
        cl = np.ascontiguousarray(np.broadcast_to(np.asarray(clearance, np.float32), (len(sb),)))
        start, goal, ok = np.zeros((len(sb), 2), np.float32), np.zeros((len(sb), 2), np.float32), np.zeros(len(sb), np.uint8)
        self._ck(self.L.ecmgpu_draw_spawns(self.h, len(sb), _p(sb, f32p), _p(gb, f32p), _p(cl, f32p), int(seed), int(counter), int(max_attempts),
                                           _p(start, f32p), _p(goal, f32p), _p(ok, u8p)))
        return start, goal, ok

    def set_neighbor_mode(self, mode: int):
        """NEIGHBORS_EXACT (default) or NEIGHBORS_KDTREE: the reference's own KD-tree lists (parity mode)."""
        self._ck(self.L.ecmgpu_set_neighbor_mode(self.h, int(mode)))

    def query_obstacles(self, slot, cap=256):
        out = np.zeros(cap, np.int32)
        n = C.c_int(0)
        self._ck(self.L.ecmgpu_find_obstacles(self.h, int(slot), _p(out, i32p), cap, C.byref(n)))
        return out[: min(n.value, cap)].copy()

    # -- introspection
    def stats(self) -> dict:
        st = Stats()
        self._ck(self.L.ecmgpu_get_stats(self.h, C.byref(st)))
        return {name: getattr(st, name) for name, _ in Stats._fields_}

    def set_profiling(self, on: bool):
        self._ck(self.L.ecmgpu_set_profiling(self.h, 1 if on else 0))

    def mark(self, which: int):
        self._ck(self.L.ecmgpu_mark(self.h, int(which)))

    def elapsed_ms(self, a: int, b: int) -> float:
        out = C.c_float(0)
        self._ck(self.L.ecmgpu_mark_elapsed_ms(self.h, int(a), int(b), C.byref(out)))
        return float(out.value)

    def last_tick_ms(self):
        out = (C.c_float * 4)()
        self._ck(self.L.ecmgpu_last_tick_ms(self.h, out))
        return {"tick": out[0], "grid": out[1], "attract": out[2], "orca": out[3]}

    def last_tick_phases(self):
        """Like last_tick_ms with k_orca and k_fallback apart (ecmgpu_last_tick_phases)."""
        out = (C.c_float * 5)()
        self._ck(self.L.ecmgpu_last_tick_phases(self.h, out))
        return {"tick": out[0], "grid": out[1], "attract": out[2], "orca": out[3], "fallback": out[4]}
