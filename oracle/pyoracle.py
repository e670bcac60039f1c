"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

ctypes binding of oracle/libecmoracle.so (the plain-C restatement, oracle/ecm_oracle.c).  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)

KNN_REF_KDTREE = 0
KNN_EXACT = 1
_MODES = {"ref-kdtree": KNN_REF_KDTREE, "exact-knn": KNN_EXACT}


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build(force: bool = False) -> str:
    """Compiles the C restatement (gcc, seconds).  Building the checker is not using it."""
    so = os.path.join(_HERE, "libecmoracle.so")
    src = [os.path.join(_HERE, f) for f in ("ecm_oracle.c", "ecm_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def _load():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build(), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.eo_create.restype = C.c_void_p
        L.eo_create.argtypes = [C.c_int, f32p, f32p, C.c_int, i32p, f32p, C.c_int, f32p, i32p, i32p, u8p, C.c_int, C.c_float, C.c_int]
        L.eo_destroy.argtypes = [C.c_void_p]
        L.eo_bulk_load.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, i32p, f32p, i32p]
        L.eo_set_path.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
        L.eo_set_kinematics.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
        L.eo_set_attraction.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 2
        L.eo_destroy_agent.argtypes = [C.c_void_p, C.c_int]
        L.eo_step.argtypes = [C.c_void_p]
        for name in ("eo_num_replans", "eo_num_destroyed", "eo_num_agents", "eo_last_index"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.eo_replans.restype = i32p
        L.eo_replans.argtypes = [C.c_void_p]
        L.eo_destroyed.restype = i32p
        L.eo_destroyed.argtypes = [C.c_void_p]
        L.eo_get_state.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, f32p, f32p, u8p]
        L.eo_query_cells.argtypes = [C.c_void_p, C.c_int, f32p, i32p]
        L.eo_retract.argtypes = [C.c_void_p, C.c_int, f32p, u8p, f32p, i32p]
        L.eo_query_neighbors.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.eo_query_obstacles.argtypes = [C.c_void_p, C.c_int, i32p, C.c_int]
        L.eo_orca_velocity.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p, f32p]
        L.eo_counters.restype = C.POINTER(C.c_longlong)
        L.eo_counters.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class OracleSim:
    """Same surface as oracle.pyref.RefSim, backed by the C restatement."""

    def __init__(self, world, max_agents: int, step: float, mode: str = "exact-knn"):
        self.L = _load()
        self.mode = mode
        self.max_agents = int(max_agents)
        w = world
        a = [np.ascontiguousarray(x) for x in (w.vert_xy, w.vert_clear, w.edge_v, w.edge_cl, w.obst_xy, w.obst_next,
                                               w.obst_prev, w.obst_convex)]
        self.h = self.L.eo_create(w.n_vertices, _p(a[0], f32p), _p(a[1], f32p), w.n_edges, _p(a[2], i32p), _p(a[3], f32p),
                                  w.n_obst_vertices, _p(a[4], f32p), _p(a[5], i32p), _p(a[6], i32p), _p(a[7], u8p),
                                  self.max_agents, float(step), _MODES[mode])

    def close(self):
        if self.h:
            self.L.eo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bulk_load(self, pos, radius, speed, path_off, path_xy):
        pos = np.ascontiguousarray(pos, np.float32)
        radius = np.ascontiguousarray(radius, np.float32)
        speed = np.ascontiguousarray(speed, np.float32)
        path_off = np.ascontiguousarray(path_off, np.int32)
        path_xy = np.ascontiguousarray(path_xy, np.float32)
        slots = np.full(len(pos), -1, np.int32)
        self.L.eo_bulk_load(self.h, len(pos), _p(pos, f32p), _p(radius, f32p), _p(speed, f32p), _p(path_off, i32p),
                            _p(path_xy, f32p), _p(slots, i32p))
        return slots

    def set_path(self, slot, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        self.L.eo_set_path(self.h, int(slot), _p(xy, f32p), len(xy))

    def destroy_agent(self, slot):
        self.L.eo_destroy_agent(self.h, int(slot))

    def set_kinematics(self, slot, pos, vel):
        self.L.eo_set_kinematics(self.h, int(slot), float(pos[0]), float(pos[1]), float(vel[0]), float(vel[1]))

    def set_attraction(self, slot, p):
        self.L.eo_set_attraction(self.h, int(slot), float(p[0]), float(p[1]))

    def step(self, n: int = 1):
        """Steps n ticks; returns (replans, destroyed) slot arrays of the LAST tick."""
        for _ in range(int(n)):
            self.L.eo_step(self.h)
        nr, nd = self.L.eo_num_replans(self.h), self.L.eo_num_destroyed(self.h)
        rp = np.ctypeslib.as_array(self.L.eo_replans(self.h), shape=(max(nr, 1),))[:nr].copy()
        ds = np.ctypeslib.as_array(self.L.eo_destroyed(self.h), shape=(max(nd, 1),))[:nd].copy()
        return rp, ds

    @property
    def num_agents(self):
        return self.L.eo_num_agents(self.h)

    @property
    def last_index(self):
        return self.L.eo_last_index(self.h)

    def counters(self):
        c = self.L.eo_counters(self.h)
        return {"lp3d": c[0], "lp_calls": c[1], "max_obstacle_neighbours": c[2], "irm_failures": c[3],
                "agent_updates": c[5], "concave_segments": c[6], "oblique_segments": c[7]}

    def state(self, count=None):
        n = self.max_agents if count is None else int(count)
        out = {k: np.zeros((n, 2), np.float32) for k in ("pos", "vel", "prefvel", "attraction", "force")}
        act = np.zeros(n, np.uint8)
        self.L.eo_get_state(self.h, n, _p(out["pos"], f32p), _p(out["vel"], f32p), _p(out["prefvel"], f32p),
                            _p(out["attraction"], f32p), _p(out["force"], f32p), _p(act, u8p))
        out["active"] = act
        return out

    def query_cells(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.int32)
        self.L.eo_query_cells(self.h, len(xy), _p(xy, f32p), _p(out, i32p))
        return out

    def retract(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        n = len(xy)
        ok = np.zeros(n, np.uint8)
        out = np.zeros((n, 2), np.float32)
        edge = np.zeros(n, np.int32)
        self.L.eo_retract(self.h, n, _p(xy, f32p), _p(ok, u8p), _p(out, f32p), _p(edge, i32p))
        return ok, out, edge

    def query_neighbors(self, count=None):
        n = self.max_agents if count is None else int(count)
        ids = np.full((n, 5), -1, np.int32)
        cnt = np.full(n, -1, np.int32)
        self.L.eo_query_neighbors(self.h, n, _p(ids, i32p), _p(cnt, i32p))
        return ids, cnt

    def query_obstacles(self, slot, cap=256):
        out = np.zeros(cap, np.int32)
        n = self.L.eo_query_obstacles(self.h, int(slot), _p(out, i32p), cap)
        return out[: min(n, cap)].copy()

    def orca_velocity(self, slot, neighbors):
        nb = np.ascontiguousarray(neighbors, np.int32)
        out = np.zeros(2, np.float32)
        self.L.eo_orca_velocity(self.h, int(slot), len(nb), _p(nb, i32p), _p(out, f32p))
        return out
