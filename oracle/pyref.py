"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

ctypes binding of oracle/_ref/libecmref.so ("ref-kdtree": the unmodified reference hot path) and
oracle/_ref/libecmref_exact.so ("exact-knn": same, KDTree.cpp swapped for oracle/kdtree_exact.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Each process can load both libraries side by side (RTLD_LOCAL keeps their symbols apart).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def available(mode: str = "ref-kdtree") -> bool:
    return os.path.exists(_path(mode))


def _path(mode):
    name = {"ref-kdtree": "libecmref.so", "exact-knn": "libecmref_exact.so"}[mode]
    return os.path.join(_HERE, "_ref", name)


def _load(mode):
    if mode not in _LIBS:
        L = C.CDLL(_path(mode), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.ecmref_create.restype = C.c_void_p
        L.ecmref_create.argtypes = [f32p, C.c_int, f32p, f32p, i32p, C.c_int, i32p, f32p, i32p, C.c_int, i32p, f32p, C.c_int, C.c_float]
        L.ecmref_destroy.argtypes = [C.c_void_p]
        L.ecmref_get_obstacles.argtypes = [C.c_void_p, f32p, i32p, i32p, u8p]
        L.ecmref_plan_path.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_int]
        L.ecmref_bulk_load.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, f32p, i32p, f32p, i32p]
        L.ecmref_spawn.argtypes = [C.c_void_p] + [C.c_float] * 6
        L.ecmref_set_kinematics.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
        L.ecmref_set_attraction.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 2
        L.ecmref_path_len.argtypes = [C.c_void_p, C.c_int]
        L.ecmref_destroy_agent.argtypes = [C.c_void_p, C.c_int]
        L.ecmref_get_path.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
        L.ecmref_step.argtypes = [C.c_void_p, C.c_int]
        L.ecmref_num_agents.argtypes = [C.c_void_p]
        L.ecmref_last_index.argtypes = [C.c_void_p]
        L.ecmref_get_state.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, f32p, f32p, u8p]
        L.ecmref_query_cells.argtypes = [C.c_void_p, C.c_int, f32p, i32p]
        L.ecmref_retract.argtypes = [C.c_void_p, C.c_int, f32p, u8p, f32p, i32p]
        L.ecmref_query_neighbors.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.ecmref_query_obstacles.argtypes = [C.c_void_p, C.c_int, i32p, C.c_int]
        L.ecmref_orca_velocity.argtypes = [C.c_void_p, C.c_int, f32p]
        L.ecmref_add_spawn_area.argtypes = [C.c_void_p] + [C.c_float] * 6
        L.ecmref_add_goal_area.argtypes = [C.c_void_p] + [C.c_float] * 4
        L.ecmref_connect_areas.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.ecmref_add_obstacle_area.argtypes = [C.c_void_p] + [C.c_float] * 4
        L.ecmref_srand.argtypes = [C.c_uint]
        L.ecmref_valid_spawn_location.argtypes = [C.c_void_p] + [C.c_float] * 3
        _LIBS[mode] = L
    return _LIBS[mode]


class RefSim:
    """The reference Simulator (ECMAgentSimulator/Simulator.h:59-188) on a flat world."""

    def __init__(self, world, max_agents: int, step: float, mode: str = "ref-kdtree"):
        self.L = _load(mode)
        self.mode = mode
        self.max_agents = int(max_agents)
        w = world
        self._keep = [np.ascontiguousarray(a) for a in (w.bbox, w.vert_xy, w.vert_clear, w.vert_he, w.edge_v, w.edge_cl,
                                                        w.he_next, w.obst_first, w.obst_xy)]
        k = self._keep
        self.h = self.L.ecmref_create(_p(k[0], f32p), w.n_vertices, _p(k[1], f32p), _p(k[2], f32p), _p(k[3], i32p),
                                      w.n_edges, _p(k[4], i32p), _p(k[5], f32p), _p(k[6], i32p), w.n_obstacles,
                                      _p(k[7], i32p), _p(k[8], f32p), self.max_agents, float(step))

    def close(self):
        if self.h:
            self.L.ecmref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def obstacles(self):
        n = self.L.ecmref_get_obstacles(self.h, None, None, None, None)
        xy = np.zeros((n, 2), np.float32)
        nx = np.zeros(n, np.int32)
        pv = np.zeros(n, np.int32)
        cv = np.zeros(n, np.uint8)
        self.L.ecmref_get_obstacles(self.h, _p(xy, f32p), _p(nx, i32p), _p(pv, i32p), _p(cv, u8p))
        return xy, nx, pv, cv

    def plan_path(self, start, goal, clearance: float, cap: int = 4096):
        out = np.zeros((cap, 2), np.float32)
        n = self.L.ecmref_plan_path(self.h, float(start[0]), float(start[1]), float(goal[0]), float(goal[1]),
                                    float(clearance), _p(out, f32p), cap)
        return None if n < 0 else out[:n].copy()

    def bulk_load(self, pos, goal, radius, speed, path_off=None, path_xy=None):
        pos = np.ascontiguousarray(pos, np.float32)
        goal = np.ascontiguousarray(goal if goal is not None else pos, np.float32)
        radius = np.ascontiguousarray(radius, np.float32)
        speed = np.ascontiguousarray(speed, np.float32)
        n = len(pos)
        slots = np.full(n, -1, np.int32)
        if path_off is not None:
            path_off = np.ascontiguousarray(path_off, np.int32)
            path_xy = np.ascontiguousarray(path_xy, np.float32)
        self.L.ecmref_bulk_load(self.h, n, _p(pos, f32p), _p(goal, f32p), _p(radius, f32p), _p(speed, f32p),
                                _p(path_off, i32p), _p(path_xy, f32p), _p(slots, i32p))
        return slots

    def spawn(self, start, goal, clearance, speed) -> int:
        return self.L.ecmref_spawn(self.h, float(start[0]), float(start[1]), float(goal[0]), float(goal[1]),
                                   float(clearance), float(speed))

    def add_spawn_area(self, pos, half, clearance, speed) -> int:
        return self.L.ecmref_add_spawn_area(self.h, float(pos[0]), float(pos[1]), float(half[0]), float(half[1]), float(clearance), float(speed))

    def add_goal_area(self, pos, half) -> int:
        return self.L.ecmref_add_goal_area(self.h, float(pos[0]), float(pos[1]), float(half[0]), float(half[1]))

    def connect_areas(self, spawn_id, goal_id, rate):
        self.L.ecmref_connect_areas(self.h, int(spawn_id), int(goal_id), float(rate))

    def add_obstacle_area(self, pos, half) -> int:
        return self.L.ecmref_add_obstacle_area(self.h, float(pos[0]), float(pos[1]), float(half[0]), float(half[1]))

    def srand(self, seed: int):
        self.L.ecmref_srand(int(seed))

    def valid_spawn_location(self, p, clearance) -> bool:
        return bool(self.L.ecmref_valid_spawn_location(self.h, float(p[0]), float(p[1]), float(clearance)))

    def set_kinematics(self, slot, pos, vel):
        self.L.ecmref_set_kinematics(self.h, int(slot), float(pos[0]), float(pos[1]), float(vel[0]), float(vel[1]))

    def set_attraction(self, slot, p):
        self.L.ecmref_set_attraction(self.h, int(slot), float(p[0]), float(p[1]))

    def destroy_agent(self, slot):
        self.L.ecmref_destroy_agent(self.h, int(slot))

    def path_len(self, slot) -> int:
        return self.L.ecmref_path_len(self.h, int(slot))

    def path(self, slot):
        n = self.path_len(slot)
        xy = np.zeros((max(n, 1), 2), np.float32)
        self.L.ecmref_get_path(self.h, int(slot), _p(xy, f32p), n)
        return xy[:n]

    def paths(self, count):
        """(path_off[count+1], path_xy[total,2]) of slots [0,count)."""
        lens = np.array([self.L.ecmref_path_len(self.h, i) for i in range(count)], np.int32)
        off = np.zeros(count + 1, np.int32)
        np.cumsum(lens, out=off[1:])
        xy = np.zeros((max(int(off[-1]), 1), 2), np.float32)
        for i in range(count):
            if lens[i] > 0:
                seg = xy[off[i]:off[i + 1]]
                self.L.ecmref_get_path(self.h, i, _p(seg, f32p), int(lens[i]))
        return off, xy[: int(off[-1])]

    def step(self, n: int = 1):
        self.L.ecmref_step(self.h, int(n))

    @property
    def num_agents(self):
        return self.L.ecmref_num_agents(self.h)

    @property
    def last_index(self):
        return self.L.ecmref_last_index(self.h)

    def state(self, count=None):
        n = self.max_agents if count is None else int(count)
        out = {k: np.zeros((n, 2), np.float32) for k in ("pos", "vel", "prefvel", "attraction", "force")}
        act = np.zeros(n, np.uint8)
        self.L.ecmref_get_state(self.h, n, _p(out["pos"], f32p), _p(out["vel"], f32p), _p(out["prefvel"], f32p),
                                _p(out["attraction"], f32p), _p(out["force"], f32p), _p(act, u8p))
        out["active"] = act
        return out

    def query_cells(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.int32)
        self.L.ecmref_query_cells(self.h, len(xy), _p(xy, f32p), _p(out, i32p))
        return out

    def retract(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        n = len(xy)
        ok = np.zeros(n, np.uint8)
        out = np.zeros((n, 2), np.float32)
        edge = np.zeros(n, np.int32)
        self.L.ecmref_retract(self.h, n, _p(xy, f32p), _p(ok, u8p), _p(out, f32p), _p(edge, i32p))
        return ok, out, edge

    def query_neighbors(self, count=None):
        n = self.max_agents if count is None else int(count)
        ids = np.full((n, 5), -1, np.int32)
        cnt = np.full(n, -1, np.int32)
        self.L.ecmref_query_neighbors(self.h, n, _p(ids, i32p), _p(cnt, i32p))
        return ids, cnt

    def query_obstacles(self, slot, cap=256):
        out = np.zeros(cap, np.int32)
        n = self.L.ecmref_query_obstacles(self.h, int(slot), _p(out, i32p), cap)
        return out[: min(n, cap)].copy()

    def orca_velocity(self, slot):
        out = np.zeros(2, np.float32)
        self.L.ecmref_orca_velocity(self.h, int(slot), _p(out, f32p))
        return out
