"""ctypes binding of libecmsim.so: the C++17 drop-in `ECM::Simulation::Simulator`
(csrc/dropin/Simulator.h) driven the way the reference's Application drives its Simulator."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import host

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libecmsim.so")
_lib = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ecmsim_last_error.restype = C.c_char_p
        L.ecmsim_create.restype = vp
        L.ecmsim_create.argtypes = [vp, C.c_int, C.c_float, C.c_int]
        L.ecmsim_destroy.argtypes = [vp]
        L.ecmsim_spawn_agent.argtypes = [vp] + [C.c_float] * 6
        L.ecmsim_destroy_agent.argtypes = [vp, C.c_int]
        L.ecmsim_update.argtypes = [vp, C.c_float]
        L.ecmsim_reset.argtypes = [vp]
        L.ecmsim_update_path.argtypes = [vp, C.c_int] + [C.c_float] * 4
        L.ecmsim_add_position.argtypes = [vp, C.c_int, C.c_float, C.c_float]
        L.ecmsim_num_agents.argtypes = [vp]
        L.ecmsim_last_index.argtypes = [vp]
        for name in ("positions", "velocities", "preferred_velocities", "attraction_points", "clearances"):
            f = getattr(L, "ecmsim_" + name)
            f.restype = f32p
            f.argtypes = [vp]
        L.ecmsim_active_flags.restype = u8p
        L.ecmsim_active_flags.argtypes = [vp]
        L.ecmsim_path.argtypes = [vp, C.c_int, f32p, C.c_int]
        L.ecmsim_valid_spawn_location.argtypes = [vp, C.c_float, C.c_float, C.c_float]
        L.ecmsim_find_neighbors.argtypes = [vp, C.c_int, i32p]
        L.ecmsim_find_obstacles.argtypes = [vp, C.c_int, C.c_float, i32p, C.c_int]
        L.ecmsim_add_spawn_area.argtypes = [vp] + [C.c_float] * 6
        L.ecmsim_set_spawn_mode.argtypes = [vp, C.c_int, C.c_uint64]
        L.ecmsim_spawn_checks.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.ecmsim_add_goal_area.argtypes = [vp] + [C.c_float] * 4
        L.ecmsim_connect_areas.argtypes = [vp, C.c_int, C.c_int, C.c_float]
        L.ecmsim_add_obstacle_area.argtypes = [vp] + [C.c_float] * 4 + [C.c_int]
        L.ecmsim_num_obstacle_vertices.argtypes = [vp]
        L.ecmsim_find_neighbors_via.argtypes = [vp, C.c_int, C.c_int, i32p]
        L.ecmsim_set_neighbor_mode.argtypes = [vp, C.c_int]
        # the same library also carries the ecmhost_* entry points (one FlatWorld layout)
        L.ecmhost_world_from_arrays.restype = vp
        L.ecmhost_world_from_arrays.argtypes = [C.POINTER(host._WorldView)]
        L.ecmhost_world_free.argtypes = [vp]
        _lib = L
    return _lib


class Simulator:
    """ECM::Simulation::Simulator (drop-in) on a flat world."""

    def __init__(self, world, max_agents: int, step: float, device: int = 0):
        self.L = lib()
        self.max_agents = int(max_agents)
        self._wh = _world_in(self.L, world)  # the world is copied into libecmsim's own FlatWorld
        self.h = self.L.ecmsim_create(self._wh[0], self.max_agents, float(step), int(device))
        if not self.h:
            raise RuntimeError("ecmsim_create failed: " + self.L.ecmsim_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.ecmsim_destroy(self.h)
            self.h = None
            self.L.ecmhost_world_free(self._wh[0])

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc == -2:
            raise RuntimeError(self.L.ecmsim_last_error().decode())
        return rc

    def spawn_agent(self, start, goal, clearance, speed) -> int:
        return self._ck(self.L.ecmsim_spawn_agent(self.h, float(start[0]), float(start[1]), float(goal[0]), float(goal[1]),
                                                  float(clearance), float(speed)))

    def destroy_agent(self, idx):
        self._ck(self.L.ecmsim_destroy_agent(self.h, int(idx)))

    def update(self, dt: float = 0.0):
        self._ck(self.L.ecmsim_update(self.h, float(dt)))

    def reset(self):
        self._ck(self.L.ecmsim_reset(self.h))

    @property
    def num_agents(self):
        return self.L.ecmsim_num_agents(self.h)

    @property
    def last_index(self):
        return self.L.ecmsim_last_index(self.h)

    def _arr(self, name, k):
        p = getattr(self.L, "ecmsim_" + name)(self.h)
        return np.ctypeslib.as_array(p, shape=(self.max_agents * k,)).reshape(self.max_agents, k).copy() if k > 1 else \
            np.ctypeslib.as_array(p, shape=(self.max_agents,)).copy()

    def state(self, count=None):
        n = self.max_agents if count is None else int(count)
        act = np.ctypeslib.as_array(self.L.ecmsim_active_flags(self.h), shape=(self.max_agents,)).copy()
        return {"pos": self._arr("positions", 2)[:n], "vel": self._arr("velocities", 2)[:n],
                "prefvel": self._arr("preferred_velocities", 2)[:n], "attraction": self._arr("attraction_points", 2)[:n],
                "active": act[:n]}

    def path(self, slot, cap=4096):
        out = np.zeros((cap, 2), np.float32)
        n = self.L.ecmsim_path(self.h, int(slot), out.ctypes.data_as(f32p), cap)
        return out[:n].copy()

    def valid_spawn_location(self, p, clearance) -> bool:
        return bool(self.L.ecmsim_valid_spawn_location(self.h, float(p[0]), float(p[1]), float(clearance)))

    def find_neighbors(self, agent):
        out = np.full(5, -1, np.int32)
        n = self._ck(self.L.ecmsim_find_neighbors(self.h, int(agent), out.ctypes.data_as(i32p)))
        return out, n

    def add_spawn_area(self, pos, half, clearance, speed) -> int:
        return self.L.ecmsim_add_spawn_area(self.h, float(pos[0]), float(pos[1]), float(half[0]), float(half[1]), float(clearance), float(speed))

    def add_goal_area(self, pos, half) -> int:
        return self.L.ecmsim_add_goal_area(self.h, float(pos[0]), float(pos[1]), float(half[0]), float(half[1]))

    SPAWN_RAND_BATCHED, SPAWN_RAND_SEQUENTIAL, SPAWN_DEVICE_COUNTER = 0, 1, 2

    def set_spawn_mode(self, mode: int, seed: int = 0):
        """Simulator::SetSpawnMode: rand() in the reference's order with GPU validity batches (default), the reference's
        own loop, or the counter-based generator on the device."""
        self.L.ecmsim_set_spawn_mode(self.h, int(mode), int(seed))

    def spawn_checks(self):
        """(validity tests answered by the GPU, by the host scan) since construction."""
        out = (C.c_longlong * 2)()
        self.L.ecmsim_spawn_checks(self.h, out)
        return int(out[0]), int(out[1])

    def connect_areas(self, spawn_id, goal_id, rate):
        self.L.ecmsim_connect_areas(self.h, int(spawn_id), int(goal_id), float(rate))

    def find_obstacles(self, agent, range_squared: float, cap: int = 256):
        """Simulator::FindNearestObstacles: flat obstacle-vertex indices in (obstacle, vertex) order."""
        out = np.full(cap, -1, np.int32)
        n = self.L.ecmsim_find_obstacles(self.h, int(agent), float(range_squared), out.ctypes.data_as(i32p), cap)
        return out[: min(n, cap)].copy()

    def add_obstacle_area(self, pos, half, update_ecm: bool = False) -> int:
        return self._ck(self.L.ecmsim_add_obstacle_area(self.h, float(pos[0]), float(pos[1]), float(half[0]), float(half[1]), int(update_ecm)))

    def num_obstacle_vertices(self) -> int:
        return int(self.L.ecmsim_num_obstacle_vertices(self.h))

    def set_neighbor_mode(self, mode: int):
        """gpu.NEIGHBORS_KDTREE: the tick uses the reference's own KD-tree lists (parity runs against the unmodified reference)."""
        if self.L.ecmsim_set_neighbor_mode(self.h, int(mode)) != 0:
            raise RuntimeError(self.L.ecmsim_last_error().decode())

    def find_neighbors_via(self, agent, route: str):
        """route 'kdtree': GetKDTree()->KNearestAgents; 'deprecated': FindNNearestNeighborsDeprecated."""
        out = np.full(5, -1, np.int32)
        n = self._ck(self.L.ecmsim_find_neighbors_via(self.h, int(agent), 0 if route == "kdtree" else 1, out.ctypes.data_as(i32p)))
        return out, n


def _world_in(L, w):
    keep = [np.ascontiguousarray(a) for a in (w.vert_xy, w.vert_clear, w.vert_he, w.edge_v, w.edge_cl, w.he_next, w.obst_xy,
                                              w.obst_next, w.obst_prev, w.obst_convex, w.obst_first)]
    v = host._WorldView()
    for i in range(4):
        v.bbox[i] = float(w.bbox[i])
    v.n_vertices, v.n_edges, v.n_obst_vertices, v.n_obstacles = w.n_vertices, w.n_edges, w.n_obst_vertices, w.n_obstacles
    v.vert_xy, v.vert_clear, v.vert_he = host.fptr(keep[0]), host.fptr(keep[1]), host.iptr(keep[2])
    v.edge_v, v.edge_cl, v.he_next = host.iptr(keep[3]), host.fptr(keep[4]), host.iptr(keep[5])
    v.obst_xy, v.obst_next, v.obst_prev, v.obst_convex, v.obst_first = (host.fptr(keep[6]), host.iptr(keep[7]), host.iptr(keep[8]),
                                                                        host.u8ptr(keep[9]), host.iptr(keep[10]))
    h = L.ecmhost_world_from_arrays(C.byref(v))
    return (h, keep)
