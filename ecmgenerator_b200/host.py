"""ctypes binding of the host-side C ABI (include/ecm_b200_host.h, libecmhost.so).

Host-side = everything that is *input preparation* for the GPU tick: ECM construction for lattice
worlds, flattening to structure-of-arrays, global path planning.  The reference does these in
ECMGenerator::GenerateECM (/root/reference/ECMGenerator/ECMGenerator.cpp:235-256),
Environment (/root/reference/ECMGenerator/Environment.cpp:188-229) and ECMPathPlanner::FindPath
(/root/reference/ECMGenerator/ECMPathPlanner.cpp:22-136).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libecmhost.so")
_lib = None

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)


class _WorldView(C.Structure):
    _fields_ = [
        ("bbox", C.c_float * 4),
        ("n_vertices", C.c_int),
        ("n_edges", C.c_int),
        ("n_obst_vertices", C.c_int),
        ("n_obstacles", C.c_int),
        ("vert_xy", c_float_p),
        ("vert_clear", c_float_p),
        ("vert_he", c_int_p),
        ("edge_v", c_int_p),
        ("edge_cl", c_float_p),
        ("he_next", c_int_p),
        ("obst_xy", c_float_p),
        ("obst_next", c_int_p),
        ("obst_prev", c_int_p),
        ("obst_convex", c_u8_p),
        ("obst_first", c_int_p),
    ]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_LIB_PATH)
        L.ecmhost_lattice_world.restype = C.c_void_p
        L.ecmhost_lattice_world.argtypes = [C.c_int, c_float_p, C.c_int, c_float_p, C.c_float, C.c_float, C.c_float]
        L.ecmhost_polygon_world.restype = C.c_void_p
        L.ecmhost_polygon_world.argtypes = [c_float_p, C.c_int, c_int_p, c_float_p, C.c_char_p, C.c_int]
        L.ecmhost_world_from_arrays.restype = C.c_void_p
        L.ecmhost_world_from_arrays.argtypes = [C.POINTER(_WorldView)]
        L.ecmhost_world_free.argtypes = [C.c_void_p]
        L.ecmhost_world_get_view.argtypes = [C.c_void_p, C.POINTER(_WorldView)]
        L.ecmhost_plan_paths.restype = C.c_void_p
        L.ecmhost_plan_paths.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int]
        L.ecmhost_paths_count.argtypes = [C.c_void_p]
        L.ecmhost_paths_succeeded.argtypes = [C.c_void_p]
        L.ecmhost_paths_offsets.restype = c_int_p
        L.ecmhost_paths_offsets.argtypes = [C.c_void_p]
        L.ecmhost_paths_xy.restype = c_float_p
        L.ecmhost_paths_xy.argtypes = [C.c_void_p]
        L.ecmhost_paths_free.argtypes = [C.c_void_p]
        L.ecmhost_find_cells.argtypes = [C.c_void_p, C.c_int, c_float_p, c_int_p]
        _lib = L
    return _lib


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_float_p)


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_int_p)


def u8ptr(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(c_u8_p)


@dataclass
class World:
    """Flat ECM + obstacles (numpy copies; layout of csrc/host/flat_world.h)."""

    bbox: np.ndarray         # (4,) f32 xmin ymin xmax ymax
    vert_xy: np.ndarray      # (nV,2) f32
    vert_clear: np.ndarray   # (nV,) f32
    vert_he: np.ndarray      # (nV,) i32
    edge_v: np.ndarray       # (nE,2) i32
    edge_cl: np.ndarray      # (nE,4,2) f32  L0 R0 L1 R1
    he_next: np.ndarray      # (2nE,) i32
    obst_xy: np.ndarray      # (nO,2) f32
    obst_next: np.ndarray    # (nO,) i32
    obst_prev: np.ndarray    # (nO,) i32
    obst_convex: np.ndarray  # (nO,) u8
    obst_first: np.ndarray   # (nObst+1,) i32
    # lattice metadata (None for worlds that did not come from lattice_world())
    street_width: float | None = None
    blocks_x: np.ndarray | None = None
    blocks_y: np.ndarray | None = None

    def with_box_obstacle(self, pos, half) -> "World":
        """Copy of the world with one more box obstacle, vertices in the order of Simulator::AddObstacleArea
        (Simulator.cpp:404-407: ++, -+, --, +-; counter-clockwise) and the links / convexity flags of
        Obstacle::Initialize (ECMDataTypes.cpp:36-59).  The ECM is unchanged, as with updateECM = false."""
        import dataclasses

        x, y, hx, hy = (np.float32(v) for v in (pos[0], pos[1], half[0], half[1]))
        box = np.array([[x + hx, y + hy], [x - hx, y + hy], [x - hx, y - hy], [x + hx, y - hy]], np.float32)
        base = int(self.obst_next.shape[0])
        nxt = base + np.array([1, 2, 3, 0], np.int32)
        prv = base + np.array([3, 0, 1, 2], np.int32)
        conv = np.zeros(4, np.uint8)
        for i in range(4):
            p, q, c = box[prv[i] - base], box[nxt[i] - base], box[i]
            a, b = p - q, c - p  # isConvex = det(prev - next, p - prev) >= 0, float arithmetic
            conv[i] = 1 if np.float32(a[0] * b[1]) - np.float32(a[1] * b[0]) >= 0 else 0
        return dataclasses.replace(
            self, obst_xy=np.concatenate([self.obst_xy, box]), obst_next=np.concatenate([self.obst_next, nxt]),
            obst_prev=np.concatenate([self.obst_prev, prv]), obst_convex=np.concatenate([self.obst_convex, conv]),
            obst_first=np.concatenate([self.obst_first, [base + 4]]).astype(np.int32))

    def with_obstacle_polygons(self, polys) -> "World":
        """Copy of the world whose ORCA obstacles are the given polygons (lists of counter-clockwise points), in that
        order, with the links and convexity flags of Obstacle::Initialize (ECMDataTypes.cpp:23-61: isConvex =
        det(prev - next, p - prev) >= 0 in float arithmetic).  The ECM is unchanged, as after the reference's
        Environment::AddObstacle without a new ComputeECM (Simulator::AddObstacleArea with updateECM = false,
        Simulator.cpp:395-420): obstacles only feed FindNearestObstacles / ORCA."""
        import dataclasses

        xy, nxt, prv, conv, first = [], [], [], [], [0]
        for poly in polys:
            P = np.asarray(poly, np.float32).reshape(-1, 2)
            m, base = len(P), first[-1]
            for i in range(m):
                ip, inx = (i - 1) % m, (i + 1) % m
                a, b = P[ip] - P[inx], P[i] - P[ip]
                conv.append(1 if m == 2 or np.float32(a[0] * b[1]) - np.float32(a[1] * b[0]) >= 0 else 0)
                nxt.append(base + inx)
                prv.append(base + ip)
            xy.append(P)
            first.append(base + m)
        return dataclasses.replace(
            self, obst_xy=np.concatenate(xy).astype(np.float32), obst_next=np.array(nxt, np.int32), obst_prev=np.array(prv, np.int32),
            obst_convex=np.array(conv, np.uint8), obst_first=np.array(first, np.int32))

    def obstacle_polygons(self):
        return [self.obst_xy[self.obst_first[k]:self.obst_first[k + 1]].copy() for k in range(self.n_obstacles)]

    def with_recessed_obstacles(self, seed: int, depth=(1.0, 3.0), every: int = 1) -> "World":
        """Every `every`-th obstacle gets recesses cut into its footprint: a notch in the middle of one side (a U shape,
        two concave vertices) and one corner cut away (an L shape, one concave vertex), so that the !isConvex legs of
        ORCA::GenerateConstraints (ORCA.cpp:146-212) and the convexity rule (ECMDataTypes.cpp:52-59) see real input.
        Only the obstacle polygons change (see with_obstacle_polygons); the streets and their ECM stay as they are."""
        rng = np.random.default_rng(seed)
        out = []
        for k, P in enumerate(self.obstacle_polygons()):
            P = P.astype(np.float64)
            m = len(P)
            if k % every or m != 4:
                out.append(P)
                continue
            side, corner = int(rng.integers(0, 4)), int(rng.integers(0, 4))
            d_side, d_corner = rng.uniform(depth[0], depth[1], size=2)
            Q = []
            for i in range(4):
                a, b = P[i], P[(i + 1) % 4]
                e = b - a
                L = float(np.hypot(*e))
                u = e / L
                n_in = np.array([-u[1], u[0]])  # counter-clockwise polygon: the interior lies to the left of every edge
                pprev = P[(i - 1) % 4]
                uprev = (a - pprev) / float(np.hypot(*(a - pprev)))
                if i == corner and min(L, float(np.hypot(*(a - pprev)))) > 4 * d_corner:
                    # L shape: vertex a is replaced by three points (the middle one is concave)
                    Q += [a - uprev * d_corner, a - uprev * d_corner + u * d_corner, a + u * d_corner]
                else:
                    Q.append(a)
                if i == side and L > 6 * d_side:
                    t0, t1 = 0.3 * L, 0.7 * L
                    Q += [a + u * t0, a + u * t0 + n_in * d_side, a + u * t1 + n_in * d_side, a + u * t1]
            out.append(np.array(Q))
        return self.with_obstacle_polygons(out)

    def rotated(self, angle: float) -> "World":
        """Copy of the world turned by `angle` (radians, counter-clockwise about the origin): every vertex, closest
        point and obstacle vertex is rotated in double precision and rounded to float once, so all cell edges and
        obstacle segments become oblique.  bbox becomes the bounding box of the turned walkable area."""
        import dataclasses

        c, s_ = np.cos(angle), np.sin(angle)
        R = np.array([[c, -s_], [s_, c]])

        def rot(a):
            return (a.astype(np.float64).reshape(-1, 2) @ R.T).astype(np.float32).reshape(a.shape)

        bb = self.bbox.astype(np.float64)
        corners = np.array([[bb[0], bb[1]], [bb[2], bb[1]], [bb[2], bb[3]], [bb[0], bb[3]]]) @ R.T
        nb = np.array([corners[:, 0].min(), corners[:, 1].min(), corners[:, 0].max(), corners[:, 1].max()], np.float32)
        w = dataclasses.replace(self, bbox=nb, vert_xy=rot(self.vert_xy), edge_cl=rot(self.edge_cl), obst_xy=rot(self.obst_xy),
                                street_width=None, blocks_x=None, blocks_y=None)
        # convexity is a float predicate of the turned coordinates (ECMDataTypes.cpp:52-59): recompute it
        return w.with_obstacle_polygons(w.obstacle_polygons())

    @property
    def n_vertices(self) -> int:
        return int(self.vert_clear.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.edge_v.shape[0])

    @property
    def n_cells(self) -> int:
        return 2 * self.n_edges

    @property
    def n_obstacles(self) -> int:
        return int(self.obst_first.shape[0]) - 1

    @property
    def n_obst_vertices(self) -> int:
        return int(self.obst_next.shape[0])


def _copy(ptr, n, dtype, shape):
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True).reshape(shape)


def lattice_world(blocks_x, blocks_y, street_width: float, x0: float = 0.0, y0: float = 0.0) -> World:
    """Blocks blocks_x[i] x blocks_y[j] separated by streets of one width (see lattice_world.h)."""
    bx = np.ascontiguousarray(blocks_x, dtype=np.float32)
    by = np.ascontiguousarray(blocks_y, dtype=np.float32)
    L = lib()
    h = L.ecmhost_lattice_world(len(bx), fptr(bx), len(by), fptr(by), float(street_width), float(x0), float(y0))
    if not h:
        raise ValueError("invalid lattice world parameters")
    try:
        v = _WorldView()
        L.ecmhost_world_get_view(h, C.byref(v))
        nV, nE, nO, nB = v.n_vertices, v.n_edges, v.n_obst_vertices, v.n_obstacles
        w = World(
            bbox=np.array(list(v.bbox), dtype=np.float32),
            vert_xy=_copy(v.vert_xy, 2 * nV, np.float32, (nV, 2)),
            vert_clear=_copy(v.vert_clear, nV, np.float32, (nV,)),
            vert_he=_copy(v.vert_he, nV, np.int32, (nV,)),
            edge_v=_copy(v.edge_v, 2 * nE, np.int32, (nE, 2)),
            edge_cl=_copy(v.edge_cl, 8 * nE, np.float32, (nE, 4, 2)),
            he_next=_copy(v.he_next, 2 * nE, np.int32, (2 * nE,)),
            obst_xy=_copy(v.obst_xy, 2 * nO, np.float32, (nO, 2)),
            obst_next=_copy(v.obst_next, nO, np.int32, (nO,)),
            obst_prev=_copy(v.obst_prev, nO, np.int32, (nO,)),
            obst_convex=_copy(v.obst_convex, nO, np.uint8, (nO,)),
            obst_first=_copy(v.obst_first, nB + 1, np.int32, (nB + 1,)),
            street_width=float(street_width),
            blocks_x=bx.copy(),
            blocks_y=by.copy(),
        )
    finally:
        L.ecmhost_world_free(h)
    return w


def _world_from_handle(L, h, **meta) -> World:
    v = _WorldView()
    L.ecmhost_world_get_view(h, C.byref(v))
    nV, nE, nO, nB = v.n_vertices, v.n_edges, v.n_obst_vertices, v.n_obstacles
    return World(
        bbox=np.array(list(v.bbox), dtype=np.float32),
        vert_xy=_copy(v.vert_xy, 2 * nV, np.float32, (nV, 2)),
        vert_clear=_copy(v.vert_clear, nV, np.float32, (nV,)),
        vert_he=_copy(v.vert_he, nV, np.int32, (nV,)),
        edge_v=_copy(v.edge_v, 2 * nE, np.int32, (nE, 2)),
        edge_cl=_copy(v.edge_cl, 8 * nE, np.float32, (nE, 4, 2)),
        he_next=_copy(v.he_next, 2 * nE, np.int32, (2 * nE,)),
        obst_xy=_copy(v.obst_xy, 2 * nO, np.float32, (nO, 2)),
        obst_next=_copy(v.obst_next, nO, np.int32, (nO,)),
        obst_prev=_copy(v.obst_prev, nO, np.int32, (nO,)),
        obst_convex=_copy(v.obst_convex, nO, np.uint8, (nO,)),
        obst_first=_copy(v.obst_first, nB + 1, np.int32, (nB + 1,)),
        **meta,
    )


def polygon_world(bbox, polygons) -> World:
    """ECM of a rectangular walkable area with polygonal obstacles (counter-clockwise point lists strictly inside it),
    built without Boost by csrc/host/polygon_world.cpp; e.g. the reference's DEBUG1 scene (Environment.cpp:147-171):
    polygon_world((-500, -500, 500, 500), [[(-50, 50), (-150, 50), (-150, -50), (-50, -50)], [(150, 50), (50, 50), (50, -50), (150, -50)]])."""
    bb = np.ascontiguousarray(bbox, np.float32)
    first = np.zeros(len(polygons) + 1, np.int32)
    pts = []
    for k, poly in enumerate(polygons):
        p = np.asarray(poly, np.float32).reshape(-1, 2)
        pts.append(p)
        first[k + 1] = first[k] + len(p)
    xy = np.ascontiguousarray(np.concatenate(pts) if pts else np.zeros((0, 2), np.float32))
    L = lib()
    err = C.create_string_buffer(256)
    h = L.ecmhost_polygon_world(fptr(bb), len(polygons), iptr(first), fptr(xy) if len(xy) else None, err, 256)
    if not h:
        raise ValueError(f"polygon_world: {err.value.decode() or 'invalid input'}")
    try:
        return _world_from_handle(L, h)
    finally:
        L.ecmhost_world_free(h)


class _WorldHandle:
    """A World re-materialised inside libecmhost (ecmhost_world_from_arrays)."""

    def __init__(self, w: World):
        self.L = lib()
        self._keep = [np.ascontiguousarray(a) for a in (w.vert_xy, w.vert_clear, w.vert_he, w.edge_v, w.edge_cl, w.he_next,
                                                        w.obst_xy, w.obst_next, w.obst_prev, w.obst_convex, w.obst_first)]
        k = self._keep
        v = _WorldView()
        for i in range(4):
            v.bbox[i] = float(w.bbox[i])
        v.n_vertices, v.n_edges, v.n_obst_vertices, v.n_obstacles = w.n_vertices, w.n_edges, w.n_obst_vertices, w.n_obstacles
        v.vert_xy, v.vert_clear, v.vert_he = fptr(k[0]), fptr(k[1]), iptr(k[2])
        v.edge_v, v.edge_cl, v.he_next = iptr(k[3]), fptr(k[4]), iptr(k[5])
        v.obst_xy, v.obst_next, v.obst_prev, v.obst_convex, v.obst_first = fptr(k[6]), iptr(k[7]), iptr(k[8]), u8ptr(k[9]), iptr(k[10])
        self.h = self.L.ecmhost_world_from_arrays(C.byref(v))

    def __del__(self):
        try:
            if self.h:
                self.L.ecmhost_world_free(self.h)
                self.h = None
        except Exception:
            pass


def plan_paths(w: World, start, goal, clearance, threads: int = 0):
    """Batched ECMPathPlanner::FindPath.  Returns (path_off[n+1], path_xy[total,2], n_ok)."""
    start = np.ascontiguousarray(start, np.float32)
    goal = np.ascontiguousarray(goal, np.float32)
    clearance = np.ascontiguousarray(clearance, np.float32)
    n = len(start)
    wh = _WorldHandle(w)
    L = lib()
    p = L.ecmhost_plan_paths(wh.h, n, fptr(start), fptr(goal), fptr(clearance), int(threads))
    if not p:
        raise RuntimeError("ecmhost_plan_paths failed")
    try:
        off = np.ctypeslib.as_array(L.ecmhost_paths_offsets(p), shape=(n + 1,)).astype(np.int32, copy=True)
        total = int(off[-1])
        xy = (np.ctypeslib.as_array(L.ecmhost_paths_xy(p), shape=(2 * total,)).astype(np.float32, copy=True).reshape(total, 2)
              if total > 0 else np.zeros((0, 2), np.float32))
        ok = L.ecmhost_paths_succeeded(p)
    finally:
        L.ecmhost_paths_free(p)
    return off, xy, ok


def find_cells(w: World, xy):
    xy = np.ascontiguousarray(xy, np.float32)
    out = np.zeros(len(xy), np.int32)
    wh = _WorldHandle(w)
    lib().ecmhost_find_cells(wh.h, len(xy), fptr(xy), iptr(out))
    return out
