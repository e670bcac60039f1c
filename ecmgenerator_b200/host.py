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
        L.ecmhost_world_from_arrays.restype = C.c_void_p
        L.ecmhost_world_from_arrays.argtypes = [C.POINTER(_WorldView)]
        L.ecmhost_world_free.argtypes = [C.c_void_p]
        L.ecmhost_world_get_view.argtypes = [C.c_void_p, C.POINTER(_WorldView)]
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
