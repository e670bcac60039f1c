"""Shared helpers of the test-suite (test infrastructure)."""
import os

import numpy as np

from ecmgenerator_b200.host import World
from ecmgenerator_b200.scenarios import Crowd

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ("c1_small", "c2_small", "jam_small", "oblique_small", "concave_small", "yard_small")
WORLD_KEYS = ("bbox", "vert_xy", "vert_clear", "vert_he", "edge_v", "edge_cl", "he_next", "obst_xy", "obst_next",
              "obst_prev", "obst_convex", "obst_first")


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        kw = {k: self.z["world/" + k] for k in WORLD_KEYS}
        sw = float(self.z["world/street_width"])
        lattice = sw == sw  # NaN: a turned / edited world without the lattice metadata
        self.world = World(street_width=sw if lattice else None, blocks_x=self.z["world/blocks_x"] if lattice else None,
                           blocks_y=self.z["world/blocks_y"] if lattice else None, **kw)
        self.crowd = Crowd(self.z["crowd/pos"], self.z["crowd/goal"], self.z["crowd/radius"], self.z["crowd/speed"])
        self.path_off, self.path_xy = self.z["crowd/path_off"], self.z["crowd/path_xy"]
        self.step = float(self.z["step"])
        self.n = self.crowd.n

    def ticks(self, mode):
        return self.z[f"{mode}/pos"].shape[0]

    def events_at(self, mode, t):
        """[(slot, kind, new_path or None)] the harness applied after tick t."""
        ev = self.z[f"{mode}/events"]
        out = []
        for j, (tt, slot, kind) in enumerate(ev):
            if tt == t:
                out.append((int(slot), int(kind), self.z[f"{mode}/newpath_{j}"] if kind == 0 else None))
        return out


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(bits(a), bits(b)):
        bad = np.nonzero((bits(a) != bits(b)).reshape(a.shape[0], -1).any(axis=1))[0]
        raise AssertionError(f"{what}: {len(bad)} rows differ, first {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}")


def apply_events(sim, events):
    for slot, kind, path in events:
        if kind == 1:
            sim.destroy_agent(slot)
        else:
            sim.set_path(slot, path)


def check_window_against_oracle(world, step, before, after, radius, speed, path_off, path_xy, window, margin, vel_tol, label=""):
    """Locality check that scales to any crowd size: one tick of the agents inside `window` = (x0, y0, x1, y1) depends
    only on the agents within their 5-NN reach, so the C oracle stepping the SUB-crowd `window + margin` must reproduce
    what the device computed for the window's agents in the full crowd - neighbour lists, cells, attraction points and
    preferred velocities bit for bit, velocities within `vel_tol`.

    before: pos, vel, attraction, active of every slot before the tick; after: pos, vel, prefvel, attraction, active,
    nbr, nbr_cnt, cell after it.  Returns statistics."""
    from oracle.pyoracle import OracleSim

    x0, y0, x1, y1 = window
    p = before["pos"]
    act = before["active"] > 0
    member = act & (p[:, 0] >= x0 - margin) & (p[:, 0] < x1 + margin) & (p[:, 1] >= y0 - margin) & (p[:, 1] < y1 + margin)
    members = np.flatnonzero(member)  # ascending slots: the (distance, slot id) tie-break keeps its meaning
    inner_l = np.flatnonzero((p[members, 0] >= x0) & (p[members, 0] < x1) & (p[members, 1] >= y0) & (p[members, 1] < y1))
    inner = members[inner_l]
    assert len(inner) >= 50, f"{label}: window holds only {len(inner)} agents"
    m = len(members)
    local = np.full(len(p), -1, np.int64)
    local[members] = np.arange(m)
    nb = after["nbr"][inner]
    cnt = after["nbr_cnt"][inner]
    assert (cnt == 5).all()
    assert (local[nb] >= 0).all(), f"{label}: margin {margin} m does not cover the 5-NN reach of the window"
    lens = (path_off[1:] - path_off[:-1])[members]
    off = np.zeros(m + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    idx = np.repeat(path_off[:-1][members].astype(np.int64), lens) + (np.arange(int(lens.sum())) - np.repeat(off[:-1].astype(np.int64), lens))
    ora = OracleSim(world, m + 8, step, "exact-knn")
    ora.bulk_load(p[members], radius[members], speed[members], off, path_xy[idx])
    for k, s in enumerate(members):
        ora.set_kinematics(k, p[s], before["vel"][s])
        ora.set_attraction(k, before["attraction"][s])
    ids_o, cnt_o = ora.query_neighbors(m)
    cells_o = ora.query_cells(p[inner])
    ora.step(1)
    st = ora.state(m)
    ora.close()
    assert np.array_equal(cnt_o[inner_l], cnt), f"{label}: neighbour counts"
    assert np.array_equal(members[ids_o[inner_l]], nb), f"{label}: neighbour lists"
    assert np.array_equal(st["active"][inner_l] > 0, after["active"][inner] > 0), f"{label}: active flags"
    alive = st["active"][inner_l] > 0
    cell = after["cell"][inner]
    evaluated = cell != -2
    assert np.array_equal(cell[evaluated], cells_o[evaluated]), f"{label}: ECM cells"
    assert_bits_equal(after["attraction"][inner], st["attraction"][inner_l], f"{label}: attraction points")
    assert_bits_equal(after["prefvel"][inner][alive], st["prefvel"][inner_l][alive], f"{label}: preferred velocities")
    dv = np.abs(after["vel"][inner][alive] - st["vel"][inner_l][alive])
    assert dv.max() <= vel_tol, f"{label}: max |dv| = {dv.max()}"
    dp = np.abs(after["pos"][inner][alive] - st["pos"][inner_l][alive])
    assert dp.max() <= vel_tol, f"{label}: max |dp| = {dp.max()}"
    same = (after["vel"][inner][alive].view(np.uint32) == st["vel"][inner_l][alive].view(np.uint32)).all(axis=1)
    return {"members": m, "inner": int(len(inner)), "evaluated_cells": int(evaluated.sum()), "max_dv": float(dv.max()),
            "velocity_rows_bit_identical": float(same.mean()), "moving": float((np.linalg.norm(before["vel"][inner], axis=1) > 1e-3).mean()),
            "mean_speed": float(np.linalg.norm(before["vel"][inner], axis=1).mean())}
