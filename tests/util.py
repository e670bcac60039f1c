"""Shared helpers of the test-suite (test infrastructure)."""
import os

import numpy as np

from ecmgenerator_b200.host import World
from ecmgenerator_b200.scenarios import Crowd

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ("c1_small", "c2_small", "jam_small")
WORLD_KEYS = ("bbox", "vert_xy", "vert_clear", "vert_he", "edge_v", "edge_cl", "he_next", "obst_xy", "obst_next",
              "obst_prev", "obst_convex", "obst_first")


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        kw = {k: self.z["world/" + k] for k in WORLD_KEYS}
        self.world = World(street_width=float(self.z["world/street_width"]), blocks_x=self.z["world/blocks_x"],
                           blocks_y=self.z["world/blocks_y"], **kw)
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
