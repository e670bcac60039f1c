"""Synthetic worlds and crowds for the configurations named in BASELINE.json.

Everything is seeded (numpy PCG64) and deterministic.  Worlds are lattice worlds
(csrc/host/lattice_world.h); crowds are jittered-grid samples of the street area so that no two
agents overlap at t=0, with goals drawn from the same free space at a minimum distance.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .host import World, lattice_world


@dataclass
class Crowd:
    pos: np.ndarray     # (N,2) f32 start positions
    goal: np.ndarray    # (N,2) f32
    radius: np.ndarray  # (N,) f32   (the reference's "clearance")
    speed: np.ndarray   # (N,) f32   preferred speed

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    def take(self, idx) -> "Crowd":
        return Crowd(self.pos[idx].copy(), self.goal[idx].copy(), self.radius[idx].copy(), self.speed[idx].copy())


def _street_centres(w: World):
    bx, by, W = w.blocks_x.astype(np.float64), w.blocks_y.astype(np.float64), float(w.street_width)
    x0, y0 = float(w.bbox[0]), float(w.bbox[1])
    cx = x0 + np.cumsum(bx)[:-1] + W * np.arange(len(bx) - 1) + 0.5 * W
    cy = y0 + np.cumsum(by)[:-1] + W * np.arange(len(by) - 1) + 0.5 * W
    return cx, cy


def free_mask(w: World, pts: np.ndarray, margin: float) -> np.ndarray:
    """True where a point lies in a street at least `margin` away from every wall."""
    assert w.street_width is not None, "free_mask needs a lattice world"
    cx, cy = _street_centres(w)
    h = 0.5 * float(w.street_width) - margin
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)

    def near(c, v):
        if len(c) == 0:
            return np.zeros(v.shape, dtype=bool)
        k = np.searchsorted(c, v)
        lo = c[np.clip(k - 1, 0, len(c) - 1)]
        hi = c[np.clip(k, 0, len(c) - 1)]
        return np.minimum(np.abs(v - lo), np.abs(v - hi)) <= h

    inside = (x >= w.bbox[0] + margin) & (x <= w.bbox[2] - margin) & (y >= w.bbox[1] + margin) & (y <= w.bbox[3] - margin)
    return inside & (near(cx, x) | near(cy, y))


def free_area(w: World) -> float:
    bx, by = w.blocks_x.astype(np.float64), w.blocks_y.astype(np.float64)
    total = float(w.bbox[2] - w.bbox[0]) * float(w.bbox[3] - w.bbox[1])
    return total - float(bx.sum() * by.sum())


def sample_crowd(w: World, n: int, seed: int, radius=(0.3, 0.3), speed=(1.4, 1.4), min_goal_dist: float = 0.0,
                 wall_margin: float = 0.15, window=None) -> Crowd:
    """n agents on a jittered grid over the streets (optionally restricted to window=(x0,y0,x1,y1))."""
    rng = np.random.default_rng(seed)
    rmax = float(max(radius))
    margin = rmax + wall_margin
    bb = np.array(window if window is not None else w.bbox, dtype=np.float64)
    # usable street area shrinks by the margin on both sides of every street
    W = float(w.street_width)
    frac = max(W - 2 * margin, 1e-3) / W
    area = free_area(w) * frac * ((bb[2] - bb[0]) * (bb[3] - bb[1])) / (float(w.bbox[2] - w.bbox[0]) * float(w.bbox[3] - w.bbox[1]))
    s = np.sqrt(area / (n * 1.15))
    for _ in range(12):
        gx = np.arange(bb[0] + 0.5 * s, bb[2], s)
        gy = np.arange(bb[1] + 0.5 * s, bb[3], s)
        jitter = max(0.0, 0.5 * (s - 2 * rmax - 0.05))
        X, Y = np.meshgrid(gx, gy, indexing="xy")
        pts = np.stack([X.ravel(), Y.ravel()], axis=1)
        pts += rng.uniform(-jitter, jitter, size=pts.shape)
        pts = pts[free_mask(w, pts, margin)]
        if len(pts) >= n:
            break
        s *= 0.93
    else:
        raise ValueError(f"could not place {n} agents (free area too small?)")
    if s < 2 * rmax + 0.05:
        raise ValueError(f"crowd of {n} does not fit without overlap (spacing {s:.3f})")
    sel = rng.permutation(len(pts))[:n]
    pos = pts[sel]
    # goals: other free points, far enough away
    pool = pts if len(pts) > 4 * n else None
    if pool is None:
        gs = max(s * 0.5, 0.25)
        gx = np.arange(float(w.bbox[0]) + 0.5 * gs, float(w.bbox[2]), gs)
        gy = np.arange(float(w.bbox[1]) + 0.5 * gs, float(w.bbox[3]), gs)
        if len(gx) * len(gy) > 4_000_000:
            k = int(np.ceil(np.sqrt(len(gx) * len(gy) / 4_000_000)))
            gx, gy = gx[::k], gy[::k]
        X, Y = np.meshgrid(gx, gy, indexing="xy")
        pool = np.stack([X.ravel(), Y.ravel()], axis=1)
        pool = pool[free_mask(w, pool, margin)]
    goal = pool[rng.integers(0, len(pool), size=n)]
    for _ in range(64):
        bad = np.hypot(*(goal - pos).T) < min_goal_dist
        if not bad.any():
            break
        goal[bad] = pool[rng.integers(0, len(pool), size=int(bad.sum()))]
    rad = rng.uniform(radius[0], radius[1], size=n) if radius[1] > radius[0] else np.full(n, radius[0])
    spd = rng.uniform(speed[0], speed[1], size=n) if speed[1] > speed[0] else np.full(n, speed[0])
    return Crowd(pos.astype(np.float32), goal.astype(np.float32), rad.astype(np.float32), spd.astype(np.float32))


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs.  dt = 1/60 everywhere.
# ------------------------------------------------------------------------------------------------
DT = np.float32(1.0 / 60.0)


def _blocks(n, lo, hi, rng):
    return np.round(rng.uniform(lo, hi, size=n)).astype(np.float32)


def world_c1() -> World:
    """C1: 300 x 300 world, 3 x 3 box obstacles of 80 x 80, 30-wide streets (small environment)."""
    return lattice_world([80.0, 80.0, 80.0], [80.0, 80.0, 80.0], 30.0, -150.0, -150.0)


def world_c2(seed: int = 2) -> World:
    """C2: ~500 x 500, 200 blocks (20 x 10 non-uniform lattice), 8-wide streets."""
    rng = np.random.default_rng(seed)
    bx = _blocks(20, 14.0, 20.0, rng)
    by = _blocks(10, 36.0, 48.0, rng)
    return lattice_world(bx, by, 8.0, 0.0, 0.0)


def world_c3() -> World:
    """C3: 45 x 45 = 2025 blocks of 40 x 40 with 20-wide streets (2.68 km square, centred on the origin).

    Worlds must keep |coordinate| < 2048: beyond that half a float ulp exceeds the reference's
    EPSILON (1e-4), Point::Approximate(p, p) turns false and its funnel never terminates."""
    return lattice_world(np.full(45, 40.0), np.full(45, 40.0), 20.0, -1340.0, -1340.0)


def world_c4(seed: int = 4) -> World:
    """C4: ~2 km x 2 km = 4 km^2, four districts with different block sizes (~8k blocks), 10-wide streets."""
    rng = np.random.default_rng(seed)
    bx = np.concatenate([_blocks(50, 8.0, 12.0, rng), _blocks(40, 12.0, 18.0, rng)])
    by = np.concatenate([_blocks(45, 10.0, 14.0, rng), _blocks(45, 10.0, 16.0, rng)])
    w = float(bx.sum() + 10.0 * (len(bx) - 1)), float(by.sum() + 10.0 * (len(by) - 1))
    return lattice_world(bx, by, 10.0, -round(w[0] / 2), -round(w[1] / 2))


def world_c5() -> World:
    """C5: bidirectional corridor stress: long 10-wide streets between 100-long blocks."""
    return lattice_world(np.full(12, 100.0), np.full(24, 12.0), 10.0, -655.0, -259.0)


def crowd_c1(w: World, n: int = 5_000, seed: int = 1) -> Crowd:
    return sample_crowd(w, n, seed, radius=(0.3, 0.3), speed=(1.4, 1.4), min_goal_dist=40.0)


def crowd_c2(w: World, n: int = 50_000, seed: int = 2) -> Crowd:
    return sample_crowd(w, n, seed, radius=(0.2, 0.4), speed=(1.0, 1.6), min_goal_dist=100.0)


def crowd_c3(w: World, n: int = 1_000_000, seed: int = 3) -> Crowd:
    return sample_crowd(w, n, seed, radius=(0.3, 0.3), speed=(1.4, 1.4), min_goal_dist=300.0)


def crowd_c4(w: World, n: int = 4_000_000, seed: int = 4) -> Crowd:
    return sample_crowd(w, n, seed, radius=(0.25, 0.25), speed=(1.2, 1.5), min_goal_dist=300.0, wall_margin=0.05)


def crowd_c5(w: World, n: int = 250_000, seed: int = 5) -> Crowd:
    """Half of the agents head east, half west, along the same horizontal streets."""
    c = sample_crowd(w, n, seed, radius=(0.3, 0.3), speed=(1.4, 1.4), min_goal_dist=0.0)
    xmid = 0.5 * float(w.bbox[0] + w.bbox[2])
    goal = c.pos.copy()
    span = float(w.bbox[2] - w.bbox[0])
    east = c.pos[:, 0] < xmid
    goal[:, 0] = np.where(east, c.pos[:, 0] + 0.45 * span, c.pos[:, 0] - 0.45 * span).astype(np.float32)
    # goals must lie in free space: keep y (same street) when the start is in a horizontal street,
    # otherwise fall back to the sampled goal
    ok = free_mask(w, goal, 0.45)
    goal[~ok] = c.goal[~ok]
    return Crowd(c.pos, goal.astype(np.float32), c.radius, c.speed)


# ------------------------------------------------------------------------------------------------
# Polygonal scenes for host.polygon_world (the Boost-free ECM generator): the reference's own test environments
# (Environment.cpp:27-185) and one with oblique, non-rectangular obstacles.  (bbox, [counter-clockwise polygons]).
# ------------------------------------------------------------------------------------------------
SCENES = {
    # Environment::TestEnvironment::DEBUG1 (Environment.cpp:147-171): two boxes
    "debug1": ((-500, -500, 500, 500), [[(-50, 50), (-150, 50), (-150, -50), (-50, -50)], [(150, 50), (50, 50), (50, -50), (150, -50)]]),
    # Environment::TestEnvironment::CLASSIC (Environment.cpp:27-48): one U-shaped obstacle (two concave vertices)
    "classic": ((-500, -500, 500, 500), [[(200, -250), (200, 250), (100, 250), (100, -150), (-100, -150), (-100, 250), (-200, 250), (-200, -250)]]),
    # Environment::TestEnvironment::SQUARE (Environment.cpp:49-66)
    "square": ((-500, -500, 500, 500), [[(200, -250), (200, 250), (-200, 250), (-200, -250)]]),
    # not from the reference: a triangle, a turned rectangle, an L and a pentagon in a 120 x 90 yard
    "yard": ((-60, -45, 60, 45), [[(-40, -30), (-22, -33), (-30, -12)],
                                  [(-5, -20), (18, -28), (22, -17), (-1, -9)],
                                  [(30, -5), (48, -5), (48, 25), (40, 25), (40, 3), (30, 3)],
                                  [(-35, 8), (-20, 5), (-12, 18), (-22, 30), (-38, 24)]]),
}


def scene_polygons(name: str):
    bbox, polys = SCENES[name]
    return bbox, polys


def crowd_in_scene(w: World, n: int, seed: int, radius=0.3, speed=1.4, window=None, min_goal_dist: float = 30.0, spacing: float = 1.2) -> Crowd:
    """n agents on a jittered grid over the free space of a polygonal scene (inside `window` = x0 y0 x1 y1), goals drawn
    from the same points at least min_goal_dist away."""
    rng = np.random.default_rng(seed)
    bb = np.array(window if window is not None else w.bbox, np.float64)
    gx = np.arange(bb[0] + spacing, bb[2] - spacing, spacing)
    gy = np.arange(bb[1] + spacing, bb[3] - spacing, spacing)
    X, Y = np.meshgrid(gx, gy, indexing="xy")
    pts = np.stack([X.ravel(), Y.ravel()], axis=1) + rng.uniform(-0.25, 0.25, size=(X.size, 2)) * (spacing - 2 * radius - 0.1)
    # keep points at least radius + 0.3 away from every obstacle edge and outside the obstacles
    keep = np.ones(len(pts), bool)
    for k in range(w.n_obstacles):
        poly = w.obst_xy[w.obst_first[k]:w.obst_first[k + 1]].astype(np.float64)
        inside = np.zeros(len(pts), bool)
        dmin = np.full(len(pts), np.inf)
        for i in range(len(poly)):
            p, q = poly[i], poly[(i + 1) % len(poly)]
            with np.errstate(divide="ignore", invalid="ignore"):
                inside ^= ((p[1] > pts[:, 1]) != (q[1] > pts[:, 1])) & (pts[:, 0] < (q[0] - p[0]) * (pts[:, 1] - p[1]) / (q[1] - p[1]) + p[0])
            d = q - p
            t = np.clip(((pts - p) @ d) / (d @ d), 0, 1)
            dmin = np.minimum(dmin, np.linalg.norm(pts - (p + t[:, None] * d), axis=1))
        keep &= ~inside & (dmin > radius + 0.3)
    pts = pts[keep]
    if len(pts) < n:
        raise ValueError(f"only {len(pts)} free grid points for {n} agents")
    sel = rng.permutation(len(pts))[:n]
    pos = pts[sel]
    goal = pts[rng.integers(0, len(pts), size=n)]
    for _ in range(64):
        bad = np.hypot(*(goal - pos).T) < min_goal_dist
        if not bad.any():
            break
        goal[bad] = pts[rng.integers(0, len(pts), size=int(bad.sum()))]
    return Crowd(pos.astype(np.float32), goal.astype(np.float32), np.full(n, radius, np.float32), np.full(n, speed, np.float32))


CONFIGS = {
    "c1_5k": (world_c1, crowd_c1),
    "c2_50k": (world_c2, crowd_c2),
    "c3_1m": (world_c3, crowd_c3),
    "c4_4m": (world_c4, crowd_c4),
    "c5_250k": (world_c5, crowd_c5),
}
