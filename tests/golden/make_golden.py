"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference compiled here
(oracle/_ref, see oracle/Makefile).  Run in the build container only (needs /root/reference to have
been compiled):   python tests/golden/make_golden.py

Each <name>.npz holds a small world, a crowd with paths planned by the reference's own
ECMPathPlanner, and what the reference computed: per-tick component arrays for both neighbour
modes ("ref-kdtree" = reference as is, "exact-knn" = KDTree.cpp swapped for oracle/kdtree_exact.cpp),
neighbour lists, cell ids and retractions of probe points.  The GPU box has no /root/reference;
tests there compare against these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ecmgenerator_b200 import scenarios as S  # noqa: E402
from ecmgenerator_b200.host import lattice_world  # noqa: E402
from oracle.pyref import RefSim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STEP = np.float32(1.0 / 60.0)

WORLD_KEYS = ("bbox", "vert_xy", "vert_clear", "vert_he", "edge_v", "edge_cl", "he_next", "obst_xy", "obst_next",
              "obst_prev", "obst_convex", "obst_first")


def probe_points(w, rng, n):
    """Uniform points over the bbox (+ a margin outside), points on cell corners/edges and exact grid lines."""
    bb = w.bbox.astype(np.float64)
    a = rng.uniform([bb[0] - 5, bb[1] - 5], [bb[2] + 5, bb[3] + 5], size=(n, 2))
    corners = w.vert_xy[rng.integers(0, w.n_vertices, size=n // 8)].astype(np.float64)
    near = corners + rng.normal(0, 2e-4, size=corners.shape)
    cl = w.edge_cl.reshape(-1, 2)[rng.integers(0, 4 * w.n_edges, size=n // 8)].astype(np.float64)
    level = a[: n // 8].copy()
    level[:, 1] = corners[: len(level), 1]  # exactly level with a vertex: the even-odd test's blind spot
    mid = 0.5 * (w.vert_xy[w.edge_v[:, 0]] + w.vert_xy[w.edge_v[:, 1]])[rng.integers(0, w.n_edges, size=n // 8)]
    return np.concatenate([a, corners, near, cl, level, mid]).astype(np.float32)


def run_mode(w, crowd, mode, ticks, paths=None):
    n = crowd.n
    r = RefSim(w, n + 8, STEP, mode)
    if paths is None:
        slots = r.bulk_load(crowd.pos, crowd.goal, crowd.radius, crowd.speed)
    else:
        slots = r.bulk_load(crowd.pos, None, crowd.radius, crowd.speed, paths[0], paths[1])
    assert (slots == np.arange(n)).all(), "every agent must get a usable path"
    off, pxy = r.paths(n)
    nb0 = r.query_neighbors(n)
    keys = ("pos", "vel")  # per tick; the other components are kept at the ticks in `full_at`
    allkeys = ("pos", "vel", "prefvel", "attraction", "force")
    full_at = sorted({0, ticks // 2, ticks - 1})
    full = {}
    hist = {k: np.zeros((ticks, n, 2), np.float32) for k in keys}
    act = np.zeros((ticks, n), np.uint8)
    events = []  # (tick, slot, kind) kind 0 = replanned (new path stored), 1 = poisoned -> destroyed by the harness
    new_paths = {}
    prev_len = np.diff(off)
    plist = [pxy[off[i]:off[i + 1]].copy() for i in range(n)]
    for t in range(ticks):
        r.step(1)
        st = r.state(n)
        for k in keys:
            hist[k][t] = st[k]
        act[t] = st["active"]
        if t in full_at:
            for k in allkeys:
                full[f"full{t}_{k}"] = st[k].copy()
        # detect replans: the reference replaced the path inside the tick (Simulator.cpp:581-587)
        for i in range(n):
            if not st["active"][i]:
                continue
            L = r.path_len(i)
            p = r.path(i) if L > 0 else np.zeros((0, 2), np.float32)
            if L != len(plist[i]) or not np.array_equal(p, plist[i]):
                if L < 2:
                    # FindPath failed: the reference now holds a 0-point path and would read path.x[-1]
                    # next tick (undefined behaviour).  The harness destroys the agent on every side.
                    r.destroy_agent(i)
                    events.append((t, i, 1))
                else:
                    events.append((t, i, 0))
                    new_paths[(t, i)] = p.copy()
                plist[i] = p.copy()
    nb1 = r.query_neighbors(n)
    final_state = r.state(n)
    r.close()
    out = {f"{mode}/{k}": v for k, v in hist.items()}
    out[f"{mode}/active"] = act
    out[f"{mode}/full_at"] = np.array(full_at, np.int32)
    for k, v in full.items():
        out[f"{mode}/{k}"] = v
    out[f"{mode}/nbr0_ids"], out[f"{mode}/nbr0_cnt"] = nb0
    out[f"{mode}/nbr1_ids"], out[f"{mode}/nbr1_cnt"] = nb1
    out[f"{mode}/final_pos"] = final_state["pos"]
    ev = np.array(events, np.int32).reshape(-1, 3)
    out[f"{mode}/events"] = ev
    for j, (t, i, kind) in enumerate(events):
        if kind == 0:
            out[f"{mode}/newpath_{j}"] = new_paths[(t, i)]
    return out, (off, pxy)


def make(name, w, crowd, ticks, seed):
    rng = np.random.default_rng(seed)
    data = {f"world/{k}": getattr(w, k) for k in WORLD_KEYS}
    # lattice metadata; NaN / empty for worlds that were turned or otherwise edited (no closed-form street mask)
    data["world/street_width"] = np.float32(w.street_width if w.street_width is not None else np.nan)
    data["world/blocks_x"] = w.blocks_x if w.blocks_x is not None else np.zeros(0, np.float32)
    data["world/blocks_y"] = w.blocks_y if w.blocks_y is not None else np.zeros(0, np.float32)
    data["step"] = STEP
    data["crowd/pos"], data["crowd/goal"] = crowd.pos, crowd.goal
    data["crowd/radius"], data["crowd/speed"] = crowd.radius, crowd.speed
    out_k, paths = run_mode(w, crowd, "ref-kdtree", ticks)
    out_e, paths_e = run_mode(w, crowd, "exact-knn", ticks)
    assert np.array_equal(paths[0], paths_e[0]) and np.array_equal(paths[1], paths_e[1])
    data["crowd/path_off"], data["crowd/path_xy"] = paths
    data.update(out_k)
    data.update(out_e)
    # static probes (any mode: no agents involved)
    r = RefSim(w, 8, STEP, "exact-knn")
    pts = probe_points(w, rng, 1600)
    data["probe/xy"] = pts
    data["probe/cell"] = r.query_cells(pts)
    ok, rxy, redge = r.retract(pts)
    data["probe/retract_ok"], data["probe/retract_xy"], data["probe/retract_edge"] = ok, rxy, redge
    # planner probes: start/goal pairs -> the reference planner's polyline (for the host planner's parity)
    m = min(crowd.n, 64)
    plens, ppts = [0], []
    for i in range(m):
        p = r.plan_path(crowd.pos[i], crowd.goal[i], float(crowd.radius[i]))
        p = np.zeros((0, 2), np.float32) if p is None else p
        ppts.append(p)
        plens.append(plens[-1] + len(p))
    data["plan/off"] = np.array(plens, np.int32)
    data["plan/xy"] = np.concatenate(ppts) if ppts else np.zeros((0, 2), np.float32)
    r.close()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    ek, ee = data["ref-kdtree/events"], data["exact-knn/events"]
    print(f"{name}: {crowd.n} agents, {ticks} ticks, events kd={len(ek)} exact={len(ee)}, {os.path.getsize(path) / 1e6:.2f} MB")


def turned(crowd, angle):
    """The crowd of a world that is then turned by `angle` (host.World.rotated): same rotation, rounded to float once."""
    c, s_ = np.cos(angle), np.sin(angle)
    R = np.array([[c, -s_], [s_, c]])
    rot = lambda a: (a.astype(np.float64) @ R.T).astype(np.float32)
    return S.Crowd(rot(crowd.pos), rot(crowd.goal), crowd.radius.copy(), crowd.speed.copy())


def main(only=None):
    if only:
        return main_new(only)
    # (1) the 5k-config world, thinned: open streets, few obstacle interactions
    w1 = S.world_c1()
    make("c1_small", w1, S.crowd_c1(w1, n=320, seed=11), 96, 101)
    # (2) narrow streets, mixed radii and speeds, many obstacle constraints (10+ segments in range)
    w2 = lattice_world([16, 14, 18, 15, 17], [40, 36, 44], 8.0, 0.0, 0.0)
    make("c2_small", w2, S.sample_crowd(w2, 320, 12, radius=(0.2, 0.4), speed=(1.0, 1.6), min_goal_dist=40.0), 96, 102)
    # (3) jam: two opposing groups in one corridor world -> collisions, LP failures, LP3D
    w3 = lattice_world([30, 30], [12, 12, 12], 6.0, 0.0, 0.0)
    c3 = S.sample_crowd(w3, 360, 13, radius=(0.3, 0.3), speed=(1.4, 1.4), min_goal_dist=25.0, wall_margin=0.05)
    make("jam_small", w3, c3, 160, 103)
    main_new(None)


def main_new(only):
    """Round 2: obstacle geometry the lattice scenes never had - oblique segments and concave vertices
    (ORCA.cpp:146-239, ECMDataTypes.cpp:52-59, UtilityFunctions.cpp:54-86 on non-axis-aligned cell edges)."""
    # (4) the narrow-street world turned by an angle with no special relation to the axes: every ECM cell edge and
    #     every obstacle segment is oblique, coordinates carry rounding of the rotation
    if only in (None, "oblique_small"):
        w4 = lattice_world([16, 14, 18, 15, 17], [40, 36, 44], 8.0, -40.0, -60.0)
        c4 = S.sample_crowd(w4, 280, 14, radius=(0.2, 0.4), speed=(1.0, 1.6), min_goal_dist=40.0)
        a4 = 0.6154797
        make("oblique_small", w4.rotated(a4), turned(c4, a4), 72, 104)
    # (5) recessed blocks (U and L shaped obstacle polygons => concave vertices) in the jam world, turned as well: crowd
    #     pressure pushes agents along and into the recesses, so the !isConvex legs, the foreign-leg tests against
    #     non-rectangular neighbours and LP3D all run on oblique, concave input
    # (6) a scene from the Boost-free ECM generator (host.polygon_world): oblique triangle / rectangle / pentagon and an
    #     L-shaped obstacle with a CONSISTENT medial axis (the recesses of (5) leave the lattice ECM as it was)
    if only in (None, "yard_small"):
        from ecmgenerator_b200.host import polygon_world

        w6 = polygon_world(*S.scene_polygons("yard"))
        c6 = S.crowd_in_scene(w6, 300, 16, radius=0.3, speed=1.4, min_goal_dist=45.0, spacing=1.6)
        make("yard_small", w6, c6, 140, 106)
    if only in (None, "concave_small"):
        w5 = lattice_world([30, 30], [12, 12, 12], 6.0, -33.0, -24.0).with_recessed_obstacles(7, depth=(0.8, 2.0))
        c5 = S.sample_crowd(lattice_world([30, 30], [12, 12, 12], 6.0, -33.0, -24.0), 360, 15, radius=(0.3, 0.3), speed=(1.4, 1.4),
                            min_goal_dist=25.0, wall_margin=0.05)
        a5 = -0.2449787
        make("concave_small", w5.rotated(a5), turned(c5, a5), 160, 105)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
