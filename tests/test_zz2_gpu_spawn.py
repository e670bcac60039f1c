"""GPU (-m gpu): ecmgpu_valid_spawn_locations against Simulator::ValidSpawnLocation's scan over
every agent (Simulator.cpp:295-311), restated in float32; tests/test_hostdev_kernels.py pins the kernel on the CPU."""
import numpy as np
import pytest

from ecmgenerator_b200 import gpu
from tests.util import Golden

# a kernel that never returns must not hang the box: the watchdog thread ends the run instead
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


@pytest.mark.parametrize("cell", [0.0, 0.6, 40.0])
def test_valid_spawn_locations_equal_the_reference_scan(cell):
    g = Golden("c2_small")
    sim = gpu.GpuSim(g.world, g.n + 8, g.step, neighbor_cell=cell)
    sim.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    sim.update(12)  # agents under way: the check runs on the CURRENT positions
    sim.destroy_agent(5)
    pos = sim.read(gpu.POS, 0, g.n)
    act = sim.read(gpu.ACTIVE, 0, g.n) > 0
    assert not act[5]
    rng = np.random.default_rng(22)
    x0, y0, x1, y1 = (float(v) for v in g.world.bbox)
    q = np.concatenate([rng.uniform([x0 - 30, y0 - 30], [x1 + 30, y1 + 30], size=(4000, 2)),
                        pos[rng.integers(0, g.n, 2000)] + rng.normal(0, 0.4, size=(2000, 2)), pos[5:6]]).astype(np.float32)
    cl = rng.choice(np.float32([0.25, 0.5, 1.0, 7.5]), size=len(q)).astype(np.float32)
    cl[-1] = 0.2
    out = sim.valid_spawn_locations(q, cl)
    dx = q[:, None, 0] - pos[None, act, 0]
    dy = q[:, None, 1] - pos[None, act, 1]
    want = ~((dx * dx + dy * dy) < (cl * cl)[:, None]).any(axis=1)
    assert np.array_equal(out > 0, want)
    assert 0.2 < want.mean() < 0.95
    # an empty simulator accepts everything
    empty = gpu.GpuSim(g.world, 8, g.step)
    assert empty.valid_spawn_locations(q[:10], 0.3).all()
