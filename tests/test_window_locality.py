"""CPU: the window check the full-size GPU test relies on (tests/util.py check_window_against_oracle), exercised with
the emulated kernels on a crowd small enough for the CPU: the C oracle stepping only `window + margin` reproduces what
the kernels computed for the window's agents inside the whole crowd."""
import numpy as np

from ecmgenerator_b200 import host
from ecmgenerator_b200 import scenarios as S
from tests.test_hostdev_kernels import EmuDevice, emu  # noqa: F401
from tests.util import check_window_against_oracle


class _Scene:
    def __init__(self, world, crowd, off, pxy, step):
        self.world, self.crowd, self.path_off, self.path_xy, self.step, self.n = world, crowd, off, pxy, step, crowd.n


def test_window_of_an_emulated_crowd_matches_the_oracle_on_the_sub_crowd(emu):
    w = S.world_c1()
    c = S.crowd_c1(w, n=1500, seed=5)
    off, pxy, _ = host.plan_paths(w, c.pos, c.goal, c.radius, threads=0)
    keep = np.flatnonzero(np.diff(off) >= 2)
    c = c.take(keep)
    off, pxy, _ = host.plan_paths(w, c.pos, c.goal, c.radius, threads=0)
    assert (np.diff(off) >= 2).all()
    g = _Scene(w, c, off, pxy, float(S.DT))
    d = EmuDevice(emu, g, 4.0)
    for _ in range(25):  # let the crowd pick up speed
        assert emu.emu_tick(d.h) == 0
    before = d.state()
    assert emu.emu_tick(d.h) == 0
    after = d.state()
    x0, y0, x1, y1 = (float(v) for v in w.bbox)
    stats = check_window_against_oracle(w, g.step, before, after, c.radius, c.speed, off, pxy, (x0 + 45, y0 + 45, x1 - 45, y1 - 45),
                                        margin=30.0, vel_tol=0.0, label="emulated c1")
    print(stats)
    assert stats["velocity_rows_bit_identical"] == 1.0 and stats["moving"] > 0.5
    d.close()
