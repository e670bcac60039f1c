"""GPU (-m gpu): strip decomposition.  In-process strips on ONE device exercise pack / exchange /
migration / ghosts exactly like the NCCL path (same kernels, peer copies instead of send/recv);
the NCCL transport itself is tested with torchrun when the box has >= 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from ecmgenerator_b200 import gpu
from ecmgenerator_b200 import multigpu as M
from ecmgenerator_b200 import scenarios as S
from ecmgenerator_b200.host import plan_paths
from tests.conftest import ROOT
from tests.util import assert_bits_equal

pytestmark = pytest.mark.gpu


def _crowd(n, seed):
    w = S.world_c1()
    c = S.crowd_c1(w, n=n, seed=seed)
    off, pxy, ok = plan_paths(w, c.pos, c.goal, c.radius)
    assert ok == n
    return w, c, off, pxy


@pytest.mark.parametrize("n_strips", [2, 3, 5])
def test_in_process_strips_match_single_gpu_bitwise(n_strips):
    n, ticks = 6000, 240
    w, c, off, pxy = _crowd(n, 41)
    single = gpu.GpuSim(w, n, float(S.DT), path_pool_points=int(off[-1] * 1.25) + 4096)
    single.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    strips = M.LocalStrips(w, c, off, pxy, n_strips, devices=(0,))
    own0 = M.owner_of(c.pos[:, 0], strips.bounds)
    for t in range(ticks):
        single.update(1)
        strips.update(1)
    single.sync()
    strips.sync()
    pos, owners = strips.gather(gpu.POS)
    act = single.read(gpu.ACTIVE, 0, n)
    assert np.array_equal(owners, act), "every active agent has exactly one owner"
    a = act > 0
    assert_bits_equal(pos[a], single.read(gpu.POS, 0, n)[a], "positions")
    vel, _ = strips.gather(gpu.VEL)
    assert_bits_equal(vel[a], single.read(gpu.VEL, 0, n)[a], "velocities")
    att, _ = strips.gather(gpu.ATTRACTION)
    assert_bits_equal(att[a], single.read(gpu.ATTRACTION, 0, n)[a], "attraction points")
    st = strips.stats()
    assert sum(s["halo_misses"] for s in st) == 0
    # agents did cross strip borders, i.e. migration was exercised
    own1 = M.owner_of(pos[:, 0], strips.bounds)
    moved = int(((own0 != own1) & a).sum())
    print(f"{n_strips} strips: {moved} agents changed owner in {ticks} ticks; halo {strips.halo:.2f}")
    assert moved > 10
    # each strip really only works on its share
    per = [s["n_active"] for s in st]
    assert max(per) < n  # n_active counts owned + ghosts of the last grid build
    strips.close()
    single.close()


def test_rebalance_moves_the_borders_and_keeps_results_bitwise():
    """Strips that start badly balanced (20 % / 30 % / 50 % of the crowd) are re-balanced mid-run: ownership moves
    wholesale, the trajectory does not change by a bit."""
    n, ticks = 6000, 90
    w, c, off, pxy = _crowd(n, 44)
    xs = np.sort(c.pos[:, 0])
    skew = np.array([xs[0] - 1.0, xs[int(0.2 * n)], xs[int(0.5 * n)], xs[-1] + 1.0], np.float32)
    single = gpu.GpuSim(w, n, float(S.DT), path_pool_points=int(off[-1] * 1.25) + 4096)
    single.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    strips = M.LocalStrips(w, c, off, pxy, 3, devices=(0,), bounds=skew)
    single.update(2 * ticks)
    strips.update(ticks)
    before = [int(s.read(gpu.ACTIVE, 0, n).sum()) for s in strips.sims]
    strips.rebalance()
    after = [int(s.read(gpu.ACTIVE, 0, n).sum()) for s in strips.sims]
    assert max(before) > 0.45 * n and max(after) < 0.36 * n, (before, after)
    assert sum(before) == sum(after)
    strips.update(ticks)
    strips.sync()
    pos, owners = strips.gather(gpu.POS)
    act = single.read(gpu.ACTIVE, 0, n)
    assert np.array_equal(owners, act)
    a = act > 0
    assert_bits_equal(pos[a], single.read(gpu.POS, 0, n)[a], "positions")
    assert_bits_equal(strips.gather(gpu.VEL)[0][a], single.read(gpu.VEL, 0, n)[a], "velocities")
    assert_bits_equal(strips.gather(gpu.ATTRACTION)[0][a], single.read(gpu.ATTRACTION, 0, n)[a], "attraction points")
    assert sum(s["halo_misses"] for s in strips.stats()) == 0
    strips.close()
    single.close()


def test_automatic_rebalancing_keeps_results_bitwise():
    """rebalance_every = M: every M ticks the drivers compare the strips' agent counts and move the borders when one
    holds more than 15 % over its share - here twice, starting from 20 / 30 / 50 %."""
    n, ticks = 6000, 120
    w, c, off, pxy = _crowd(n, 45)
    xs = np.sort(c.pos[:, 0])
    skew = np.array([xs[0] - 1.0, xs[int(0.2 * n)], xs[int(0.5 * n)], xs[-1] + 1.0], np.float32)
    single = gpu.GpuSim(w, n, float(S.DT), path_pool_points=int(off[-1] * 1.25) + 4096)
    single.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    strips = M.LocalStrips(w, c, off, pxy, 3, devices=(0,), bounds=skew, rebalance_every=40)
    single.update(ticks)
    strips.update(ticks)
    strips.check_exact()
    assert strips.rebalances >= 1 and max(strips.owned_counts()) < 0.40 * n
    pos, owners = strips.gather(gpu.POS)
    act = single.read(gpu.ACTIVE, 0, n)
    assert np.array_equal(owners, act)
    a = act > 0
    assert_bits_equal(pos[a], single.read(gpu.POS, 0, n)[a], "positions")
    assert_bits_equal(strips.gather(gpu.VEL)[0][a], single.read(gpu.VEL, 0, n)[a], "velocities")
    strips.close()
    single.close()


def test_halo_miss_is_detected_when_the_halo_is_too_small():
    n = 3000
    w, c, off, pxy = _crowd(n, 42)
    strips = M.LocalStrips(w, c, off, pxy, 2, devices=(0,), halo=0.05)
    strips.update(3)
    strips.sync()
    assert sum(s["halo_misses"] for s in strips.stats()) > 0
    with pytest.raises(RuntimeError, match="halo misses"):
        strips.check_exact()  # loud: the strips no longer compute what one GPU computes
    strips.close()


def test_strip_validation_errors():
    n = 500
    w, c, off, pxy = _crowd(n, 43)
    s = gpu.GpuSim(w, n, float(S.DT))
    s.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    with pytest.raises(gpu.EcmGpuError, match="null unique id"):
        s._ck(s.L.ecmgpu_comm_init(s.h, None, 0, 2))
    with pytest.raises(gpu.EcmGpuError, match="comm_init first"):
        s.comm_set_strips(np.array([-1.0, 0.0, 1.0], np.float32), 5.0)
    s2 = gpu.GpuSim(w, n, float(S.DT))
    s2.comm_init_local(0, 1, None, None)
    with pytest.raises(gpu.EcmGpuError, match="ascend"):
        s2.comm_set_strips(np.array([1.0, 0.0], np.float32), 5.0)
    s2.comm_set_strips(np.array([-1000.0, 1000.0], np.float32), 5.0)  # a single strip is fine
    s2.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    s2.update(2)
    s.update(2)
    assert_bits_equal(s2.read(gpu.POS, 0, n), s.read(gpu.POS, 0, n), "one strip == no strips")


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_multi_process_strips_match_single_gpu_bitwise(transport):
    """One process per GPU under torchrun.  transport = peer: entries stored into the neighbour's inbox over
    NVLink (CUDA IPC); nccl: ncclSend/ncclRecv of the staged message."""
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    out = os.path.join(ROOT, "gpurun_out", f"strips_{transport}.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if transport == "peer" else "29518", os.path.join(ROOT, "tests", "nccl_strips_worker.py"), out]
    env = dict(os.environ, ECMGPU_P2P="1" if transport == "peer" else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    print(res)
    assert res["pos_equal"] and res["vel_equal"] and res["halo_misses"] == 0 and res["owners_ok"] and res["moved"] > 10
    assert res["p2p"] == (transport == "peer")
    assert res["io_owned_ok"] and sum(res["io_owned_counts"]) > 0
