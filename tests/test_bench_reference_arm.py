"""CPU: `bench.py --impl reference` (the reference's own Simulator::Update timed on host cores) prints the contract's
JSON line, bounds its sample by the time budget and labels the all-core replica farm as an upper bound."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1_5k", "--steps", "3",
                        "--warmup", "1", "--cpu-budget", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] and line["unit"] == "agent-updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 3 and line["vs_baseline"] is None
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["cores"] == 1 and cb["kind"] in ("reference", "port") and cb["value"] == line["value"]
    assert "agents nearest the crowd centroid" in cb["sample"]
    if (os.cpu_count() or 1) > 1:
        farm = cb["replica_farm"]
        assert farm["replicas"] == min(os.cpu_count(), 128) and farm["value"] > 0 and "upper bound" in farm["note"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--config", "c1_5k"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
