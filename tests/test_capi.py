"""CPU: the C-ABI libraries load and export every symbol their headers declare; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT, has_cuda


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def test_gpu_library_exports_every_declared_symbol():
    from ecmgenerator_b200 import gpu

    names = _declared("ecm_b200.h", "ecmgpu_")
    assert len(names) >= 25
    L = C.CDLL(gpu.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(gpu.EXPORTS) == names, "python binding out of sync with include/ecm_b200.h"


def test_host_library_exports_every_declared_symbol():
    from ecmgenerator_b200 import host

    names = _declared("ecm_b200_host.h", "ecmhost_")
    L = host.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_param_struct_layout_matches_header():
    from ecmgenerator_b200 import gpu

    assert C.sizeof(gpu.Params) == 32
    assert gpu.Stats.ticks.offset % 8 == 0


def test_agent_record_layout_matches_header():
    """ecmgpu_agent_rec (ecmgpu_update_io_owned) is five 4-byte fields, slot first, as the numpy dtype says."""
    from ecmgenerator_b200 import gpu

    text = open(os.path.join(ROOT, "include", "ecm_b200.h")).read()
    m = re.search(r"typedef struct ecmgpu_agent_rec \{(.*?)\} ecmgpu_agent_rec;", text, flags=re.S)
    assert m, "ecmgpu_agent_rec not declared"
    fields = [f.strip() for f in m.group(1).replace("\n", " ").split(";") if f.strip()]
    assert fields == ["int32_t slot", "float x, y, vx, vy"]
    assert gpu.AGENT_REC.itemsize == 20 and gpu.AGENT_REC.names == ("slot", "x", "y", "vx", "vy")
    assert [gpu.AGENT_REC.fields[n][1] for n in gpu.AGENT_REC.names] == [0, 4, 8, 12, 16]


def test_tools_and_bench_compile():
    import py_compile

    for rel in ["bench.py", "__graft_entry__.py"] + [os.path.join("tools", f) for f in sorted(os.listdir(os.path.join(ROOT, "tools"))) if f.endswith(".py")]:
        py_compile.compile(os.path.join(ROOT, rel), doraise=True)


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a CUDA device")
def test_no_cpu_fallback_without_gpu():
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200.host import lattice_world

    w = lattice_world([20, 20], [20, 20], 6.0)
    with pytest.raises(gpu.EcmGpuError, match="no CUDA device"):
        gpu.GpuSim(w, 16, 1 / 60)


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a CUDA device")
def test_dropin_fails_loudly_without_gpu():
    """libecmsim.so (the C++17 drop-in Simulator) loads, plans on the host, and refuses to run without a device."""
    from ecmgenerator_b200 import dropin
    from ecmgenerator_b200.host import lattice_world

    w = lattice_world([20, 20], [20, 20], 6.0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        dropin.Simulator(w, 16, 1 / 60)
