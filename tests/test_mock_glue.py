"""CPU: the REAL host glue of the C ABI (csrc/ecmgpu.cu) driven by the GPU test-suite itself, without a GPU.

tests/hostdev/mock_cuda builds csrc/ecmgpu.cu for the host against a synchronous stand-in for the CUDA runtime and runs
its kernels under the SIMT emulator (tests/hostdev/shim/simt.h).  The tests below are a subset of the `-m gpu` tests,
unchanged, pointed at that library through the ECMGPU_LIB hook: launch order, buffer sizing, mode switches (KD-tree
neighbour mode, compact strips), the strip phases with the in-process transport, the planner / spawn /
query entry points, the pipelined host I/O calls and the C++ drop-in Simulator are exercised through the same bindings a
GPU run uses.
The tick runs as a captured graph like on the GPU (the mock records the capture and replays it).  It cannot show timing,
overlap, the peer / NCCL transports or the library sort.
TEST INFRASTRUCTURE: the product library has no CPU path and nothing in the product can load this build."""
import os
import subprocess
import sys

from tests.conftest import ROOT

SUBSET = [
    "tests/test_gpu_parity.py::test_cells_and_retraction_match_reference_golden[c2_small]",
    "tests/test_gpu_parity.py::test_neighbours_bit_exact_vs_reference_golden[0.7-c2_small]",
    "tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[c2_small]",
    "tests/test_gpu_parity.py::test_ties_and_colocated_agents",
    "tests/test_gpu_parity.py::test_arrival_destroy_and_replan_events",
    "tests/test_gpu_parity.py::test_outside_world_and_bin_fallback_paths",
    "tests/test_gpu_parity.py::test_obstacle_lists_match_oracle",
    "tests/test_gpu_parity.py::test_update_io_pipeline_matches_plain_update",
    "tests/test_gpu_parity.py::test_update_io_owned_records_match_plain_update[1]",
    "tests/test_gpu_parity.py::test_update_io_owned_records_match_plain_update[0]",
    # the C++ drop-in Simulator (libecmsim.so): the mock library is preloaded, so its ecmgpu_* symbols are the ones bound
    "tests/test_gpu_simulator_dropin.py::test_spawn_update_getters_match_reference",
    "tests/test_gpu_simulator_dropin.py::test_dropin_matches_c_oracle_with_host_planner",
    "tests/test_gpu_simulator_dropin.py::test_batched_spawn_checks_keep_the_reference_rand_stream_through_rewinds",
    "tests/test_gpu_strips.py::test_halo_miss_is_detected_when_the_halo_is_too_small",
    "tests/test_gpu_strips.py::test_strip_validation_errors",
    "tests/test_zz2_gpu_spawn.py::test_valid_spawn_locations_equal_the_reference_scan[0.0]",
    "tests/test_zz2_gpu_compact.py::test_compact_walk_in_the_graph_tick_with_spawns_and_destroys",
    "tests/test_zz3_gpu_kdtree.py::test_kd_neighbour_lists_equal_the_unmodified_reference[c2_small]",
    "tests/test_zz3_gpu_kdtree.py::test_kd_lockstep_velocities_within_tolerance[c2_small]",
    "tests/test_zz3_gpu_kdtree.py::test_dropin_in_kd_mode_walks_like_the_unmodified_reference",
    "tests/test_zz3_gpu_kdtree.py::test_kd_mode_refuses_strips",
    "tests/test_zz4_gpu_planner.py::test_device_planner_reproduces_the_reference_polylines[c2_small]",
    "tests/test_zz4_gpu_planner.py::test_device_planner_small_pool_is_retried",
]


def test_gpu_suite_subset_through_the_real_c_abi_on_the_mock_runtime():
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostdev", "mock_cuda"))
    import make_mock

    so = make_mock.build()
    env = dict(os.environ, ECMGPU_LIB=so, LD_PRELOAD=so)  # LD_PRELOAD: for libecmsim.so, linked against libecmgpu
    env.pop("ECMGPU_GRAPH", None)  # graphs on, as on the GPU: the mock records a capture as closures and replays them
    for k in ("ECMGPU_COMPACT",):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + SUBSET, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    print(tail)
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    assert f"{len(SUBSET)} passed" in r.stdout or f"{len(SUBSET) - 3} passed, 3 skipped" in r.stdout  # three need oracle/_ref


def _threaded(name, ranks, transport, compact, ticks):
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostdev", "mock_cuda", "threaded_strips.py"), name, str(ranks), transport,
                        str(compact), str(ticks)], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_one_strip_per_thread_over_the_nccl_and_peer_transports():
    """The multi-rank tick as `torchrun` drives it - ecmgpu_update on every rank, NCCL send / recv or the peer transport
    (k_pack storing into the neighbour's inbox, k_exchange_p2p spinning on the sequence number the neighbour writes, the
    tick replayed as a captured graph per inbox generation) - with the ranks as THREADS of one process on the mock
    runtime (mock_nccl.cpp, mock IPC handles).  Bit for bit against the reference's golden trajectory, migrations
    included, with and without the compact walk."""
    for transport, compact, ticks in (("nccl", 0, 60), ("p2p", 1, 160)):
        res = _threaded("jam_small", 3, transport, compact, ticks)
        print(transport, compact, res)
        assert res["owners_ok"] and res["pos_equal"] and res["vel_equal"] and res["halo_misses"] == 0
        assert res["moved"] >= (3 if ticks > 100 else 1) and max(res["owned"]) < 200
