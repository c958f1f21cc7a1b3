"""The WHOLE library on the CPU: api.cu (handle state machine, readback, re-sort, snapshots), every kernel file and the host map code are
compiled for the host SIMT emulator (tests/cuda_emu/build_emu_lib.py -> libmsim_emu.so, same C ABI) and driven through the ordinary ctypes
binding, loaded as a second copy of the package so that the product module and its real library are not touched.

What runs here: sim ticks in every rebuild / flag mode, the asynchronous snapshots, the display quadtree, the C++ drop-in with its headless
runner, and a sample of the required GPU parity tests at emulator-friendly sizes.  Same bar: bit-exact against
the oracle.  This is a development loop and a logic check; it is not hardware and proves nothing about speed, timing-dependent behaviour or
the real memory system.  TEST INFRASTRUCTURE ONLY: the product never loads this library (its binding has no switch for it)."""
import os

import numpy as np
import pytest

import test_gpu_parity as parity
import test_gpu_readback as readback
from conftest import ROOT, assert_entities_equal, oracle_dispatch, oracle_map, to_oracle_entities

EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")


@pytest.fixture(scope="module")
def emu_msim():
    from conftest import load_library_under_emulator

    return load_library_under_emulator()


@pytest.fixture(scope="module")
def emu_city(emu_msim):
    return emu_msim.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7)  # tests/conftest.py::small_city, as the emulated copy's own Map type


@pytest.fixture(scope="module")
def emu_test_map(emu_msim):
    return emu_msim.Map.load_json(os.path.join(ROOT, "tests", "golden", "test_map.json"))


# The emulator runs one OS thread per CUDA thread: the default CPU suite keeps one case per kernel family (default rebuild = counting sort with
# re-sorts, onesweep) plus the snapshot, quadtree and grid-change cases; MSIM_TEST_SLOW=1 adds the remaining flag combinations, the enqueued-ticks
# state-machine test and the C++ drop-in Simulator with its headless runner (all of them run on the GPU in the required suite)
SLOW = os.environ.get("MSIM_TEST_SLOW") == "1"
slow = pytest.mark.skipif(not SLOW, reason="emulator: long case, MSIM_TEST_SLOW=1 runs it")


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("flag_names", [(), pytest.param(("FLAG_NO_REORDER",), marks=slow), pytest.param(("FLAG_NO_REORDER", "FLAG_SORT_COUNTING"), marks=slow),
                                        ("FLAG_SORT_ONESWEEP",), pytest.param(("FLAG_NO_PAIR_COUNT",), marks=slow)])
def test_sim_ticks_through_the_c_abi(emu_msim, orc, emu_city, flag_names, monkeypatch):
    """Blocking dispatches, every rebuild mode of the neighbour structure, a cell re-sort every 3 collision passes, readback at several points."""
    monkeypatch.setenv("MSIM_REORDER_EVERY", "3")
    flags = 0
    for name in flag_names:
        flags |= getattr(emu_msim, name)
    n = 2500
    ents = emu_city.init_entities(n, seed=12)
    om = oracle_map(orc, emu_city)
    want = to_oracle_entities(orc, ents)
    with emu_msim.Simulation(emu_city, ents, radius=10.0, flags=flags) as sim:
        for tick in range(2, 2 + 2 * 9):
            sim.dispatch(tick)
            pairs = oracle_dispatch(orc, want, om, 10.0, tick)
            if tick % 2 == 1 and "FLAG_NO_PAIR_COUNT" not in flag_names:
                assert sim.stats()["last_pair_count"] == pairs, f"tick {tick}"
            if tick in (3, 8, 9, 19):
                assert_entities_equal(sim.read_entities(), want, what=f"{flag_names}: tick {tick}")
        assert (sim.read_collision_flags() == orc.collision_flags(want)).all()
        assert (sim.read_positions() == want["pos"]).all()


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("mode", ["default", pytest.param("counting", marks=slow)])
def test_radius_back_and_forth(emu_msim, orc, emu_city, mode):
    """tests/test_gpu_parity.py::test_radius_back_and_forth_reuses_the_counter_table through the emulated library (the sequence the round-1
    advisor reproduced a crash with)."""
    parity.test_radius_back_and_forth_reuses_the_counter_table(emu_msim, orc, emu_city, mode)


@pytest.mark.timeout(1800)
def test_snapshots(emu_msim, orc, emu_city, emu_test_map):
    """msim_snapshot_begin / poll / end through the real handle code (the emulated runtime copies synchronously: buffer rotation, state-at-begin
    semantics and the error paths are what is checked)."""
    n = 1500
    ents = emu_city.init_entities(n, seed=21)
    om = oracle_map(orc, emu_city)
    want = to_oracle_entities(orc, ents)
    with emu_msim.Simulation(emu_city, ents, radius=10.0) as sim:
        sim.dispatch(2)
        oracle_dispatch(orc, want, om, 10.0, 2)
        tick = 3

        def advance(k):
            nonlocal tick
            sim.enqueue_ticks(k, True)
            for _ in range(k):
                oracle_dispatch(orc, want, om, 10.0, tick + 1)
                oracle_dispatch(orc, want, om, 10.0, tick + 2)
                tick += 2

        advance(3)
        sim.snapshot_begin()
        want_a = want.copy()
        advance(4)
        snap_a = sim.snapshot_end(copy=False)
        assert_entities_equal(snap_a, want_a, what="snapshot A = state at begin")
        sim.snapshot_begin()
        want_b = want.copy()
        advance(1)
        assert sim.snapshot_ready()
        snap_b = sim.snapshot_end(copy=False)
        assert_entities_equal(snap_b, want_b, what="snapshot B")
        assert_entities_equal(snap_a, want_a, what="snapshot A is still intact (other pinned buffer)")
        assert_entities_equal(sim.read_entities(), want, what="blocking readback afterwards")
    readback.test_snapshot_argument_errors(emu_msim, emu_test_map)


@pytest.mark.timeout(1800)
def test_required_parity_tests_at_emulator_size(emu_msim, orc, emu_test_map, emu_city):
    """Two of the required GPU tests as they are (they are small enough): 300 enqueued move passes, and the CUDA path against the reference's
    compiled shader (250 dispatches of 30 k entities would take too long here: the same test body at 2 k)."""
    parity.test_enqueue_ticks_matches_dispatch(emu_msim, orc, emu_test_map)
    if orc.ref_shader_available():
        ents = emu_city.init_entities(2000, seed=17)
        om = oracle_map(orc, emu_city)
        want = to_oracle_entities(orc, ents)
        with emu_msim.Simulation(emu_city, ents, flags=emu_msim.FLAG_NO_COLLISIONS) as sim:
            for step in range(1 + 60):
                sim.dispatch(2 + 2 * step)
                orc.ref_shader_move_pass(want, om)
            assert_entities_equal(sim.read_entities(), want, what="emulated library vs compiled shader")


@pytest.mark.timeout(1800)
@slow
def test_asynchronous_ticks_interleaved_with_host_operations(emu_msim, orc, emu_city, monkeypatch):
    """The required GPU test as it is, at emulator size: the handle's state machine around enqueued ticks (joins, re-upload, grid change)."""
    parity.test_asynchronous_ticks_interleaved_with_host_operations(emu_msim, orc, emu_city, monkeypatch, n=1201)


def test_display_quadtree_through_the_c_abi(emu_msim, orc, emu_city):
    """msim_read_quadtree_nodes (device histogram with dynamic shared memory + host builder) equals its host twin."""
    ents = emu_city.init_entities(3000, seed=3)
    with emu_msim.Simulation(emu_city, ents, radius=10.0) as sim:
        for tick in range(2, 18):
            sim.dispatch(tick)
        got = sim.read_quadtree_nodes()
        pos = sim.read_positions()
    want = emu_msim.quadtree_from_positions(pos, emu_city.width, emu_city.height, 8, 10)
    assert got.shape == want.shape and got.tobytes() == want.tobytes()


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("mode", [pytest.param("blocking", marks=slow), pytest.param("async", marks=slow)])
def test_cpp_simulator_and_headless_runner_on_the_emulated_library(emu_msim, orc, emu_city, tmp_path, mode):
    """The drop-in sim::Simulator (worker thread, hand-off protocol, CSV) and the headless runner, linked against libmsim_emu.so: a consumer takes
    the entity buffer every 2 ms like the UI does per frame; blocking readback and the asynchronous snapshot path must give the same simulation."""
    import re
    import subprocess

    runner = os.path.join(EMU_DIR, "msim_headless_emu")
    sim_dir = os.path.join(ROOT, "movement-sim_b200", "sim")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", runner, os.path.join(sim_dir, "main_headless.cpp"), os.path.join(sim_dir, "Simulator.cpp"),
                        "-L" + EMU_DIR, "-lmsim_emu", "-Wl,-rpath," + EMU_DIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    path = str(tmp_path / "city.msimmap")
    emu_city.save_binary(path)
    n, ticks = 1500, 16  # (emulator speed: every tick is read back in full by the blocking mode)
    dump = str(tmp_path / "entities.bin")
    cmd = [runner, "--headless", "--quiet", "--map", path, "--entities", str(n), "--seed", "7", "--ticks", str(ticks), "--consume-entities", "--dump", dump,
           "--csv", str(tmp_path / "t.csv")] + (["--async-readback"] if mode == "async" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-1500:]
    assert int(re.search(r"entity_frames=(\d+)", r.stdout).group(1)) >= 2
    want = to_oracle_entities(orc, emu_city.init_entities(n, seed=7))
    om = oracle_map(orc, emu_city)
    for tick in range(2, 2 + 2 * ticks):
        oracle_dispatch(orc, want, om, 10.0, tick)
    assert_entities_equal(np.fromfile(dump, dtype=emu_msim.ENTITY_DTYPE), want, what=f"C++ Simulator on the emulated library, {mode} readback")
    rows = open(tmp_path / "t.csv").read().strip().splitlines()
    assert len(rows) == ticks and rows[-1].split(";")[1] == str(ticks)
