"""The C++ drop-in sim::Simulator (movement-sim_b200/sim/, public API of
/root/reference/src/sim/Simulator.hpp:92-119) driven through its headless runner, against the oracle.
K calls of Simulator::sim_tick issue dispatch ticks 2..2K+1: the first only initialises
(/root/reference/src/sim/Simulator.cpp:101,220-235), so K ticks = K-1 move passes + K collision passes."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, assert_entities_equal, oracle_dispatch, oracle_map, to_oracle_entities

pytestmark = pytest.mark.gpu
RUNNER = os.path.join(ROOT, "movement-sim_b200", "msim_headless")


def run_headless(tmp_path, *args):
    dump = str(tmp_path / "entities.bin")
    csv = str(tmp_path / "ticks.csv")
    cmd = [RUNNER, "--headless", "--quiet", "--dump", dump, "--csv", csv, *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    return dump, csv, r.stdout


def test_config1_through_the_cpp_simulator(msim, orc, test_map, tmp_path):
    """BASELINE config 1 end to end through the drop-in: test_map.json, 10 k entities, seed 42,
    1000 move passes (= 1001 sim ticks), collisions off."""
    dump, csv, out = run_headless(tmp_path, "--map", os.path.join(GOLDEN, "test_map.json"), "--entities", "10000", "--seed", "42",
                                  "--ticks", "1001", "--no-collisions")
    got = np.fromfile(dump, dtype=msim.ENTITY_DTYPE)
    want = to_oracle_entities(orc, test_map.init_entities(10_000, seed=42))
    om = oracle_map(orc, test_map)
    for _ in range(1 + 1000):
        orc.move_pass(want, om)
    assert_entities_equal(got, want, what="C++ Simulator, config 1")
    rows = open(csv).read().strip().splitlines()
    assert len(rows) == 1001 and rows[-1].split(";")[1] == "1001"  # time;tick/2;secUpdate;secCollision;secAll
    assert "ticks=1001" in out


def test_collisions_through_the_cpp_simulator(msim, orc, tmp_path):
    path = str(tmp_path / "city.json")
    import ctypes as C

    L = msim.lib()
    h = C.c_void_p()
    assert L.msim_map_generate_city(1500.0, 1000.0, 35.0, 0.3, 0.12, 3, C.byref(h)) == 0
    assert L.msim_map_save_json(h, path.encode()) == 0
    L.msim_map_free(h)
    m = msim.Map.load_json(path)
    dump, csv, out = run_headless(tmp_path, "--map", path, "--entities", "30000", "--seed", "7", "--ticks", "25")
    got = np.fromfile(dump, dtype=msim.ENTITY_DTYPE)
    want = to_oracle_entities(orc, m.init_entities(30_000, seed=7))
    om = oracle_map(orc, m)
    for tick in range(2, 2 + 2 * 25):
        oracle_dispatch(orc, want, om, 10.0, tick)
    assert_entities_equal(got, want, what="C++ Simulator, collisions on")


def test_missing_map_fails_like_the_reference(tmp_path):
    r = subprocess.run([RUNNER, "--map", str(tmp_path / "missing.json"), "--ticks", "1", "--quiet", "--csv", str(tmp_path / "x.csv")],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "could not be loaded" in r.stderr
