"""GPU parity tests of code paths that were written AFTER the round's GPU budget was spent: they compile for sm_100a,
their host-side state machine is reviewed, but they have never run on hardware.  They are opt-in in the library
(a flag or an environment variable; the defaults are the measured, parity-green configuration) and these tests are
skipped unless MSIM_TEST_UNVERIFIED=1, so that the required `-m gpu` suite only contains verified paths.  First thing
to run on the next GPU box:

    MSIM_TEST_UNVERIFIED=1 python -m pytest tests/test_zz_gpu_unverified.py -m gpu -x -q

Same bar as tests/test_gpu_parity.py: bit-exact against the oracle on every field."""
import os

import numpy as np
import pytest

from conftest import assert_entities_equal, oracle_dispatch, oracle_map, to_oracle_entities

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MSIM_TEST_UNVERIFIED") != "1", reason="paths not yet run on hardware: set MSIM_TEST_UNVERIFIED=1")]


# ---- MSIM_FLAG_FUSED_ARRIVE: pass B of a move served by the next move kernel ------------------------------------
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1023, 4097, 50_001])
def test_fused_arrive_collisions_off(msim, orc, test_map, n):
    """Asynchronous move passes with the pending next-waypoint pass consumed by the following move kernel; test_map has
    63.64 m roads, so every entity arrives every ~46 passes and 4-way junctions draw from the RNG."""
    ents = test_map.init_entities(n, seed=7 + n)
    omap = oracle_map(orc, test_map)
    want = to_oracle_entities(orc, ents)
    orc.move_pass(want, omap)  # init dispatch
    with msim.Simulation(test_map, ents, flags=msim.FLAG_NO_COLLISIONS | msim.FLAG_FUSED_ARRIVE) as sim:
        sim.dispatch(2)
        done = 0
        for chunk in (1, 2, 45, 47, 200, 5):  # a readback between chunks completes the pending pass with the stand-alone kernel
            sim.enqueue_ticks(chunk, False)
            for _ in range(chunk):
                orc.move_pass(want, omap)
            done += chunk
            assert_entities_equal(sim.read_entities(), want, what=f"n={n}: {done} fused move passes")
        launches = sim.stats()["kernel_launches"]
        sim.enqueue_ticks(100, False)
        # 100 move kernels with pass B riding inside them, plus ONE stand-alone pass B for the last move (msim_get_stats synchronises)
        assert sim.stats()["kernel_launches"] - launches == 101
        for _ in range(100):
            orc.move_pass(want, omap)
        assert_entities_equal(sim.read_entities(), want, what=f"n={n}: final")


@pytest.mark.parametrize("mode", ["default", "onesweep"])
def test_fused_arrive_with_collisions(msim, orc, small_city, mode, monkeypatch):
    """Full sim ticks: the pending pass B survives the collision pass (which needs positions only) and is consumed by the
    next move kernel; the periodic cell re-sort permutes target / road / rng and therefore completes it first."""
    monkeypatch.setenv("MSIM_REORDER_EVERY", "5")
    n = 60_000
    flags = msim.FLAG_FUSED_ARRIVE | (msim.FLAG_NO_REORDER if mode == "onesweep" else 0)
    ents = small_city.init_entities(n, seed=11)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(small_city, ents, radius=10.0, flags=flags) as sim:
        sim.dispatch(2)
        oracle_dispatch(orc, want, omap, 10.0, 2)
        tick = 3
        for chunk in (1, 3, 8, 20):
            sim.enqueue_ticks(chunk, True)
            sim.sync()
            pairs = 0
            for _ in range(chunk):
                # enqueue_ticks = (move, collide) per sim tick: even dispatch then odd dispatch
                oracle_dispatch(orc, want, omap, 10.0, tick + 1)
                pairs = oracle_dispatch(orc, want, omap, 10.0, tick + 2)
                tick += 2
            assert sim.stats()["last_pair_count"] == pairs, f"after {tick} dispatches"
            assert_entities_equal(sim.read_entities(), want, what=f"{mode}: tick {tick}")


def test_fused_arrive_blocking_dispatch_is_unchanged(msim, orc, small_city):
    """msim_dispatch synchronises, so every pending pass is completed by the stand-alone kernel: same results, no fusion."""
    n = 20_000
    ents = small_city.init_entities(n, seed=5)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(small_city, ents, radius=10.0, flags=msim.FLAG_FUSED_ARRIVE) as sim:
        for t in range(2, 2 + 2 * 10):
            sim.dispatch(t)
            oracle_dispatch(orc, want, omap, 10.0, t)
        assert_entities_equal(sim.read_entities(), want, what="blocking dispatches")


# ---- launch tuning from the environment (read once per process: run each value in its own process) ---------------
@pytest.mark.parametrize("env", [{"MSIM_MOVE_MIN_BLOCKS": "5"}, {"MSIM_MOVE_MIN_BLOCKS": "6"}, {"MSIM_MOVE_GRID": "occupancy"},
                                 {"MSIM_MOVE_MIN_BLOCKS": "6", "MSIM_MOVE_GRID": "occupancy"}, {"MSIM_SCAN_MIN_BLOCKS": "8"},
                                 {"MSIM_ARRIVE_GRID": "persistent"}, {"MSIM_CSORT_MAX_CELLS_LOG2": "27"}, {"MSIM_QUERY_PAIRED": "1"}, {"MSIM_L2_PERSIST_ROADS": "1"}])
def test_move_tuning_variants_in_subprocess(env):
    """The register-capped instantiations of the move kernel and the occupancy-sized grid: smoke() (bit-exact against the
    oracle over 8 sim ticks) in a fresh process per setting."""
    import subprocess
    import sys

    from conftest import ROOT

    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=dict(os.environ, **env), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


def test_persistent_arrive_grid_at_a_size_that_strides():
    """MSIM_ARRIVE_GRID=persistent only changes the launch above 1184 CTAs x 8192 entities: 10.5 M entities on the test map, 60 move
    passes against the oracle, in a fresh process (the knob is read once)."""
    import subprocess
    import sys

    from conftest import ROOT

    code = (
        "import numpy as np, movement_sim_b200 as M\n"
        "from oracle import oracle as O\n"
        "m = M.Map.load_json('tests/golden/test_map.json')\n"
        "ents = m.init_entities(10_500_000, seed=3)\n"
        "om = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)\n"
        "want = np.ascontiguousarray(ents).view(O.ENTITY_DTYPE).copy()\n"
        "for _ in range(61): O.move_pass(want, om, threads=16)\n"
        "with M.Simulation(m, ents, flags=M.FLAG_NO_COLLISIONS) as sim:\n"
        "    sim.dispatch(2); sim.enqueue_ticks(60, False); got = sim.read_entities()\n"
        "assert got.tobytes() == want.tobytes(), 'mismatch'\n"
        "print('persistent arrive ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, MSIM_ARRIVE_GRID="persistent"), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "persistent arrive ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


# ---- msim_snapshot_*: asynchronous readback into pinned double buffers ----------------------------------------------
def test_snapshot_is_the_state_at_begin(msim, orc, small_city):
    """A snapshot holds the state as of msim_snapshot_begin even though more ticks are enqueued before it is collected; two
    snapshots alternate between two pinned buffers, so the first stays intact while the second is filled."""
    n = 30_000
    ents = small_city.init_entities(n, seed=21)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(small_city, ents, radius=10.0) as sim:
        sim.dispatch(2)
        oracle_dispatch(orc, want, omap, 10.0, 2)
        tick = 3

        def advance(k):
            nonlocal tick
            sim.enqueue_ticks(k, True)
            for _ in range(k):
                oracle_dispatch(orc, want, omap, 10.0, tick + 1)
                oracle_dispatch(orc, want, omap, 10.0, tick + 2)
                tick += 2

        advance(5)
        sim.snapshot_begin()
        want_a = want.copy()
        advance(7)  # runs while the copy engine drains the image
        snap_a = sim.snapshot_end(copy=False)
        assert_entities_equal(snap_a, want_a, what="snapshot A = state at begin")
        sim.snapshot_begin()
        want_b = want.copy()
        advance(2)
        while not sim.snapshot_ready():
            pass
        snap_b = sim.snapshot_end(copy=False)
        assert_entities_equal(snap_b, want_b, what="snapshot B")
        assert_entities_equal(snap_a, want_a, what="snapshot A after B was taken (other pinned buffer)")
        assert_entities_equal(sim.read_entities(), want, what="blocking readback afterwards")


def test_snapshot_argument_errors(msim, test_map):
    with msim.Simulation(test_map, test_map.init_entities(100), flags=msim.FLAG_NO_COLLISIONS) as sim:
        with pytest.raises(msim.MsimError) as ei:
            sim.snapshot_end()
        assert ei.value.status == msim.MSIM_ERR_INVALID and "no snapshot" in ei.value.message
        sim.snapshot_begin()
        got = sim.snapshot_end()
        assert got.shape[0] == 100 and np.array_equal(got["road_index"], sim.read_entities()["road_index"])


# ---- the drop-in Simulator with the asynchronous readback (MSIM_ASYNC_READBACK / --async-readback) -------------------
@pytest.mark.parametrize("mode", ["blocking", "async"])
def test_cpp_simulator_with_a_consumer(msim, orc, small_city, tmp_path, mode):
    """A consumer takes the entity buffer every 2 ms, like the UI does per frame; the simulation result must not depend on
    how the readback is done, and frames must actually flow."""
    import re
    import subprocess

    from conftest import ROOT

    runner = os.path.join(ROOT, "movement-sim_b200", "msim_headless")
    path = str(tmp_path / "city.msimmap")
    small_city.save_binary(path)
    dump = str(tmp_path / "entities.bin")
    cmd = [runner, "--headless", "--quiet", "--map", path, "--entities", "40000", "--seed", "7", "--ticks", "400", "--consume-entities", "--dump", dump,
           "--csv", str(tmp_path / "t.csv")] + (["--async-readback"] if mode == "async" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    frames = int(re.search(r"entity_frames=(\d+)", r.stdout).group(1))
    assert frames >= 2, r.stdout
    got = np.fromfile(dump, dtype=msim.ENTITY_DTYPE)
    want = to_oracle_entities(orc, small_city.init_entities(40_000, seed=7))
    om = oracle_map(orc, small_city)
    for tick in range(2, 2 + 2 * 400):
        oracle_dispatch(orc, want, om, 10.0, tick)
    assert_entities_equal(got, want, what=f"C++ Simulator with a consumer, {mode} readback")


def test_background_pass_b_beside_the_query():
    """MSIM_ARRIVE_BESIDE_CTAS=1: pass B rides beside the query as a strided grid of one CTA per SM (148 x 8192 entities per stride, so
    1.5 M entities make it stride).  Six sim ticks on the bench's map against the oracle, in a fresh process (the knob is read once)."""
    import subprocess
    import sys

    from conftest import ROOT

    code = (
        "import numpy as np, movement_sim_b200 as M\n"
        "from oracle import oracle as O\n"
        "m = M.Map.city()\n"
        "ents = m.init_entities(1_500_000, seed=4)\n"
        "om = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)\n"
        "want = np.ascontiguousarray(ents).view(O.ENTITY_DTYPE).copy()\n"
        "pairs = 0\n"
        "for t in range(2, 2 + 2 * 7):\n"
        "    if t % 2 == 0: O.move_pass(want, om, threads=16)\n"
        "    else: pairs = O.collide_pass(want, om.world_w, om.world_h, 10.0, threads=16)\n"
        "with M.Simulation(m, ents, radius=10.0) as sim:\n"
        "    sim.dispatch(2); sim.dispatch(3); sim.enqueue_ticks(6, True); sim.sync()\n"
        "    assert sim.stats()['last_pair_count'] == pairs, (sim.stats()['last_pair_count'], pairs)\n"
        "    got = sim.read_entities()\n"
        "assert got.tobytes() == want.tobytes(), 'mismatch'\n"
        "print('background pass B ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, MSIM_ARRIVE_BESIDE_CTAS="1"), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "background pass B ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


@pytest.mark.parametrize("env", [{"MSIM_QUERY_PAIRED": "1"},
                                 {"MSIM_MOVE_MIN_BLOCKS": "6", "MSIM_MOVE_GRID": "occupancy", "MSIM_SCAN_MIN_BLOCKS": "8", "MSIM_ARRIVE_GRID": "persistent",
                                  "MSIM_ARRIVE_BESIDE_CTAS": "2", "MSIM_QUERY_PAIRED": "1"}])
def test_whole_parity_file_under_the_knobs(env):
    """Every test of tests/test_gpu_parity.py (ragged sizes, city population tick by tick with pair counts, point clouds with duplicates and
    exact-radius pairs, stacked start, all rebuild modes) in a fresh process with the launch knobs set."""
    import subprocess
    import sys

    from conftest import ROOT

    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-x", "-q"], cwd=ROOT, env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-3000:]
