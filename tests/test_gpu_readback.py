"""GPU tests of the steps right after the hot path: the asynchronous readback (msim_snapshot_*: device image packed on
the tick stream, copied into two alternating pinned host buffers while later ticks run), the drop-in sim::Simulator with a consumer that
takes the entity buffer every few milliseconds like the UI does per frame, and the next-waypoint pass riding beside the collision query as a
strided background grid.  Same bar as tests/test_gpu_parity.py: bit-exact against the oracle on every field.  (Round 1 kept these behind
MSIM_TEST_UNVERIFIED because they had been written without a GPU; they ran green on hardware at the start of round 2.)"""
import os

import numpy as np
import pytest

from conftest import assert_entities_equal, oracle_dispatch, oracle_map, to_oracle_entities

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("ctas", ["0", "1", "2"])
def test_background_pass_b_beside_the_query(ctas):
    """MSIM_ARRIVE_BESIDE_CTAS: pass B rides beside the query as a strided grid of k CTAs per SM (148 k x 8192 entities per stride, so
    1.5 M entities make it stride; 0 = full grid).  Six sim ticks on the bench's map against the oracle, in a fresh process (the knob is read once)."""
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
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, MSIM_ARRIVE_BESIDE_CTAS=ctas), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "background pass B ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


