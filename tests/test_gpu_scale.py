"""Full-size checks (BASELINE.json sizes) through size-independent properties plus an oracle comparison
on the largest size the multi-threaded oracle still finishes in seconds."""
import numpy as np
import pytest

from conftest import ORACLE_THREADS, assert_entities_equal, oracle_dispatch, oracle_map, to_oracle_entities

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def munich(msim):
    return msim.Map.city()  # the bench's Munich stand-in: 29007.46 x 16463.77 m, 701 590 roads


def test_munich_1m_vs_oracle(msim, orc, munich):
    """BASELINE config 2 size (1 M entities) with collisions on as well: every field after 20 sim ticks."""
    n = 1_000_000
    ents = munich.init_entities(n, seed=42)
    om = oracle_map(orc, munich)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(munich, ents, radius=10.0) as sim:
        for t in range(2, 2 + 2 * 20):
            sim.dispatch(t)
            p = oracle_dispatch(orc, want, om, 10.0, t)
            if t % 2 == 1:
                assert sim.stats()["last_pair_count"] == p
        assert_entities_equal(sim.read_entities(), want, what="munich 1M")


def test_munich_10m_properties(msim, orc, munich):
    """BASELINE config 3 size on one GPU.  Properties that hold at any size:
    * the two neighbour-structure rebuilds (cell-ordered storage + counting sort / upload order + onesweep
      radix sort) agree on every flag and on the pair count, and with the flags-only query;
    * flagged entities = entities with a partner: sum over blue == 'last_flagged_count';
    * a 200 k-entity sample of the 10 M state equals an oracle continuation of that sample for the
      movement fields (movement never reads another entity);
    * positions stay inside the world and every road index is valid."""
    n = 10_000_000
    ents = munich.init_entities(n, seed=42)
    results = []
    for flags in (0, msim.FLAG_NO_REORDER, msim.FLAG_NO_PAIR_COUNT):
        with msim.Simulation(munich, ents, radius=10.0, flags=flags) as sim:
            sim.dispatch(2)
            sim.enqueue_ticks(12, True)
            sim.sync()
            st = sim.stats()
            got = sim.read_entities()
            results.append((st, got))
    (st_a, a), (st_b, b), (st_c, c) = results
    assert a.tobytes() == b.tobytes() == c.tobytes()
    assert st_a["last_pair_count"] == st_b["last_pair_count"] > 0
    assert st_a["last_flagged_count"] == st_b["last_flagged_count"] == st_c["last_flagged_count"]
    blue = (a["color"] == np.array([0, 0, 1, 1], dtype=np.float32)).all(axis=1)
    green = (a["color"] == np.array([0, 1, 0, 1], dtype=np.float32)).all(axis=1)
    assert (blue | green).all() and int(blue.sum()) == st_a["last_flagged_count"]
    assert (a["pos"] >= 0).all() and (a["pos"][:, 0] <= munich.width).all() and (a["pos"][:, 1] <= munich.height).all()
    assert (a["road_index"] < munich.roads.shape[0]).all() and (a["initialized"] == 1).all()
    # oracle continuation of a strided sample (movement fields only)
    sample = slice(0, n, 50)
    want = to_oracle_entities(orc, ents[sample])
    om = oracle_map(orc, munich)
    for _ in range(1 + 12):
        orc.move_pass(want, om, threads=ORACLE_THREADS)
    assert_entities_equal(a[sample], want, fields=["pos", "target", "rng", "road", "dir", "initialized"], what="10M sample")


def test_munich_10m_move_only_long_run(msim, orc, munich):
    """Collisions off, 10 M entities, 300 move passes via the asynchronous tick queue; a strided sample is
    compared with the oracle bit for bit."""
    n = 10_000_000
    ents = munich.init_entities(n, seed=7)
    with msim.Simulation(munich, ents, flags=msim.FLAG_NO_COLLISIONS) as sim:
        sim.dispatch(2)
        sim.enqueue_ticks(300, False)
        sim.sync()
        got = sim.read_entities()
    sample = slice(3, n, 97)
    want = to_oracle_entities(orc, ents[sample])
    om = oracle_map(orc, munich)
    for _ in range(1 + 300):
        orc.move_pass(want, om, threads=ORACLE_THREADS)
    assert_entities_equal(got[sample], want, what="10M move-only sample")
