"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Bar: bit-exact on every field (road index, RNG state, target, position, direction, colour
flag, pair count).  The stated tolerance for float positions is therefore 0 ULP.

Configs follow BASELINE.json: config 1 (test_map, 10 k entities, seed 42, 1000 move passes, collisions
off) in full; the Munich-style city at sizes the oracle finishes in seconds; full sizes through
size-independent properties in test_gpu_scale.py."""
import numpy as np
import pytest

from conftest import assert_entities_equal, oracle_dispatch, oracle_map, to_oracle_entities

pytestmark = pytest.mark.gpu


def MODES(msim):
    """flag sets selecting the neighbour-structure rebuild (include/msim.h)"""
    return {"default": 0, "onesweep": msim.FLAG_NO_REORDER, "counting": msim.FLAG_NO_REORDER | msim.FLAG_SORT_COUNTING,
            "reorder+onesweep": msim.FLAG_SORT_ONESWEEP}


def run_oracle(O, ents, omap, radius, ticks):
    """ticks = list of dispatch tick numbers, as Simulator::sim_tick issues them (2,3,4,5,...)."""
    e = to_oracle_entities(O, ents)
    pairs = []
    for t in ticks:
        p = oracle_dispatch(O, e, omap, radius, t)
        if t % 2 == 1:
            pairs.append(p)
    return e, pairs


def test_config1_test_map_10k_1000_moves(msim, orc, test_map):
    """BASELINE config 1: test_map.json, 10 k entities, seed 42, 1000 move passes, collisions off."""
    ents = test_map.init_entities(10_000, seed=42)
    omap = oracle_map(orc, test_map)
    want = to_oracle_entities(orc, ents)
    orc.move_pass(want, omap)  # first dispatch: initialise only
    with msim.Simulation(test_map, ents, flags=msim.FLAG_NO_COLLISIONS) as sim:
        sim.dispatch(2)  # init-only dispatch (random_move.comp:863-867)
        got = sim.read_entities()
        assert_entities_equal(got, want, what="after init dispatch")
        assert sim.read_debug()[0] == 10_000
        for step in range(1000):
            sim.dispatch(4 + 2 * step)
            orc.move_pass(want, omap)
            if step in (0, 1, 45, 46, 47, 499, 999):
                assert_entities_equal(sim.read_entities(), want, what=f"move pass {step + 1}")
        st = sim.stats()
        assert st["move_passes"] == 1000 and st["collide_passes"] == 0


def test_cuda_path_matches_compiled_shader(msim, orc, small_city):
    """The CUDA move path against the reference's OWN movement code: random_move.comp compiled for the CPU
    (oracle/_ref/libref_shader_move.so, see tests/test_oracle_vs_ref_shader.py) instead of the oracle's restatement."""
    if not orc.ref_shader_available():
        pytest.skip("oracle/_ref/libref_shader_move.so not built (needs /root/reference at build time)")
    ents = small_city.init_entities(30_000, seed=17)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(small_city, ents, flags=msim.FLAG_NO_COLLISIONS) as sim:
        for step in range(1 + 250):
            sim.dispatch(2 + 2 * step)
            orc.ref_shader_move_pass(want, omap)
            if step in (0, 1, 2, 30, 100, 250):
                assert_entities_equal(sim.read_entities(), want, what=f"dispatch {step} vs compiled shader")


def test_cuda_path_matches_the_whole_compiled_shader(msim, orc, small_city):
    """Full sim ticks of the CUDA path against the reference's ENTIRE shader compiled for the CPU (oracle/_ref/libref_shader_full.so:
    main() with the lock-based quadtree's insert / update / collision walk): every field incl. the colours, dispatch by dispatch.
    The world handed to both is one metre larger than the map (the shader never terminates for an entity on the map's maximum
    coordinate, tests/test_oracle_vs_ref_shader.py)."""
    if not orc.ref_shader_full_available():
        pytest.skip("oracle/_ref/libref_shader_full.so not built (needs /root/reference at build time)")
    padded = msim.Map(small_city.width + 1.0, small_city.height + 1.0, small_city.roads, small_city.connections)
    ents = padded.init_entities(20_000, seed=23)
    om = oracle_map(orc, padded)
    want = to_oracle_entities(orc, ents)
    ref = orc.RefShaderSim(want, om, radius=10.0)
    with msim.Simulation(padded, ents, radius=10.0) as sim:
        for tick in range(2, 2 + 2 * 40):
            sim.dispatch(tick)
            ref.dispatch(tick)
            if tick in (2, 3, 4, 5, 40, 41, 80, 81):
                assert_entities_equal(sim.read_entities(), want, what=f"dispatch {tick} vs the whole compiled shader")
        assert (sim.read_collision_flags() == orc.collision_flags(want)).all()
        # debugData[1]: the shader over-counts (App. B5); the CUDA path reports the unique pairs, never more
        assert sim.read_debug()[1] <= ref.debug[1]


def test_enqueue_ticks_matches_dispatch(msim, orc, test_map):
    ents = test_map.init_entities(4097, seed=3)
    omap = oracle_map(orc, test_map)
    want = to_oracle_entities(orc, ents)
    for _ in range(1 + 300):
        orc.move_pass(want, omap)
    with msim.Simulation(test_map, ents, flags=msim.FLAG_NO_COLLISIONS) as sim:
        sim.dispatch(2)
        sim.enqueue_ticks(300, False)
        sim.sync()
        assert_entities_equal(sim.read_entities(), want, what="300 enqueued move passes")


@pytest.mark.parametrize("mode", ["default", "onesweep", "counting", "reorder+onesweep"])
@pytest.mark.parametrize("n", [0, 1, 2, 31, 63, 64, 65, 127, 1023, 4096, 4097, 12_345])
def test_ragged_sizes_full_ticks(msim, orc, small_city, n, mode, monkeypatch):
    """Empty and ragged populations through complete sim ticks (move + collide), with every rebuild of
    the neighbour structure: cell-ordered storage + counting sort (default), onesweep radix sort and
    counting sort on upload-ordered storage, onesweep on cell-ordered storage.  The state is re-sorted
    every 3 collision passes here so that several re-sorts happen inside the test."""
    monkeypatch.setenv("MSIM_REORDER_EVERY", "3")
    ents = small_city.init_entities(n, seed=100 + n)
    omap = oracle_map(orc, small_city)
    ticks = list(range(2, 2 + 2 * 12))
    want, want_pairs = run_oracle(orc, ents, omap, 10.0, ticks)
    with msim.Simulation(small_city, ents, radius=10.0, flags=MODES(msim)[mode]) as sim:
        got_pairs = []
        for t in ticks:
            sim.dispatch(t)
            if t % 2 == 1:
                got_pairs.append(sim.stats()["last_pair_count"])
        got = sim.read_entities()
    assert_entities_equal(got, want, what=f"n={n}")
    assert got_pairs == want_pairs


@pytest.mark.parametrize("mode", ["default", "onesweep", "counting", "reorder+onesweep"])
def test_city_100k_collisions_every_tick(msim, orc, small_city, mode, monkeypatch):
    """Munich-style street graph, collisions on: every field and the pair count, tick by tick."""
    monkeypatch.setenv("MSIM_REORDER_EVERY", "4")
    n = 100_000
    ents = small_city.init_entities(n, seed=42)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(small_city, ents, radius=10.0, flags=MODES(msim)[mode]) as sim:
        for t in range(2, 2 + 2 * 30):
            sim.dispatch(t)
            p = oracle_dispatch(orc, want, omap, 10.0, t)
            if t % 2 == 1:
                st = sim.stats()
                assert st["last_pair_count"] == p, f"tick {t}"
                assert st["last_flagged_count"] == int(orc.collision_flags(want).sum()), f"tick {t}"
            if t in (3, 4, 5, 21, 61):
                assert_entities_equal(sim.read_entities(), want, what=f"tick {t}")
        assert_entities_equal(sim.read_entities(), want, what="final")
        flags = sim.read_collision_flags()
        assert (flags == orc.collision_flags(want)).all()
        assert (sim.read_positions() == want["pos"]).all()
        dbg = sim.read_debug()
        assert dbg[0] == n
        assert (sim.stats()["reorders"] > 2) == (mode in ("default", "reorder+onesweep"))


def test_flags_only_mode_matches(msim, orc, small_city):
    n = 50_000
    ents = small_city.init_entities(n, seed=9)
    omap = oracle_map(orc, small_city)
    ticks = list(range(2, 2 + 2 * 20))
    want, _ = run_oracle(orc, ents, omap, 10.0, ticks)
    with msim.Simulation(small_city, ents, radius=10.0, flags=msim.FLAG_NO_PAIR_COUNT) as sim:
        for t in ticks:
            sim.dispatch(t)
        assert_entities_equal(sim.read_entities(), want, what="flags-only")


@pytest.mark.parametrize("mode", ["default", "onesweep", "counting", "reorder+onesweep"])
@pytest.mark.parametrize("radius", [0.0, 0.5, 3.0, 10.0, 37.5, 250.0])
def test_point_clouds_vs_brute_force(msim, orc, small_city, radius, mode):
    """Collision predicate on arbitrary (off-road) positions, including exact-distance edge cases."""
    rng = np.random.default_rng(int(radius * 10) + 1)
    n = 6000
    ents = small_city.init_entities(n, seed=5)
    ents["initialized"] = 1
    xy = (rng.random((n, 2)) * np.array([small_city.width, small_city.height])).astype(np.float32)
    # exact-radius pairs and duplicates: strict '<' must hold
    xy[1] = xy[0] + np.array([radius, 0], dtype=np.float32)
    xy[3] = xy[2]
    xy[5] = xy[4] + np.array([0, np.nextafter(np.float32(radius), np.float32(0))], dtype=np.float32)
    xy[6] = [0, 0]
    xy[7] = [small_city.width, small_city.height]
    ents["pos"] = xy
    ents["target"] = xy  # nobody moves anywhere sensible; only the collision pass is exercised
    want = to_oracle_entities(orc, ents)
    want_pairs = orc.collide_pass_brute(want, radius)
    with msim.Simulation(small_city, ents, radius=radius, flags=MODES(msim)[mode]) as sim:
        sim.dispatch(3)
        got = sim.read_entities()
        st = sim.stats()
    assert_entities_equal(got, want, fields=["color", "pos", "initialized"], what=f"r={radius}")
    assert st["last_pair_count"] == want_pairs


def test_stacked_entities(msim, orc, test_map):
    """App. B10: everybody starts stacked on four points; cells must take unbounded occupancy."""
    n = 3000
    ents = test_map.init_entities(n, seed=1)
    omap = oracle_map(orc, test_map)
    ticks = [2, 3, 4, 5, 6, 7]
    want, want_pairs = run_oracle(orc, ents, omap, 10.0, ticks)
    with msim.Simulation(test_map, ents, radius=10.0) as sim:
        got_pairs = []
        for t in ticks:
            sim.dispatch(t)
            if t % 2 == 1:
                got_pairs.append(sim.stats()["last_pair_count"])
        assert_entities_equal(sim.read_entities(), want, what="stacked")
    assert got_pairs == want_pairs


def test_upload_readback_roundtrip_and_reupload(msim, orc, small_city):
    ents = small_city.init_entities(5000, seed=77)
    ents["direction"] = np.random.default_rng(0).random((5000, 2)).astype(np.float32)
    with msim.Simulation(small_city, ents) as sim:
        assert_entities_equal(sim.read_entities(), ents, what="untouched round trip")
        sim.dispatch(2)
        sim.dispatch(3)
        sim.dispatch(4)
        mid = sim.read_entities()
        # re-upload the read-back state: must continue exactly like an uninterrupted run
        sim.dispatch(5)
        sim.dispatch(6)
        a = sim.read_entities()
        sim.upload(mid)
        sim.dispatch(5)
        sim.dispatch(6)
        b = sim.read_entities()
    assert_entities_equal(a, b, what="re-upload continuation")


def test_asynchronous_ticks_interleaved_with_host_operations(msim, orc, small_city, monkeypatch, n=30_011):
    """Ticks that are only enqueued keep two of them in flight on two streams (pipelined rebuild, overlapped move phase, pass B beside the
    query); every host-visible call in between must see exactly the state of the ticks enqueued so far: statistics, fast readbacks, a full
    readback, a re-upload, a radius change through the push constants, a re-sort every third collision pass."""
    monkeypatch.setenv("MSIM_REORDER_EVERY", "3")
    ents = small_city.init_entities(n, seed=19)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)

    def oracle_ticks(k, radius=10.0):
        pairs = 0
        for _ in range(k):
            orc.move_pass(want, omap, threads=4)
            pairs = orc.collide_pass(want, omap.world_w, omap.world_h, radius, threads=4)
        return pairs

    with msim.Simulation(small_city, ents, radius=10.0) as sim:
        sim.dispatch(2)  # initialise only
        orc.move_pass(want, omap, threads=4)
        for k in (1, 2, 5, 3):
            sim.enqueue_ticks(k, True)
            pairs = oracle_ticks(k)
            st = sim.stats()  # (joins the streams)
            assert (st["last_pair_count"], st["last_flagged_count"]) == (pairs, int(orc.collision_flags(want).sum())), f"after {k} more ticks"
            assert (sim.read_positions() == want["pos"]).all()
        sim.enqueue_ticks(4, True)
        oracle_ticks(4)
        assert (sim.read_collision_flags() == orc.collision_flags(want)).all()
        mid = sim.read_entities()
        assert_entities_equal(mid, want, what="full readback between enqueued ticks")
        sim.enqueue_ticks(2, True)  # ... these two are thrown away by the re-upload
        sim.upload(mid)
        sim.enqueue_ticks(3, True)
        oracle_ticks(3)
        sim.radius = 25.0  # a different grid: work still in flight under the old one is completed first
        sim.dispatch(40)
        orc.move_pass(want, omap, threads=4)
        sim.dispatch(41)
        pairs = orc.collide_pass(want, omap.world_w, omap.world_h, 25.0, threads=4)
        assert sim.stats()["last_pair_count"] == pairs
        sim.enqueue_ticks(2, True)
        pairs = oracle_ticks(2, 25.0)
        assert sim.stats()["last_pair_count"] == pairs
        assert_entities_equal(sim.read_entities(), want, what="after the radius change")


def test_collide_without_prior_move_and_radius_change(msim, orc, small_city):
    ents = small_city.init_entities(20_000, seed=11)
    ents["initialized"] = 1
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    with msim.Simulation(small_city, ents, radius=10.0) as sim:
        sim.dispatch(3)
        p0 = oracle_dispatch(orc, want, omap, 10.0, 3)
        assert sim.stats()["last_pair_count"] == p0
        sim.dispatch(4)
        oracle_dispatch(orc, want, omap, 10.0, 4)
        sim.radius = 25.0  # push constants carry the radius per dispatch (PushConsts.hpp:17)
        sim.dispatch(5)
        p1 = oracle_dispatch(orc, want, omap, 25.0, 5)
        assert sim.stats()["last_pair_count"] == p1
        assert_entities_equal(sim.read_entities(), want, what="after radius change")


@pytest.mark.parametrize("mode", ["default", "counting"])
def test_radius_back_and_forth_reuses_the_counter_table(msim, orc, small_city, mode):
    """A move pass counts the cells of grid A, the radius grows (smaller grid B, same allocation), shrinks back to A: the counters
    the first grid left beyond B's cells must not survive into A's next count (round-1 advisor finding: out-of-bounds ranks)."""
    ents = small_city.init_entities(15_000, seed=29)
    omap = oracle_map(orc, small_city)
    want = to_oracle_entities(orc, ents)
    plan = [(2, 10.0), (3, 10.0), (4, 10.0), (5, 25.0), (6, 25.0), (7, 10.0), (8, 10.0), (9, 4.0), (10, 4.0), (11, 10.0), (13, 25.0), (15, 10.0)]
    with msim.Simulation(small_city, ents, radius=10.0, flags=MODES(msim)[mode]) as sim:
        for tick, radius in plan:
            sim.radius = radius
            sim.dispatch(tick)
            pairs = oracle_dispatch(orc, want, omap, radius, tick)
            if tick % 2 == 1 and tick > 2:
                assert sim.stats()["last_pair_count"] == pairs, f"tick {tick} r={radius}"
        assert_entities_equal(sim.read_entities(), want, what="radius back and forth")


def test_invalid_inputs(msim, small_city):
    ents = small_city.init_entities(100, seed=1)
    bad = ents.copy()
    bad["road_index"][5] = small_city.roads.shape[0] + 3
    with pytest.raises(msim.MsimError) as ei:
        msim.Simulation(small_city, bad)
    assert ei.value.status == msim.MSIM_ERR_INVALID
    mixed = ents.copy()
    mixed["initialized"][::2] = 1
    with pytest.raises(msim.MsimError) as ei:
        msim.Simulation(small_city, mixed)
    assert ei.value.status == msim.MSIM_ERR_UNSUPPORTED
    with msim.Simulation(small_city, ents, flags=msim.FLAG_NO_COLLISIONS) as sim:
        with pytest.raises(msim.MsimError):
            sim.dispatch(3)
