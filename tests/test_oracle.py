"""CPU tests of the ORACLE (oracle/msim_oracle.c): known answers derived by hand / by an independent
numpy restatement of the shader text, plus regression digests in tests/golden/ (made by
tests/golden/make_golden.py).  No GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, oracle_map, to_oracle_entities

f32 = np.float32


def py_xorshift128(s):
    """Independent restatement of random_move.comp:725-736 on Python ints."""
    x, y, z, w = s
    t = w
    t ^= (t << 11) & 0xFFFFFFFF
    t ^= t >> 8
    nx = (t ^ x ^ (x >> 19)) & 0xFFFFFFFF
    return nx, [nx, x, y, z]


def np_next_range(u, lo, hi):
    """random_move.comp:738-746 in numpy binary32 (each operation rounded separately)."""
    f = f32(np.uint32(u)) * f32(2.0 ** -32)
    prod = f32(f * f32(hi - lo + 1))
    s = f32(f32(lo) + prod)
    return int(np.ceil(s)) - 1


def test_xorshift128_known_answers(orc):
    # state (1,2,3,4): t = 4 ^ (4<<11) = 8196; t ^= t>>8 -> 8228; x' = 8228 ^ 1 ^ 0 = 8229
    s = np.array([1, 2, 3, 4], dtype=np.uint32)
    assert orc.xorshift128(s) == 8229
    assert s.tolist() == [8229, 1, 2, 3]
    rng = np.random.default_rng(1)
    for _ in range(50):
        st = rng.integers(0, 2**32, 4, dtype=np.uint64).astype(np.uint32)
        ref = [int(v) for v in st]
        for _ in range(20):
            want, ref = py_xorshift128(ref)
            assert orc.xorshift128(st) == want
            assert st.tolist() == ref


def test_xorshift_zero_state_is_fixed_point(orc):
    s = np.zeros(4, dtype=np.uint32)
    assert orc.xorshift128(s) == 0 and s.tolist() == [0, 0, 0, 0]


def test_next_float_range_and_rounding(orc):
    # choose states whose next output is a given u: with state (x,y,z,w), u = t(w) ^ x ^ (x>>19)
    for u_target in (0, 1, 64, 65, 2**24 + 1, 2**31, 0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFF):
        s = np.array([0, 0, 0, 0], dtype=np.uint32)
        # w = 0 -> t = 0 -> u = x ^ (x >> 19): solve for x top-down
        x = 0
        for bit in range(31, -1, -1):
            want = (u_target >> bit) & 1
            hi = (x >> (bit + 19)) & 1 if bit + 19 <= 31 else 0
            x |= (want ^ hi) << bit
        s[0] = x
        st = s.copy()
        assert orc.xorshift128(st) == u_target
        st = s.copy()
        f = orc.next_float(st)
        assert f == f32(np.uint32(u_target)) * f32(2.0 ** -32)
        assert 0.0 <= f <= 1.0
    # u = 0xFFFFFFFF rounds to 2^32 -> exactly 1.0
    assert f32(np.uint32(0xFFFFFFFF)) * f32(2.0 ** -32) == f32(1.0)


def test_next_range_matches_numpy_restatement(orc):
    rng = np.random.default_rng(2)
    for count in (3, 4, 5, 6, 7, 100, 2**20):
        for _ in range(400):
            st = rng.integers(0, 2**32, 4, dtype=np.uint64).astype(np.uint32)
            u, _ = py_xorshift128([int(v) for v in st])
            got = orc.next_range(st, 1, count)
            assert got == np_next_range(u, 1, count)
            assert 0 <= got <= count


def test_next_range_edges(orc):
    # u tiny -> 1 + f*count rounds to 1.0 -> offset 0 ; u max -> offset == count (one past the block, App. B1)
    def state_for(u):
        x = 0
        for bit in range(31, -1, -1):
            want = (u >> bit) & 1
            hi = (x >> (bit + 19)) & 1 if bit + 19 <= 31 else 0
            x |= (want ^ hi) << bit
        return np.array([x, 0, 0, 0], dtype=np.uint32)

    assert orc.next_range(state_for(0), 1, 4) == 0
    assert orc.next_range(state_for(64), 1, 4) == 0
    assert orc.next_range(state_for(65), 1, 4) == 1
    assert orc.next_range(state_for(0xFFFFFFFF), 1, 4) == 4
    assert orc.next_range(state_for(0x7FFFFFFF), 1, 4) == 2


def test_calc_node_count_reference_asserts(orc, msim):
    # /root/reference/src/sim/Simulator.cpp:72-76
    for depth, want in ((1, 1), (2, 5), (3, 21), (4, 85), (8, 21845)):
        assert orc.calc_node_count(depth) == want
        assert msim.calc_node_count(depth) == want


def test_first_dispatch_only_initialises(orc, test_map):
    ents = test_map.init_entities(100, seed=42)
    e = to_oracle_entities(orc, ents)
    before = e.copy()
    st = orc.move_pass(e, oracle_map(orc, test_map))
    assert st["initialised"] == 100 and st["moved"] == 0
    assert (e["initialized"] == 1).all()
    e2 = e.copy()
    e2["initialized"] = 0
    assert e2.tobytes() == before.tobytes()


def test_straight_walk_and_arrival_on_test_map(orc, test_map):
    """Road 0 runs (5,5)->(50,50): length 63.64 m, 1.4 m per pass -> 45 walking passes, arrival on the 46th."""
    ents = test_map.init_entities(1, seed=42)
    ents["road_index"] = 0
    ents["pos"] = [5, 5]
    ents["target"] = [50, 50]
    ents["initialized"] = 1
    e = to_oracle_entities(orc, ents)
    om = oracle_map(orc, test_map)
    p = np.array([5, 5], dtype=f32)
    t = np.array([50, 50], dtype=f32)
    for step in range(45):
        d = t - p
        ln = np.sqrt(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])))
        assert ln > f32(1.4)
        direction = (d / ln) * f32(1.4)
        p = p + direction
        st = orc.move_pass(e, om)
        assert st["arrivals"] == 0
        assert e["pos"][0].tobytes() == p.tobytes(), step
        assert e["dir"][0].tobytes() == direction.astype(f32).tobytes()
    st = orc.move_pass(e, om)
    assert st["arrivals"] == 1 and st["rng_draws"] == 1  # junction of 4 roads -> random choice
    assert e["pos"][0].tolist() == [50.0, 50.0]
    assert e["target"][0].tolist() != [50.0, 50.0]
    # direction now points from the junction to the new target (random_move.comp:848-850)
    d = e["target"][0] - e["pos"][0]
    ln = np.sqrt(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])))
    assert e["dir"][0].tobytes() == ((d / ln) * f32(1.4)).astype(f32).tobytes()


def test_dead_end_turns_around(orc, test_map):
    ents = test_map.init_entities(1, seed=1)
    ents["road_index"] = 2
    ents["pos"] = [5.5, 94.5]
    ents["target"] = [5, 95]  # start coordinate of road 2, connectedCount 1
    ents["initialized"] = 1
    e = to_oracle_entities(orc, ents)
    rng_before = e["rng"].copy()
    st = orc.move_pass(e, oracle_map(orc, test_map))
    assert st["arrivals"] == 1 and st["uturns"] == 1 and st["rng_draws"] == 0
    assert e["pos"][0].tolist() == [5.0, 95.0]
    assert e["target"][0].tolist() == [50.0, 50.0]
    assert e["road"][0] == 2
    assert (e["rng"] == rng_before).all()


def test_out_of_bounds_connection_read_yields_road_zero(orc, test_map):
    """App. B1: road 3's end block is connections[16..19]; offset 4 reads index 20 of a 20-entry table."""
    ents = test_map.init_entities(1, seed=1)
    ents["road_index"] = 3
    ents["pos"] = [50.5, 50.5]
    ents["target"] = [50, 50]
    ents["initialized"] = 1
    x = 0
    for bit in range(31, -1, -1):  # state whose next output is 0xFFFFFFFF -> offset == count == 4
        hi = (x >> (bit + 19)) & 1 if bit + 19 <= 31 else 0
        x |= (1 ^ hi) << bit
    ents["rand_state"] = [x, 0, 0, 0]
    e = to_oracle_entities(orc, ents)
    st = orc.move_pass(e, oracle_map(orc, test_map))
    assert st["oob_reads"] == 1
    assert e["road"][0] == 0
    assert e["target"][0].tolist() == [5.0, 5.0]  # road 0 ends at (50,50) == reached point -> go to its start


def test_two_way_junction_takes_second_entry_without_rng(orc, msim):
    # chain a-b-c: coordinate b has connectedCount 2 -> connections[index + 1], no RNG draw (:802-804)
    roads = np.zeros(2, dtype=msim.ROAD_DTYPE)
    roads[0] = ((0, 0), 0, 1, (10, 0), 1, 2)
    roads[1] = ((10, 0), 1, 2, (20, 0), 3, 1)
    conns = np.array([0, 0, 1, 1], dtype=np.uint32)
    m = msim.Map(20, 1, roads, conns)
    ents = m.init_entities(1, seed=3)
    ents["road_index"] = 0
    ents["pos"] = [9.5, 0]
    ents["target"] = [10, 0]
    ents["initialized"] = 1
    e = to_oracle_entities(orc, ents)
    rng_before = e["rng"].copy()
    st = orc.move_pass(e, oracle_map(orc, m))
    assert st["arrivals"] == 1 and st["rng_draws"] == 0
    assert e["road"][0] == 1 and e["target"][0].tolist() == [20.0, 0.0]
    assert (e["rng"] == rng_before).all()


def test_in_range_predicate(orc):
    assert orc.in_range([0, 0], [3, 4], 5.0) is False  # strict '<' (random_move.comp:561)
    assert orc.in_range([0, 0], [3, 4], np.nextafter(f32(5), f32(6))) is True
    assert orc.in_range([0, 0], [11, 0], 10.0) is False
    assert orc.in_range([1, 1], [1, 1], 10.0) is True
    assert orc.in_range([1, 1], [1, 1], 0.0) is False


@pytest.mark.parametrize("radius", [0.5, 10.0, 60.0])
def test_collide_grid_equals_brute_force(orc, radius):
    rng = np.random.default_rng(int(radius * 7))
    n = 3000
    e = np.zeros(n, dtype=orc.ENTITY_DTYPE)
    e["pos"] = (rng.random((n, 2)) * [900, 700]).astype(f32)
    e["pos"][1] = e["pos"][0] + np.array([radius, 0], dtype=f32)
    e["pos"][3] = e["pos"][2]
    e["initialized"] = 1
    e["initialized"][10:20] = 0  # uninitialised entities are not in the neighbour structure
    a, b = e.copy(), e.copy()
    pa = orc.collide_pass(a, 900, 700, radius)
    pb = orc.collide_pass_brute(b, radius)
    assert pa == pb
    assert a.tobytes() == b.tobytes()
    c = e.copy()
    assert orc.collide_pass(c, 900, 700, radius, threads=4) == pa and c.tobytes() == a.tobytes()
    flagged = orc.collision_flags(a)
    assert flagged[3] == 1 and flagged[2] == 1 and flagged[10:20].sum() == 0
    assert (a["initialized"] == 1).all()


def test_move_multithreaded_equals_single(orc, small_city):
    ents = small_city.init_entities(20_000, seed=8)
    om = oracle_map(orc, small_city)
    a, b = to_oracle_entities(orc, ents), to_oracle_entities(orc, ents)
    for _ in range(60):
        sa = orc.move_pass(a, om, threads=1)
        sb = orc.move_pass(b, om, threads=5)
        assert sa == sb
    assert a.tobytes() == b.tobytes()


def test_golden_digests(orc, msim, test_map, small_city):
    """Regression pins of the oracle itself (tests/golden/oracle_digests.json, made by make_golden.py)."""
    from golden.make_golden import compute_digests

    with open(os.path.join(GOLDEN, "oracle_digests.json")) as f:
        want = json.load(f)
    got = compute_digests()
    assert got == want
