"""Host-side logic of libmsim_cuda.so: map loader (Map::load_from_file semantics), generators with the
generate_map.py connection layout, seeded entity init.  No GPU."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_load_test_map(msim, test_map):
    assert (test_map.width, test_map.height) == (100.0, 100.0)
    assert test_map.roads.shape[0] == 4 and test_map.connections.shape[0] == 20
    r = test_map.roads
    assert r["start_pos"].tolist() == [[5, 5], [95, 5], [5, 95], [95, 95]]
    assert (r["end_pos"] == 50).all()
    assert r["start_count"].tolist() == [1, 1, 1, 1] and r["end_count"].tolist() == [4, 4, 4, 4]
    assert r["end_index"].tolist() == [1, 6, 11, 16]
    assert test_map.connections.tolist() == [0, 0, 1, 2, 3, 1, 1, 0, 2, 3, 2, 2, 1, 0, 3, 3, 3, 1, 2, 0]


def test_loader_errors_mirror_reference(msim, tmp_path):
    with pytest.raises(msim.MsimError) as ei:  # Map.cpp:30-33 -> nullptr; here MSIM_ERR_IO
        msim.Map.load_json(str(tmp_path / "nope.json"))
    assert ei.value.status == msim.MSIM_ERR_IO
    doc = json.load(open(os.path.join(GOLDEN, "test_map.json")))
    for field, msg in (("maxDistLat", "'maxDistLat' field missing"), ("connectionRoadIndexList", "'connectionRoadIndexList' field missing")):
        d = dict(doc)
        del d[field]
        p = tmp_path / f"no_{field}.json"
        p.write_text(json.dumps(d))
        with pytest.raises(msim.MsimError) as ei:  # Map.cpp:42-44,... throw std::runtime_error
            msim.Map.load_json(str(p))
        assert ei.value.status == msim.MSIM_ERR_PARSE and msg in ei.value.message
    d = json.loads(json.dumps(doc))
    del d["roads"][1]["connCountEnd"]
    p = tmp_path / "no_cce.json"
    p.write_text(json.dumps(d))
    with pytest.raises(msim.MsimError) as ei:
        msim.Map.load_json(str(p))
    assert "'connCountEnd' field missing" in ei.value.message
    p = tmp_path / "garbage.json"
    p.write_text("{\"maxDistLat\": ")
    with pytest.raises(msim.MsimError):
        msim.Map.load_json(str(p))


def test_zero_length_roads_are_skipped_without_reindexing(msim, tmp_path):
    """SURVEY App. B3 / Map.cpp:124-128."""
    doc = json.load(open(os.path.join(GOLDEN, "test_map.json")))
    doc["roads"].insert(1, {"start": {"distLat": 7, "distLong": 7}, "end": {"distLat": 7, "distLong": 7},
                            "connIndexStart": 0, "connCountStart": 1, "connIndexEnd": 0, "connCountEnd": 1})
    p = tmp_path / "zero.json"
    p.write_text(json.dumps(doc))
    m = msim.Map.load_json(str(p))
    assert m.roads.shape[0] == 4 and m.connections.shape[0] == 20


def test_save_load_round_trip_is_bit_exact(msim, small_city, tmp_path):
    L = msim.lib()
    import ctypes as C

    h = C.c_void_p()
    assert L.msim_map_generate_city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7, C.byref(h)) == 0
    path = str(tmp_path / "city.json")
    assert L.msim_map_save_json(h, path.encode()) == 0
    L.msim_map_free(h)
    again = msim.Map.load_json(path)
    assert again.roads.tobytes() == small_city.roads.tobytes()
    assert again.connections.tobytes() == small_city.connections.tobytes()
    assert (again.width, again.height) == (small_city.width, small_city.height)


def check_generate_map_layout(m):
    """Connection table as map/generate_map.py:234-258 emits it: one shared block per coordinate, every
    road of the coordinate once, END-matching roads twice; connIndex = block start, connCount = roads."""
    r, c = m.roads, m.connections
    coord_roads = {}
    for i in range(r.shape[0]):
        coord_roads.setdefault((tuple(r["start_pos"][i]), int(r["start_index"][i])), []).append((i, "s"))
        coord_roads.setdefault((tuple(r["end_pos"][i]), int(r["end_index"][i])), []).append((i, "e"))
    used = 0
    for (pos, index), members in coord_roads.items():
        count = len(members)
        expect = []
        for i, kind in sorted(members):
            expect += [i] if kind == "s" else [i, i]
            got_count = r["start_count"][i] if kind == "s" else r["end_count"][i]
            assert got_count == count
        assert sorted(c[index : index + len(expect)].tolist()) == sorted(expect)
        used += len(expect)
    assert used == c.shape[0]
    # one block per geometric coordinate
    assert len({p for p, _ in coord_roads}) == len(coord_roads)


def test_city_generator_layout_and_determinism(msim, small_city):
    check_generate_map_layout(small_city)
    again = msim.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7)
    assert again.roads.tobytes() == small_city.roads.tobytes()
    other = msim.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 8)
    assert other.roads.tobytes() != small_city.roads.tobytes()
    r = small_city.roads
    ln = np.hypot(*(r["start_pos"] - r["end_pos"]).T)
    assert ln.min() > 0 and ln.max() < 120
    assert (r["start_pos"] >= 0).all() and (r["end_pos"][:, 0] <= small_city.width).all() and (r["end_pos"][:, 1] <= small_city.height).all()
    degrees = np.concatenate([r["start_count"], r["end_count"]])
    assert degrees.min() == 1 and degrees.max() >= 5  # dead ends and 5-way junctions exist
    assert (small_city.connections < r.shape[0]).all()


def test_grid_generator(msim):
    g = msim.Map.grid(6, 5, 20.0)
    assert g.roads.shape[0] == 5 * 5 + 4 * 6  # (nx-1)*ny + (ny-1)*nx
    assert (g.width, g.height) == (100.0, 80.0)
    check_generate_map_layout(g)
    assert set(np.unique(g.roads["start_pos"])) <= set(np.arange(0, 101, 20.0))


def test_entity_init_mirrors_add_entities(msim, small_city):
    """Simulator::add_entities (Simulator.cpp:114-129): pos = road.start, target = road.end, dir = 0,
    initialized = 0, colour in [0,1) with alpha 1; RNG words are raw mt19937(seed+2) outputs."""
    n = 5000
    e = small_city.init_entities(n, seed=42)
    r = small_city.roads
    assert (e["pos"] == r["start_pos"][e["road_index"]]).all()
    assert (e["target"] == r["end_pos"][e["road_index"]]).all()
    assert (e["direction"] == 0).all() and (e["initialized"] == 0).all()
    assert (e["color"][:, 3] == 1).all() and (e["color"][:, :3] >= 0).all() and (e["color"][:, :3] < 1).all()
    raw = np.frombuffer(np.random.RandomState(44).bytes(4 * 4 * n), dtype="<u4").reshape(n, 4)  # init_genrand(44) == std::mt19937(44)
    assert (e["rand_state"] == raw).all()
    assert small_city.init_entities(n, seed=42).tobytes() == e.tobytes()
    assert small_city.init_entities(n, seed=43).tobytes() != e.tobytes()
    assert len(np.unique(e["road_index"])) > 1500


def test_entity_init_box_restriction(msim, small_city):
    box = [900, 600, 1300, 1000]
    e = small_city.init_entities(2000, seed=5, box=box)
    r = small_city.roads[e["road_index"]]
    for pos in (r["start_pos"], r["end_pos"]):
        assert (pos[:, 0] >= box[0]).all() and (pos[:, 0] <= box[2]).all() and (pos[:, 1] >= box[1]).all() and (pos[:, 1] <= box[3]).all()
    with pytest.raises(msim.MsimError):
        small_city.init_entities(10, seed=5, box=[-5, -5, -1, -1])


def test_road_index_stream_alone_matches_entity_init(msim, small_city):
    """msim_entities_init_roads = the road-index generator of msim_entities_init on its own (the sharded host's partition histogram)."""
    box = np.array([0.2 * small_city.width, 0.2 * small_city.height, 0.8 * small_city.width, 0.8 * small_city.height], dtype=np.float32)
    for count, seed, bx in ((0, 1, None), (1, 2, None), (50_000, 42, None), (33_333, 7, box)):
        want = small_city.init_entities(count, seed=seed, box=bx)["road_index"]
        assert np.array_equal(small_city.init_road_indices(count, seed=seed, box=bx), want)
    with pytest.raises(msim.MsimError):
        small_city.init_road_indices(10, seed=1, box=np.array([-5.0, -5.0, -4.0, -4.0], dtype=np.float32))


def test_partition_histogram_equals_the_full_population(msim, small_city):
    from movement_sim_b200 import sharding as S

    total, seed = 2 * S.CHUNK + 12_345, 42  # three seeded chunks
    want = None
    for _, ents in S.generate_population(msim, small_city, total, seed):
        rows, ncx, ncy = msim.grid_rows(small_city.width, small_city.height, 10.0, ents["pos"])
        h = np.bincount(rows, minlength=ncy).astype(np.int64)
        want = h if want is None else want + h
    hist, gx, gy = S.global_row_histogram(msim, small_city, total, seed, 10.0)
    assert (gx, gy) == (ncx, ncy) and np.array_equal(hist, want) and int(hist.sum()) == total


def test_population_build_is_independent_of_the_thread_count(msim, small_city):
    from movement_sim_b200 import sharding as S

    total, seed = 3 * S.CHUNK + 777, 5
    h1 = S.global_row_histogram(msim, small_city, total, seed, 10.0, threads=1)
    h3 = S.global_row_histogram(msim, small_city, total, seed, 10.0, threads=3)
    assert np.array_equal(h1[0], h3[0]) and h1[1:] == h3[1:]
    e1, g1 = S.collect_band(msim, small_city, total, seed, 10.0, 20, 90, threads=1)
    e3, g3 = S.collect_band(msim, small_city, total, seed, 10.0, 20, 90, threads=3)
    assert e1.tobytes() == e3.tobytes() and np.array_equal(g1, g3) and g1.shape[0] == e1.shape[0] > 0
    assert S.host_threads(1) >= 1 and S.host_threads(10_000) == 1
