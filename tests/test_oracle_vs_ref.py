"""Pins the oracle's collision half against the reference's OWN CPU restatement of the shader's
quadtree (/root/reference/shader_validation/src/main.cpp), compiled into oracle/_ref by oracle/Makefile.
Skipped when oracle/_ref has not been built (fresh clone without /root/reference)."""
import numpy as np
import pytest

from conftest import oracle_map, to_oracle_entities

f32 = np.float32


@pytest.fixture(scope="module")
def ref(orc):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return orc


def test_reference_known_answer_test_passes(ref):
    """run_collision_detection_test_1 (shader_validation/src/main.cpp:1137-1199): the reference's only
    KAT, disabled upstream (`// run_tests();`, :1205-1207).  It asserts internally; rc 0 = all held."""
    r = ref.run_ref_kat()
    assert r.returncode == 0, r.stderr[-400:]
    assert "run_collision_detection_test_1 was successful." in r.stdout


def test_reference_kat_sequence_on_oracle(ref):
    """The same five hand-placed points through the oracle: unique pairs must be the KAT's final
    answers (1 after the second point, 2 after the fifth)."""
    w, h = f32(29007.4609), f32(16463.7656)
    pts = [(w / 2 - 3, 0), (w / 2 + 3, 0), (0, 0), (w / 4, 0), (w / 4 - 9, 0)]
    want_pairs = [0, 1, 1, 1, 2]
    for k in range(1, 6):
        e = np.zeros(k, dtype=ref.ENTITY_DTYPE)
        e["pos"] = np.array(pts[:k], dtype=f32)
        e["initialized"] = 1
        assert ref.collide_pass(e, float(w), float(h), 10.0) == want_pairs[k - 1]
        q = ref.RefQuadTree(float(w), float(h), 10.0, 100)  # harness default cap (main.cpp:240)
        q.insert(e["pos"])
        flags, _ = q.collide()
        assert (flags == ref.collision_flags(e)).all()


def test_reference_capacity_and_node_count(ref):
    assert ref.ref().ref_node_count() == 21845 == ref.calc_node_count(8)
    assert ref.ref().ref_capacity() >= 1_000_000


@pytest.mark.parametrize("seed,n,world", [(0, 3000, (2000.0, 1500.0)), (1, 20_000, (29007.4609, 16463.7656)), (2, 12_000, (400.0, 300.0))])
def test_flagged_set_equals_reference_quadtree_on_point_clouds(ref, seed, n, world):
    """Shader cap 10 / depth 8 (src/sim/Simulator.hpp:37-38): the reference tree's recoloured set must
    equal the oracle's flags; its debugData[1] may only over-count."""
    rng = np.random.default_rng(seed)
    xy = (rng.random((n, 2)) * np.array(world)).astype(f32)
    xy[: n // 10] = xy[n // 10 : 2 * (n // 10)]  # duplicates exercise same_pos_as_first (:301-307)
    e = np.zeros(n, dtype=ref.ENTITY_DTYPE)
    e["pos"] = xy
    e["initialized"] = 1
    pairs = ref.collide_pass(e, world[0], world[1], 10.0)
    q = ref.RefQuadTree(world[0], world[1], 10.0, 10)
    q.insert(xy, threads=1)
    assert q.count_in_tree() == n
    flags, debug1 = q.collide(threads=4)
    assert (flags == ref.collision_flags(e)).all()
    assert debug1 >= pairs


def test_flagged_set_equals_reference_on_street_population(ref, small_city):
    ents = small_city.init_entities(60_000, seed=42)
    om = oracle_map(ref, small_city)
    e = to_oracle_entities(ref, ents)
    for _ in range(1 + 120):
        ref.move_pass(e, om, threads=4)
    q = ref.RefQuadTree(small_city.width, small_city.height, 10.0, 10)
    q.insert(e["pos"], threads=1)
    flags, _ = q.collide(threads=4)
    ref.collide_pass(e, small_city.width, small_city.height, 10.0, threads=4)
    assert (flags == ref.collision_flags(e)).all()
