"""msim_read_quadtree_nodes: the display quadtree built from the current positions
must be the tree the REFERENCE's own quad_tree_insert builds for the same points (fresh insertion,
cap 10, depth 8; compiled reference harness in oracle/_ref), node for node in the order the UI walks it
(src/ui/widgets/opengl/QuadTreeGridGlObject.cpp:29-51)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def walk(nodes):
    """Depth-first TL, TR, BL, BR walk from node 0, like QuadTreeGridGlObject::add_node_rec."""
    rects, types, counts = [], [], []

    def rec(i):
        nd = nodes[i]
        rects.append((nd["offset_x"], nd["offset_y"], nd["width"], nd["height"]))
        types.append(int(nd["content_type"]))
        counts.append(int(nd["entity_count"]))
        if nd["content_type"] == 1:
            for k in ("next_tl", "next_tr", "next_bl", "next_br"):
                assert nodes[nd[k]]["prev_node_index"] == i
                rec(int(nd[k]))

    rec(0)
    return np.array(rects, dtype=np.float32), np.array(types), np.array(counts)


@pytest.mark.parametrize("n,seed", [(1, 0), (11, 1), (5000, 2), (200_000, 3)])
def test_display_quadtree_equals_reference_tree(msim, orc, n, seed):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    m = msim.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7) if n < 100_000 else msim.Map.city()
    ents = m.init_entities(n, seed=seed)
    ents["initialized"] = 1
    rng = np.random.default_rng(seed)
    xy = (rng.random((n, 2)) * np.array([m.width, m.height])).astype(np.float32)  # distinct points (no same-position ties)
    ents["pos"] = xy
    ents["target"] = xy
    q = orc.RefQuadTree(m.width, m.height, 10.0, 10)
    q.insert(xy, threads=1)
    want_rects, want_types, want_counts = q.dump_tree()
    with msim.Simulation(m, ents, radius=10.0) as sim:
        nodes = sim.read_quadtree_nodes()
    got_rects, got_types, got_counts = walk(nodes)
    assert len(got_types) == len(want_types) == len(nodes)
    assert (got_types == want_types).all()
    assert got_rects.tobytes() == want_rects.tobytes()
    leaves = want_types == 2
    assert (got_counts[leaves] == want_counts[leaves]).all()
    assert got_counts[leaves].sum() == n


def test_display_quadtree_follows_the_simulation(msim, orc, small_city):
    ents = small_city.init_entities(30_000, seed=4)
    with msim.Simulation(small_city, ents) as sim:
        root = sim.read_quadtree_nodes()
        assert len(root) == 1 and root[0]["content_type"] == 2 and root[0]["entity_count"] == 0  # nothing inserted yet
        assert (root[0]["width"], root[0]["height"]) == (np.float32(small_city.width), np.float32(small_city.height))
        for t in range(2, 42):
            sim.dispatch(t)
        nodes = sim.read_quadtree_nodes()
        _, types, counts = walk(nodes)
        assert counts[types == 2].sum() == 30_000 and (types == 1).sum() > 100
        assert len(nodes) <= msim.calc_node_count(8)
    with msim.Simulation(small_city, ents, flags=msim.FLAG_NO_QUADTREE) as sim:
        sim.dispatch(2)
        assert len(sim.read_quadtree_nodes()) == 1
