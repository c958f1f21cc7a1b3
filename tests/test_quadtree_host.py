"""Display quadtree without a GPU: msim_quadtree_from_positions - the host twin of msim_read_quadtree_nodes, same
builder, leaf histogram taken on the host - against the tree the reference's OWN shader code builds when it inserts the same positions
(oracle/_ref/libref_shader_full.so: quad_tree_insert / quad_tree_split_up_node of random_move.comp, first dispatch).  Compared depth first
(TL, TR, BL, BR - the order QuadTreeGridGlObject::add_node_rec walks, src/ui/widgets/opengl/QuadTreeGridGlObject.cpp:10-51): rectangle,
content type and leaf entity count of every node.  Positions are distinct: the shader keeps entities with identical positions on one leaf
beyond its capacity (quad_tree_same_pos_as_fist, :301-307), a tie effect the product does not reproduce (csrc/quadtree.cu header)."""
import numpy as np
import pytest


def dfs_reference(nodes):
    out = []

    def walk(i):
        nd = nodes[i]
        out.append((float(nd["offsetX"]), float(nd["offsetY"]), float(nd["width"]), float(nd["height"]), int(nd["contentType"]),
                    int(nd["entityCount"]) if nd["contentType"] == 2 else 0))
        if nd["contentType"] == 1:
            for child in ("nextTL", "nextTR", "nextBL", "nextBR"):
                walk(int(nd[child]))

    walk(0)
    return out


def dfs_product(nodes):
    out = []

    def walk(i):
        nd = nodes[i]
        out.append((float(nd["offset_x"]), float(nd["offset_y"]), float(nd["width"]), float(nd["height"]), int(nd["content_type"]),
                    int(nd["entity_count"]) if nd["content_type"] == 2 else 0))
        if nd["content_type"] == 1:
            for child in ("next_tl", "next_tr", "next_bl", "next_br"):
                assert nodes[int(nd[child])]["prev_node_index"] == i
                walk(int(nd[child]))

    walk(0)
    return out


def clouds(rnd, w, h, n):
    uniform = np.stack([rnd.uniform(0, w, n), rnd.uniform(0, h, n)], axis=1)
    centres = np.stack([rnd.uniform(0, w, 12), rnd.uniform(0, h, 12)], axis=1)
    clustered = np.clip(centres[rnd.integers(12, size=n)] + rnd.normal(0, min(w, h) / 150.0, (n, 2)), 0, [np.nextafter(np.float32(w), np.float32(0)),
                                                                                                          np.nextafter(np.float32(h), np.float32(0))])
    return {"uniform": uniform.astype(np.float32), "clustered": clustered.astype(np.float32)}


@pytest.mark.timeout(600)
@pytest.mark.parametrize("depth,cap", [(8, 10), (8, 1), (5, 10), (3, 4), (1, 10)])
@pytest.mark.parametrize("n", [1, 9, 11, 500, 20_000])
def test_host_quadtree_equals_the_tree_the_shader_builds(msim, orc, test_map, n, depth, cap):
    if not orc.ref_shader_full_available():
        pytest.skip("oracle/_ref/libref_shader_full.so not built (needs /root/reference at build time)")
    w, h = 29007.4609, 16463.7656
    rnd = np.random.default_rng(n * 100 + depth * 10 + cap)
    for name, xy in clouds(rnd, w, h, n).items():
        xy = np.unique(xy, axis=0)  # distinct positions (see the module docstring)
        rnd.shuffle(xy)
        e = np.zeros(xy.shape[0], dtype=orc.ENTITY_DTYPE)
        e["pos"] = xy
        om = orc.OracleMap(w, h, test_map.roads.view(orc.ROAD_DTYPE), test_map.connections)  # the map plays no role in the first dispatch
        sim = orc.RefShaderSim(e, om, radius=10.0, max_depth=depth, node_cap=cap)
        sim.dispatch(2)  # initialise: quad_tree_insert(index, 0, 1) for every entity
        want = dfs_reference(sim.nodes)
        got = dfs_product(msim.quadtree_from_positions(xy, w, h, depth, cap))
        assert len(got) == len(want), f"{name}: {len(got)} nodes, the shader built {len(want)}"
        assert got == want, f"{name}: first difference at node {next(i for i, (a, b) in enumerate(zip(got, want)) if a != b)}"
        assert sum(t[5] for t in got) == xy.shape[0]


def test_host_quadtree_argument_checks(msim):
    import ctypes as C

    L = msim.lib()
    n = C.c_uint64()
    out = np.zeros(4, dtype=msim.QUADTREE_NODE_DTYPE)
    xy = np.array([[1.0, 1.0], [2.0, 2.0], [3.0, 3.0]], dtype=np.float32)
    assert L.msim_quadtree_from_positions(None, 3, 10.0, 10.0, 8, 10, out.ctypes.data, 4, C.byref(n)) == msim.MSIM_ERR_INVALID
    assert L.msim_quadtree_from_positions(xy.ctypes.data, 3, 10.0, 10.0, 8, 1, out.ctypes.data, 4, C.byref(n)) == msim.MSIM_ERR_CAPACITY
    assert L.msim_quadtree_from_positions(xy.ctypes.data, 0, 10.0, 10.0, 8, 10, out.ctypes.data, 4, C.byref(n)) == msim.MSIM_OK and n.value == 1
    root = out[0]
    assert root["width"] == 10.0 and root["content_type"] == 2 and root["entity_count"] == 0
