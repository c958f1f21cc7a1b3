"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo runs of movement-sim_b200/sharding.py with
the oracle-backed engine, compared entity by entity with an unsharded oracle run; plus unit tests of
the band partitioner.  The same worker runs on GPUs over NCCL in test_gpu_sharding.py."""
import json
import os
import socket

import numpy as np
import pytest

from conftest import oracle_map, to_oracle_entities


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def reference_run(O, M, cfg):
    """Unsharded oracle run of the same seeded global population."""
    from movement_sim_b200 import sharding as S
    from shard_worker import make_map

    m = make_map(M, cfg)
    parts = [e for _, e in S.generate_population(M, m, cfg["entities"], cfg["seed"], cfg.get("box"))]
    e = to_oracle_entities(O, np.concatenate(parts))
    om = oracle_map(O, m)
    O.move_pass(e, om)
    pairs = []
    for _ in range(cfg["ticks"]):
        O.move_pass(e, om)
        pairs.append(O.collide_pass(e, m.width, m.height, cfg["radius"]))
    return e, pairs


def check_against_reference(O, M, cfg, outdir, world):
    want, want_pairs = reference_run(O, M, cfg)
    got = np.zeros_like(want)
    seen = np.zeros(len(want), dtype=np.int32)
    metas = []
    for r in range(world):
        e = np.load(os.path.join(outdir, f"ents_{r}.npy")).view(O.ENTITY_DTYPE).reshape(-1)
        g = np.load(os.path.join(outdir, f"gids_{r}.npy"))
        assert len(e) == len(g)
        got[g] = e
        seen[g] += 1
        metas.append(json.load(open(os.path.join(outdir, f"meta_{r}.json"))))
    assert (seen == 1).all(), "every entity must be owned by exactly one rank"
    for f in ("pos", "target", "rng", "road", "color", "dir", "initialized"):
        bad = np.nonzero((got[f] != want[f]).reshape(len(want), -1).any(axis=1))[0]
        assert bad.size == 0, f"field {f}: {bad.size} entities differ, first gid {bad[:1]}"
    for m in metas:
        assert m["pairs"] == want_pairs  # the all-reduced pair count is identical on every rank and exact
    return metas


BASE = {"map": "city", "city": [1500.0, 1000.0, 35.0, 0.3, 0.12, 3], "entities": 6000, "seed": 42, "radius": 10.0, "ticks": 40,
        "capacity": 1 << 13, "rebalance_every": 8}


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gloo_matches_unsharded_oracle(orc, msim, tmp_path, world):
    import torch.multiprocessing as mp

    from shard_worker import run

    cfg = dict(BASE)
    mp.spawn(run, args=(world, free_port(), "gloo", str(tmp_path), cfg), nprocs=world, join=True)
    metas = check_against_reference(orc, msim, cfg, str(tmp_path), world)
    assert metas[0]["exchanged_bytes"] > 0


def test_peer_memory_request_falls_back_to_the_collective_when_unavailable(orc, msim, tmp_path):
    """exchange="p2p" on an engine / box that cannot do it (here: the CPU oracle engine over gloo): every rank must agree on
    the fall-back and the run must still match."""
    import torch.multiprocessing as mp

    from shard_worker import run

    cfg = dict(BASE, exchange="p2p", ticks=20)
    mp.spawn(run, args=(2, free_port(), "gloo", str(tmp_path), cfg), nprocs=2, join=True)
    metas = check_against_reference(orc, msim, cfg, str(tmp_path), 2)
    assert all(m["exchange"] == "collective" and m["p2p_error"] for m in metas)
    assert metas[0]["exchanged_bytes"] > 0


def test_rebalancing_from_a_skewed_partition(orc, msim, tmp_path):
    """Dense crowd (BASELINE config 5 in miniature): everybody starts in one corner and the initial
    split is geometric; the histogram all-reduce must walk the boundary until the load is even."""
    import torch.multiprocessing as mp

    from shard_worker import run

    cfg = dict(BASE, box=[0.0, 0.0, 700.0, 450.0], skew_splits=True, ticks=60, rebalance_every=4, entities=5000)
    mp.spawn(run, args=(2, free_port(), "gloo", str(tmp_path), cfg), nprocs=2, join=True)
    metas = check_against_reference(orc, msim, cfg, str(tmp_path), 2)
    first, last = metas[0]["owned"][0], metas[0]["owned"][-1]
    assert first > 0.9 * cfg["entities"]  # rank 0 started with (almost) everything
    assert abs(last - cfg["entities"] / 2) < 0.15 * cfg["entities"]  # and ended near half
    assert metas[0]["splits"][0] != metas[0]["splits"][-1]


def test_balanced_splits_properties(msim):
    from movement_sim_b200 import sharding as S

    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for _ in range(20):
            hist = rng.integers(0, 50, 300) * (rng.random(300) < 0.4)
            s = S.balanced_splits(hist, world)
            assert s[0] == 0 and s[-1] == 300 and (np.diff(s) >= 1).all()
            loads = [hist[s[r] : s[r + 1]].sum() for r in range(world)]
            if hist.sum() > 0:
                assert max(loads) <= hist.sum() / world + hist.max() + 1
    s = S.balanced_splits(np.array([0, 0, 100, 0]), 4)  # degenerate: still one row per band
    assert s.tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        S.balanced_splits(np.ones(3), 4)


def test_step_towards_moves_one_row_and_keeps_bands(msim):
    from movement_sim_b200 import sharding as S

    cur, tgt = np.array([0, 10, 11, 40]), np.array([0, 30, 35, 40])
    steps = 0
    while not np.array_equal(cur, tgt):
        nxt = S.step_towards(cur, tgt)
        assert (np.abs(nxt - cur) <= 1).all() and (np.diff(nxt) >= 1).all()
        cur = nxt
        steps += 1
        assert steps < 100
    assert S.entity_range(10, 0, 3) == (0, 3) and S.entity_range(10, 2, 3) == (6, 10)


def test_grid_rows_matches_oracle_side_geometry(msim):
    xy = np.array([[0, 0], [5, 10.02], [100, 16463.7], [3, -4], [1, 1e9]], dtype=np.float32)
    rows, ncx, ncy = msim.grid_rows(29007.4609, 16463.7656, 10.0, xy)
    assert (ncx, ncy) == (2897, 1645)
    assert rows.tolist() == [0, 1, 1644, 0, 1644]
