"""GPU tests of the sharded path (include/msim_shard.h): pack / integrate / ghost kernels against the
unsharded oracle.  Two or three handles on ONE GPU exchange their buffers with device copies (this is
what the driver's single-GPU `pytest -m gpu` run exercises); with >= 2 GPUs the same worker as the CPU
gloo test runs over NCCL."""
import json
import os

import numpy as np
import pytest

from conftest import assert_entities_equal, oracle_map, to_oracle_entities
from test_sharding import BASE, check_against_reference, free_port

pytestmark = pytest.mark.gpu


def run_bands_on_one_gpu(msim, orc, m, total, seed, radius, world, ticks, capacity, box=None, rebalance_every=0, skew=False, asynchronous=False, flags=0, fused=False, p2p=False):
    import torch

    from movement_sim_b200 import sharding as S

    hist, ncx, ncy = S.global_row_histogram(msim, m, total, seed, radius, box)
    splits = np.linspace(0, ncy, world + 1).astype(np.int64) if skew else S.balanced_splits(hist, world)
    # an explicit stream: torch's default stream has handle 0, which the C ABI reads as "create your own stream",
    # and the exchange copies below would then not be ordered with the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        return _run_bands(msim, m, total, seed, radius, world, ticks, capacity, box, rebalance_every, asynchronous, splits, ncy, stream, flags, fused, p2p)


def _run_bands(msim, m, total, seed, radius, world, ticks, capacity, box, rebalance_every, asynchronous, splits, ncy, stream, flags, fused, p2p):
    import torch

    from movement_sim_b200 import sharding as S

    target = splits.copy()
    nbytes = msim.shard_buffer_bytes(capacity, capacity)
    sims, bufs = [], []
    for r in range(world):
        ents, gids = S.collect_band(msim, m, total, seed, radius, int(splits[r]), int(splits[r + 1]), box)
        sim = msim.Simulation(m, ents, radius=radius, stream=stream.cuda_stream, capacity=total + 8 * capacity, flags=flags)
        sim.shard_enable(gids, capacity, capacity)
        sim.dispatch(2)
        sims.append(sim)
        bufs.append({k: torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for k in ("sd", "su", "rd", "ru")})
    if p2p:  # the bands' receive arenas live in this process: connect them by device pointer
        arenas = [sim.shard_p2p_create()[1] for sim in sims]
        for r, sim in enumerate(sims):
            sim.shard_p2p_connect_local(arenas[r - 1] if r > 0 else None, arenas[r + 1] if r + 1 < world else None)
    pairs, owned = [], []
    for t in range(ticks):
        if not np.array_equal(splits, target):
            splits = S.step_towards(splits, target)
        if p2p:
            # every band's move kernel (stores into the neighbours' buffers, raises their flags) is enqueued before any
            # integrate kernel (spins on the flags): on ONE stream the opposite order would wait for itself
            for r, sim in enumerate(sims):
                sim.shard_p2p_move_pack(int(splits[r]), int(splits[r + 1]))
            for sim in sims:
                sim.shard_p2p_integrate()
                sim.enqueue_collide()
            pairs.append(sum(sim.stats()["last_pair_count"] for sim in sims))
            if t % 7 == 0 or t == ticks - 1:
                owned.append([s.stats()["entity_count"] for s in sims])
            if rebalance_every and (t + 1) % rebalance_every == 0:
                h = sum(s.shard_row_histogram(ncy).astype(np.int64) for s in sims)
                target = S.balanced_splits(h, world)
            continue
        for r, sim in enumerate(sims):
            band = (int(splits[r]), int(splits[r + 1]), bufs[r]["sd"].data_ptr() if r > 0 else None, bufs[r]["su"].data_ptr() if r + 1 < world else None)
            if fused:  # one kernel moves and packs; the next-waypoint pass then runs after the integrate
                sim.shard_move_pack(*band)
            else:
                sim.enqueue_move()
                sim.shard_pack(*band)
        for r in range(world):  # the "exchange": what NCCL send/recv does between processes
            if r + 1 < world:
                bufs[r + 1]["rd"].copy_(bufs[r]["su"])
                bufs[r]["ru"].copy_(bufs[r + 1]["sd"])
        total_pairs = 0
        for r, sim in enumerate(sims):
            integrate = sim.shard_integrate_async if asynchronous else sim.shard_integrate
            integrate(bufs[r]["rd"].data_ptr() if r > 0 else None, bufs[r]["ru"].data_ptr() if r + 1 < world else None)
            sim.enqueue_collide()
            if not asynchronous:
                sim.sync()
                total_pairs += sim.stats()["last_pair_count"]
        if asynchronous:  # everything of this tick is enqueued on every band before anybody waits
            total_pairs = sum(sim.stats()["last_pair_count"] for sim in sims)
        pairs.append(total_pairs)
        if not asynchronous or t % 7 == 0 or t == ticks - 1:  # asynchronous ticks: do not force the counts to the host every tick
            owned.append([s.stats()["entity_count"] for s in sims])
        if rebalance_every and (t + 1) % rebalance_every == 0:
            h = sum(s.shard_row_histogram(ncy).astype(np.int64) for s in sims)
            target = S.balanced_splits(h, world)
    if not flags & msim.FLAG_NO_REORDER and ticks >= 8:  # the bands must actually have been re-sorted into cell order
        assert all(s.stats()["reorders"] >= 2 for s in sims), [s.stats()["reorders"] for s in sims]
    got = np.zeros(total, dtype=msim.ENTITY_DTYPE)
    seen = np.zeros(total, dtype=np.int32)
    for sim in sims:
        e, g = sim.read_entities(), sim.shard_read_gids()
        got[g] = e
        seen[g] += 1
        sim.close()
    stream.synchronize()
    assert (seen == 1).all(), (f"gids missing {np.nonzero(seen == 0)[0][:8].tolist()} ({int((seen == 0).sum())}), "
                               f"duplicated {np.nonzero(seen > 1)[0][:8].tolist()} ({int((seen > 1).sum())}); owned history tail {owned[-3:]}")
    return got, pairs, owned


def oracle_reference(msim, orc, m, total, seed, radius, ticks, box=None):
    from movement_sim_b200 import sharding as S

    parts = [e for _, e in S.generate_population(msim, m, total, seed, box)]
    e = to_oracle_entities(orc, np.concatenate(parts))
    om = oracle_map(orc, m)
    orc.move_pass(e, om)
    pairs = []
    for _ in range(ticks):
        orc.move_pass(e, om, threads=8)
        pairs.append(orc.collide_pass(e, m.width, m.height, radius, threads=8))
    return e, pairs


REBUILDS = ["cell-ordered", "onesweep"]


def rebuild_flags(msim, monkeypatch, rebuild):
    """cell-ordered = the default (counting sort + periodic physical re-sort, here every 3rd collision pass so that the
    re-sort of a band is exercised many times); onesweep = the radix-sort rebuild on storage left in arrival order."""
    if rebuild == "cell-ordered":
        monkeypatch.setenv("MSIM_REORDER_EVERY", "3")
        return 0
    return msim.FLAG_SORT_ONESWEEP | msim.FLAG_NO_REORDER


@pytest.mark.parametrize("fused", [False, True], ids=["move+pack", "fused-move-pack"])
@pytest.mark.parametrize("rebuild", REBUILDS)
@pytest.mark.parametrize("asynchronous", [False, True], ids=["host-integrate", "device-integrate"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_bands_on_one_gpu_match_unsharded_oracle(msim, orc, small_city, world, asynchronous, rebuild, fused, monkeypatch):
    total, ticks = 40_000, 50
    got, pairs, owned = run_bands_on_one_gpu(msim, orc, small_city, total, 42, 10.0, world, ticks, capacity=1 << 14, asynchronous=asynchronous,
                                             flags=rebuild_flags(msim, monkeypatch, rebuild), fused=fused)
    want, want_pairs = oracle_reference(msim, orc, small_city, total, 42, 10.0, ticks)
    assert_entities_equal(got, want, what=f"{world} bands")
    assert pairs == want_pairs
    if world > 1:
        assert any(o != owned[0] for o in owned), "entities should migrate between bands"


@pytest.mark.parametrize("fused", [False, True], ids=["move+pack", "fused-move-pack"])
@pytest.mark.parametrize("rebuild", REBUILDS)
@pytest.mark.parametrize("asynchronous", [False, True], ids=["host-integrate", "device-integrate"])
def test_bands_rebalance_dense_corner(msim, orc, small_city, asynchronous, rebuild, fused, monkeypatch):
    """BASELINE config 5 in miniature: everybody starts in one corner, geometric initial split."""
    total, ticks = 30_000, 80
    box = [0.0, 0.0, 900.0, 600.0]
    got, pairs, owned = run_bands_on_one_gpu(msim, orc, small_city, total, 7, 10.0, 2, ticks, capacity=1 << 15, box=box, rebalance_every=4, skew=True,
                                             asynchronous=asynchronous, flags=rebuild_flags(msim, monkeypatch, rebuild), fused=fused)
    want, want_pairs = oracle_reference(msim, orc, small_city, total, 7, 10.0, ticks, box=box)
    assert_entities_equal(got, want, what="rebalanced bands")
    assert pairs == want_pairs
    assert owned[0][0] > 0.8 * total and abs(owned[-1][0] - total / 2) < 0.15 * total


@pytest.mark.parametrize("rebuild", REBUILDS)
@pytest.mark.parametrize("world", [1, 2, 3])
def test_bands_peer_memory_exchange_matches_unsharded_oracle(msim, orc, small_city, world, rebuild, monkeypatch):
    """msim_shard_p2p_*: the move kernel stores into the neighbours' receive buffers, the integrate kernel waits on a flag."""
    total, ticks = 40_000, 50
    got, pairs, owned = run_bands_on_one_gpu(msim, orc, small_city, total, 42, 10.0, world, ticks, capacity=1 << 14, asynchronous=True,
                                             flags=rebuild_flags(msim, monkeypatch, rebuild), p2p=True)
    want, want_pairs = oracle_reference(msim, orc, small_city, total, 42, 10.0, ticks)
    assert_entities_equal(got, want, what=f"{world} bands, peer-memory exchange")
    assert pairs == want_pairs
    if world > 1:
        assert any(o != owned[0] for o in owned), "entities should migrate between bands"


def test_bands_peer_memory_exchange_rebalance_dense_corner(msim, orc, small_city, monkeypatch):
    total, ticks = 30_000, 80
    box = [0.0, 0.0, 900.0, 600.0]
    got, pairs, owned = run_bands_on_one_gpu(msim, orc, small_city, total, 7, 10.0, 2, ticks, capacity=1 << 15, box=box, rebalance_every=4, skew=True,
                                             asynchronous=True, flags=rebuild_flags(msim, monkeypatch, "cell-ordered"), p2p=True)
    want, want_pairs = oracle_reference(msim, orc, small_city, total, 7, 10.0, ticks, box=box)
    assert_entities_equal(got, want, what="rebalanced bands, peer-memory exchange")
    assert pairs == want_pairs


def test_peer_memory_exchange_merged_kernel_on_separate_streams(msim, orc, small_city, monkeypatch):
    """The kernel multi-process ranks run (publish + wait + integrate + ghosts in one launch), exercised on one GPU: each band
    on its own library-owned stream, so that a band's wait can overlap the neighbour's launch."""
    from movement_sim_b200 import sharding as S

    monkeypatch.setenv("MSIM_P2P_MERGED", "1")
    monkeypatch.setenv("MSIM_REORDER_EVERY", "3")
    total, ticks, world, cap = 40_000, 40, 3, 1 << 14
    hist, ncx, ncy = S.global_row_histogram(msim, small_city, total, 42, 10.0)
    splits = S.balanced_splits(hist, world)
    sims = []
    for r in range(world):
        ents, gids = S.collect_band(msim, small_city, total, 42, 10.0, int(splits[r]), int(splits[r + 1]))
        sim = msim.Simulation(small_city, ents, radius=10.0, capacity=total + 8 * cap)  # own stream
        sim.shard_enable(gids, cap, cap)
        sim.dispatch(2)
        sims.append(sim)
    arenas = [s.shard_p2p_create()[1] for s in sims]
    for r, s in enumerate(sims):
        s.shard_p2p_connect_local(arenas[r - 1] if r > 0 else None, arenas[r + 1] if r + 1 < world else None)
    pairs = []
    for t in range(ticks):
        for r, s in enumerate(sims):  # nothing here waits for the GPU: three streams run three bands concurrently
            s.shard_p2p_move_pack(int(splits[r]), int(splits[r + 1]))
            s.shard_p2p_integrate()
            s.enqueue_collide()
        pairs.append(sum(s.stats()["last_pair_count"] for s in sims))
    got = np.zeros(total, dtype=msim.ENTITY_DTYPE)
    seen = np.zeros(total, dtype=np.int32)
    for s in sims:
        e, g = s.read_entities(), s.shard_read_gids()
        got[g] = e
        seen[g] += 1
        s.close()
    assert (seen == 1).all()
    want, want_pairs = oracle_reference(msim, orc, small_city, total, 42, 10.0, ticks)
    assert_entities_equal(got, want, what="3 bands, merged exchange kernel")
    assert pairs == want_pairs


def test_peer_memory_exchange_times_out_instead_of_hanging(msim, small_city, monkeypatch):
    """A neighbour that never enqueues its tick: the integrate kernel gives up after MSIM_P2P_TIMEOUT_MS and the error surfaces."""
    import torch

    from movement_sim_b200 import sharding as S

    monkeypatch.setenv("MSIM_P2P_TIMEOUT_MS", "200")
    total = 20_000
    hist, ncx, ncy = S.global_row_histogram(msim, small_city, total, 42, 10.0)
    sims = []
    for lo, hi in ((0, ncy // 2), (ncy // 2, ncy)):
        ents, gids = S.collect_band(msim, small_city, total, 42, 10.0, lo, hi)
        sim = msim.Simulation(small_city, ents, radius=10.0, capacity=total + (1 << 14))
        sim.shard_enable(gids, 1 << 12, 1 << 12)
        sim.dispatch(2)
        sims.append(sim)
    arenas = [s.shard_p2p_create()[1] for s in sims]
    sims[0].shard_p2p_connect_local(None, arenas[1])
    sims[1].shard_p2p_connect_local(arenas[0], None)
    sims[0].shard_p2p_move_pack(0, ncy // 2)  # band 1 never moves: band 0 waits for a flag that is not coming
    sims[0].shard_p2p_integrate()
    sims[0].enqueue_collide()
    with pytest.raises(msim.MsimError) as ei:
        sims[0].sync()
    assert ei.value.status == msim.MSIM_ERR_INTERNAL and "timeout" in str(ei.value)
    for s in sims:
        s.close()


def test_capacity_overflow_is_reported(msim, small_city):
    import torch

    from movement_sim_b200 import sharding as S

    total = 20_000
    hist, ncx, ncy = S.global_row_histogram(msim, small_city, total, 42, 10.0)
    ents, gids = S.collect_band(msim, small_city, total, 42, 10.0, 0, ncy // 2)
    sim = msim.Simulation(small_city, ents, radius=10.0, capacity=total)  # library-owned stream
    sim.shard_enable(gids, 8, 8)  # absurdly small
    sim.dispatch(2)
    nbytes = msim.shard_buffer_bytes(8, 8)
    up = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    rup = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # the buffers are zeroed on torch's stream
    sim.enqueue_move()
    sim.shard_pack(0, ncy // 2, None, up.data_ptr())
    with pytest.raises(msim.MsimError) as ei:
        sim.shard_integrate(None, rup.data_ptr())
    assert ei.value.status == msim.MSIM_ERR_CAPACITY
    sim.close()


def test_capacity_overflow_is_reported_by_async_ticks(msim, small_city):
    import torch

    from movement_sim_b200 import sharding as S

    total = 20_000
    hist, ncx, ncy = S.global_row_histogram(msim, small_city, total, 42, 10.0)
    ents, gids = S.collect_band(msim, small_city, total, 42, 10.0, 0, ncy // 2)
    sim = msim.Simulation(small_city, ents, radius=10.0, capacity=total)  # library-owned stream
    sim.shard_enable(gids, 8, 8)
    sim.dispatch(2)
    nbytes = msim.shard_buffer_bytes(8, 8)
    up = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    rup = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # the buffers are zeroed on torch's stream
    sim.enqueue_move()
    sim.shard_pack(0, ncy // 2, None, up.data_ptr())
    sim.shard_integrate_async(None, rup.data_ptr())  # enqueue only: cannot fail yet
    sim.enqueue_collide()
    with pytest.raises(msim.MsimError) as ei:
        sim.sync()
    assert ei.value.status == msim.MSIM_ERR_CAPACITY
    sim.close()


@pytest.mark.parametrize("exchange", ["collective", "p2p"])
def test_two_gpus_over_nccl(msim, orc, tmp_path, exchange):
    import json

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from shard_worker import run

    world = min(torch.cuda.device_count(), 4)
    cfg = dict(BASE, entities=60_000, ticks=60, capacity=1 << 14, exchange=exchange)
    mp.spawn(run, args=(world, free_port(), "nccl", str(tmp_path), cfg), nprocs=world, join=True)
    check_against_reference(orc, msim, cfg, str(tmp_path), world)
    with open(tmp_path / "meta_0.json") as f:
        meta = json.load(f)
    assert meta["exchange"] == exchange, f"fell back to {meta['exchange']}: {meta.get('p2p_error')}"
