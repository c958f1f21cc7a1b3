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


def run_bands_on_one_gpu(msim, orc, m, total, seed, radius, world, ticks, capacity, box=None, rebalance_every=0, skew=False, asynchronous=False, flags=0, fused=False):
    import torch

    from movement_sim_b200 import sharding as S

    hist, ncx, ncy = S.global_row_histogram(msim, m, total, seed, radius, box)
    splits = np.linspace(0, ncy, world + 1).astype(np.int64) if skew else S.balanced_splits(hist, world)
    # an explicit stream: torch's default stream has handle 0, which the C ABI reads as "create your own stream",
    # and the exchange copies below would then not be ordered with the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        return _run_bands(msim, m, total, seed, radius, world, ticks, capacity, box, rebalance_every, asynchronous, splits, ncy, stream, flags, fused)


def _run_bands(msim, m, total, seed, radius, world, ticks, capacity, box, rebalance_every, asynchronous, splits, ncy, stream, flags, fused):
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
    pairs, owned = [], []
    for t in range(ticks):
        if not np.array_equal(splits, target):
            splits = S.step_towards(splits, target)
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


def test_two_gpus_over_nccl(msim, orc, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from shard_worker import run

    world = min(torch.cuda.device_count(), 4)
    cfg = dict(BASE, entities=60_000, ticks=60, capacity=1 << 14)
    mp.spawn(run, args=(world, free_port(), "nccl", str(tmp_path), cfg), nprocs=world, join=True)
    check_against_reference(orc, msim, cfg, str(tmp_path), world)
