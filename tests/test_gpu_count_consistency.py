"""The counters a collision pass reports must be the counts of what it stored.

Round 2 found `last_flagged_count` one short of the number of blue flags the same pass had just written, on about every tenth tick at
10 M entities and never at the sizes the oracle comparisons run at (profiles/r2_flag_count_race.md): the count was taken by a ballot
behind divergent code inside the query kernel.  It is now taken from the stored flags by a kernel of its own (csrc/collide.cu
fold_counts_kernel); these tests hold the property at the size where it broke, tick by tick, for both neighbour-structure rebuilds and
for bands on one GPU, and compare two independent handles (different atomic orders inside the cells) with each other."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 10_000_000
TICKS = 600        # two handles in lock step: the count was wrong on every tenth tick
TICKS_BANDS = 1500  # bands against the unsharded handle: a wrong FLAG showed up about once per 700 ticks (both take a few seconds on a B200)


@pytest.fixture(scope="module")
def munich(msim):
    return msim.Map.city()


@pytest.fixture(scope="module")
def population(msim, munich):
    from movement_sim_b200 import sharding as S

    return np.concatenate([e for _, e in S.generate_population(msim, munich, N, 42, None)])


@pytest.mark.parametrize("flags_name", ["default", "FLAG_SORT_ONESWEEP"])
def test_flagged_count_equals_stored_flags_every_tick(msim, munich, population, flags_name):
    flags = 0 if flags_name == "default" else getattr(msim, flags_name)
    ticks = TICKS if flags == 0 else 12
    with msim.Simulation(munich, population, radius=10.0, flags=flags) as a, msim.Simulation(munich, population, radius=10.0, flags=flags) as b:
        for s in (a, b):
            s.dispatch(2)
            s.enqueue_ticks(100, False)  # spread along the roads: boundary cases need neighbours at all distances
        for t in range(ticks):
            a.enqueue_ticks(1, True)
            b.enqueue_ticks(1, True)
            sa, sb = a.stats(), b.stats()
            assert (sa["last_flagged_count"], sa["last_pair_count"]) == (sb["last_flagged_count"], sb["last_pair_count"]), f"tick {t}: two handles, same input"
            if t % 50 == 0 or t == ticks - 1:
                fa = a.read_collision_flags()
                assert int(fa.sum()) == sa["last_flagged_count"], f"tick {t}: count vs stored flags"
        assert a.stats()["total_flagged_count"] == b.stats()["total_flagged_count"] > 0
        assert a.stats()["total_pair_count"] == b.stats()["total_pair_count"] > 0


def test_bands_on_one_gpu_report_the_unsharded_counts_every_tick(msim, munich, population):
    """Four bands (peer-memory exchange between handles of this process, overlapped ticks) against one unsharded handle: flagged and pair
    totals agree on every tick, and the per-band counts are the counts of the flags each band holds."""
    import torch

    from movement_sim_b200 import sharding as S

    world, radius = 4, 10.0
    hist, _, _ = S.global_row_histogram(msim, munich, N, 42, radius, None)
    splits = S.balanced_splits(hist, world)
    cap = max(4096, 3 * int(hist.max()))
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ref = msim.Simulation(munich, population, radius=radius, stream=stream.cuda_stream)
        sims = []
        for r in range(world):
            e, g = S.collect_band(msim, munich, N, 42, radius, int(splits[r]), int(splits[r + 1]), None)
            s = msim.Simulation(munich, e, radius=radius, stream=stream.cuda_stream, capacity=int(e.shape[0] * 1.3) + 8 * cap + 1024)
            s.shard_enable(g, cap, cap)
            sims.append(s)
        arenas = [s.shard_p2p_create()[1] for s in sims]
        for r, s in enumerate(sims):
            s.shard_p2p_connect_local(arenas[r - 1] if r > 0 else None, arenas[r + 1] if r + 1 < world else None)
        for s in [ref] + sims:
            s.dispatch(2)
        for t in range(100 + TICKS_BANDS):
            collide = t >= 100
            ref.enqueue_ticks(1, collide)
            for r, s in enumerate(sims):  # every band's move + pack before any band's integrate (one stream: see test_gpu_sharding)
                s.shard_p2p_move_pack(int(splits[r]), int(splits[r + 1]))
            for s in sims:
                s.shard_p2p_integrate()
                if collide:
                    s.enqueue_collide()
            if not collide:
                continue
            want = ref.stats()
            got = [s.stats() for s in sims]
            assert sum(g["last_pair_count"] for g in got) == want["last_pair_count"], f"tick {t}"
            assert sum(g["last_flagged_count"] for g in got) == want["last_flagged_count"], f"tick {t}"
            if t % 250 == 0:
                for s, g in zip(sims, got):
                    s.shard_counts()
                    assert int(s.read_collision_flags().sum()) == g["last_flagged_count"], f"tick {t}"
        for s in [ref] + sims:
            s.close()
    stream.synchronize()
