"""Dry run of the gated GPU tests' OWN logic (tests/test_zz_gpu_unverified.py) on the CPU: the library's Simulation class is replaced by a
stand-in that computes with the oracle, so tick numbering, oracle bookkeeping, expected pair counts and buffer-lifetime expectations of
those tests are exercised before they cost GPU time.  It proves nothing about the kernels: it proves that a failure of a gated test on
hardware will be the kernels' fault, not the test's.  (Tests that start a subprocess of the real library are not covered.)"""
import types

import numpy as np
import pytest

import test_zz_gpu_unverified as gated
from conftest import oracle_map, to_oracle_entities


class OracleBackedSimulation:
    """The subset of movement_sim_b200.Simulation the gated tests use, with the dispatch rules of the C ABI (first dispatch after an upload of
    uninitialised entities only initialises; even tick = move pass, odd tick = collision pass; enqueue_ticks = (move [, collide]) per tick)."""

    O = None

    def __init__(self, m, entities, radius=10.0, flags=0, **_):
        O = self.O
        self.om = oracle_map(O, m)
        self.e = to_oracle_entities(O, entities)
        self.radius, self.flags = float(radius), flags
        self.count = self.e.shape[0]
        self.launches = 0
        self.last_pairs = 0
        self.total_pairs = 0
        self.snap = [None, None]
        self.slot = 0
        self.pending = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def _move(self):
        self.O.move_pass(self.e, self.om)
        self.launches += 1

    def _collide(self):
        self.last_pairs = self.O.collide_pass(self.e, self.om.world_w, self.om.world_h, self.radius)
        self.total_pairs += self.last_pairs

    def dispatch(self, tick):
        if tick % 2 == 0 or (self.count and not self.e["initialized"].all()):
            self._move()  # the oracle's move pass is the init-only dispatch for uninitialised entities
        else:
            self._collide()

    def enqueue_ticks(self, k, collide):
        for _ in range(k):
            self._move()
            if collide:
                self._collide()

    def sync(self):
        pass

    def stats(self):
        self.launches += 1 if (self.flags & gated_fused_flag()) else 0  # the stand-alone pass B a synchronising call completes
        return {"last_pair_count": self.last_pairs, "kernel_launches": self.launches, "entity_count": self.count}

    def _product(self, e):
        import movement_sim_b200 as M

        return e.view(M.ENTITY_DTYPE)  # same 64 bytes, the product's field names

    def read_entities(self):
        return self._product(self.e.copy())

    def read_collision_flags(self):
        return self.O.collision_flags(self.e)

    def read_debug(self):
        out = np.zeros(10, dtype=np.uint32)
        out[0], out[1] = self.count, self.total_pairs
        return out

    def snapshot_begin(self):
        self.slot ^= 1
        self.snap[self.slot] = self._product(self.e.copy())
        self.pending = True

    def snapshot_ready(self):
        if not self.pending:
            raise RuntimeError("no snapshot")
        return True

    def snapshot_end(self, copy=True):
        if not self.pending:
            import movement_sim_b200 as M

            raise M.MsimError(M.MSIM_ERR_INVALID, "msim_snapshot_end: no snapshot has been started")
        self.pending = False
        return self.snap[self.slot].copy() if copy else self.snap[self.slot]


def gated_fused_flag():
    import movement_sim_b200 as M

    return M.FLAG_FUSED_ARRIVE


@pytest.fixture
def fake(msim, orc, monkeypatch):
    OracleBackedSimulation.O = orc
    stand_in = types.SimpleNamespace(**{k: getattr(msim, k) for k in dir(msim) if not k.startswith("__")})
    stand_in.Simulation = OracleBackedSimulation
    return stand_in


@pytest.mark.parametrize("n", [1, 65, 4097])
def test_dry_run_fused_arrive_collisions_off(fake, orc, test_map, n):
    gated.test_fused_arrive_collisions_off(fake, orc, test_map, n)


def test_dry_run_fused_arrive_with_collisions(fake, orc, small_city, monkeypatch):
    gated.test_fused_arrive_with_collisions(fake, orc, small_city, "default", monkeypatch)


def test_dry_run_blocking_dispatch(fake, orc, small_city):
    gated.test_fused_arrive_blocking_dispatch_is_unchanged(fake, orc, small_city)


def test_dry_run_snapshots(fake, orc, small_city, test_map):
    gated.test_snapshot_is_the_state_at_begin(fake, orc, small_city)
    gated.test_snapshot_argument_errors(fake, test_map)


def test_dry_run_of_the_required_tests_against_the_compiled_shader(fake, orc, small_city):
    """The two tests added to the REQUIRED GPU suite after the last GPU run (tests/test_gpu_parity.py): their own logic, on the CPU."""
    import test_gpu_parity as parity

    fake.Map = __import__("movement_sim_b200").Map
    parity.test_cuda_path_matches_compiled_shader(fake, orc, small_city)
    parity.test_cuda_path_matches_the_whole_compiled_shader(fake, orc, small_city)
