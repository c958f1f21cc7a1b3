"""The kernels written after the round's GPU budget was spent, EXECUTED on the CPU under a host SIMT emulator (tests/cuda_emu: one OS thread
per CUDA thread, barriers for __syncthreads and for the warp collectives, GCC atomics) and compared with the oracle:
  * move_kernel<FUSE> - MSIM_FLAG_FUSED_ARRIVE: the pending next-waypoint pass served inside the following move kernel (plus the unfused
    kernel + arrive_kernel as the control, i.e. the verified path under the same emulator);
  * arrive_kernel<STRIDE> - the strided pass B of MSIM_ARRIVE_GRID / MSIM_ARRIVE_BESIDE_CTAS;
  * query_paired_kernel - MSIM_QUERY_PAIRED, tiled (shared-memory windows) and untiled (window overflow) CTAs.
The kernel sources are compiled as they are (only the <<<>>> launchers and five inline-PTX helpers have emulator twins).  This shows that
the kernels' logic - indices, masks, warp votes, window offsets, counts - is right; it says nothing about timing or about hardware effects
an emulator does not have.  Hardware parity remains tests/test_zz_gpu_unverified.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, oracle_map, to_oracle_entities

EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")
f32 = np.float32


@pytest.fixture(scope="module")
def emu():
    import sys

    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "build_emu_lib.py")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("the emulator build failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    L = C.CDLL(os.path.join(EMU_DIR, "libemu_kernels.so"))
    vp, u32, u64, i32, f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
    L.emu_move.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, vp, u64, i32, i32, u32]
    L.emu_arrive.argtypes = [u32, vp, vp, vp, vp, vp, vp, u64, i32, u32]
    L.emu_query_paired.argtypes = [u32, vp, vp, vp, f, f, f, i32, i32, vp]
    L.emu_query_window.restype = u32
    for fn in (L.emu_move, L.emu_arrive, L.emu_query_paired):
        fn.restype = None
    return L


class SoA:
    """The resident state as api.cu lays it out (capacity padded to 64 entities, ping-pong positions, 1 arrival bit per entity)."""

    def __init__(self, e, m):
        n = e.shape[0]
        cap = (n + 63) // 64 * 64
        self.n, self.cur = n, 0
        self.pos = [np.zeros((cap, 2), f32), np.zeros((cap, 2), f32)]
        self.pos[0][:n] = e["pos"]
        self.target = np.zeros((cap, 2), f32)
        self.target[:n] = e["target"]
        self.road = np.zeros(cap, np.uint32)
        self.road[:n] = e["road"]
        self.rng = np.zeros((cap, 4), np.uint32)
        self.rng[:n] = e["rng"]
        self.arrived = np.zeros(cap // 32 + 2, np.uint32)
        self.roads = np.ascontiguousarray(m.roads)
        self.conn = np.ascontiguousarray(m.connections)

    def move(self, L, fuse, consume, blocks):
        L.emu_move(self.n, self.pos[self.cur].ctypes.data, self.pos[self.cur ^ 1].ctypes.data, self.target.ctypes.data, self.arrived.ctypes.data,
                   self.road.ctypes.data, self.rng.ctypes.data, self.roads.ctypes.data, self.conn.ctypes.data, self.conn.shape[0], fuse, consume, blocks)
        self.cur ^= 1

    def arrive(self, L, stride=0, blocks=0):
        L.emu_arrive(self.n, self.arrived.ctypes.data, self.target.ctypes.data, self.road.ctypes.data, self.rng.ctypes.data, self.roads.ctypes.data,
                     self.conn.ctypes.data, self.conn.shape[0], stride, blocks)

    def check(self, want, what):
        n = self.n
        for name, got, exp in (("pos", self.pos[self.cur][:n], want["pos"]), ("target", self.target[:n], want["target"]),
                               ("road", self.road[:n], want["road"]), ("rng", self.rng[:n], want["rng"])):
            bad = np.nonzero((got.view(np.uint32).reshape(n, -1) != exp.view(np.uint32).reshape(n, -1)).any(axis=1))[0]
            assert bad.size == 0, f"{what}: {name} differs for {bad.size} of {n} entities, first {int(bad[0])}: {got[bad[0]]} vs {exp[bad[0]]}"


def initialised(orc, m, n, seed):
    om = oracle_map(orc, m)
    e = to_oracle_entities(orc, m.init_entities(n, seed=seed))
    orc.move_pass(e, om)  # the init-only dispatch (host logic in api.cu, not a kernel)
    return e, om


@pytest.mark.timeout(900)
@pytest.mark.parametrize("n,blocks", [(1, 1), (65, 1), (1500, 2), (3001, 1)])
def test_fused_move_kernel_under_the_emulator(emu, orc, test_map, n, blocks):
    """test_map: 63.64 m roads, an arrival every ~46 passes per entity, 4-way junctions (RNG draws) and the past-the-end connection read."""
    e, om = initialised(orc, test_map, n, 100 + n)
    s = SoA(e, test_map)
    for t in range(100):
        s.move(emu, 1, 1 if t else 0, blocks)  # the first pass has nothing pending
        orc.move_pass(e, om)
        if t in (0, 45, 46, 47, 99):  # what a synchronising call does: complete the pending pass with the stand-alone kernel, then look
            probe = SoA.__new__(SoA)
            probe.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else [a.copy() for a in v] if isinstance(v, list) else v) for k, v in s.__dict__.items()})
            probe.arrive(emu)
            probe.check(e, f"n={n}: fused pass {t + 1}")
    s.arrive(emu)
    s.check(e, f"n={n}: after 100 fused passes")


@pytest.mark.timeout(900)
def test_unfused_kernels_and_strided_pass_b_under_the_emulator(emu, orc, small_city):
    """The verified pair (move_kernel, arrive_kernel) as the control, and arrive_kernel<STRIDE> with a grid far smaller than the mask."""
    e, om = initialised(orc, small_city, 12_000, 5)
    plain, strided = SoA(e, small_city), SoA(e, small_city)
    for t in range(24):
        plain.move(emu, 0, 0, 3)
        plain.arrive(emu)
        strided.move(emu, 0, 0, 3)
        strided.arrive(emu, stride=1, blocks=1)  # 12 000 entities = 376 mask words = 2 CTAs' worth: one CTA strides over them
        orc.move_pass(e, om, threads=4)
    plain.check(e, "move + arrive")
    strided.check(e, "move + strided arrive")


def sorted_view(msim, pos, w, h, radius):
    g = msim.grid_params(w, h, radius)
    inv, ncx, ncy = g["inv_cell"], g["cells_x"], g["cells_y"]
    cx = np.clip(np.floor(pos[:, 0] * inv), 0, ncx - 1).astype(np.int64)
    cy = np.clip(np.floor(pos[:, 1] * inv), 0, ncy - 1).astype(np.int64)
    key = cy * ncx + cx
    order = np.argsort(key, kind="stable")
    cap = (pos.shape[0] + 63) // 64 * 64
    sp = np.zeros((cap, 2), f32)
    sp[: pos.shape[0]] = pos[order]
    cell_start = np.searchsorted(key[order], np.arange(ncx * ncy + 1)).astype(np.uint32)
    return g, order, sp, cell_start


@pytest.mark.timeout(900)
@pytest.mark.parametrize("case", ["city", "city_odd_small_radius", "dense_untiled", "three"])
def test_paired_query_kernel_under_the_emulator(emu, msim, orc, small_city, test_map, case):
    m, n, radius, passes = {"city": (small_city, 12_000, 10.0, 150), "city_odd_small_radius": (small_city, 5001, 3.3, 60),
                            "dense_untiled": (test_map, 6000, 10.0, 90), "three": (small_city, 3, 10.0, 2)}[case]
    om = oracle_map(orc, m)
    e = to_oracle_entities(orc, m.init_entities(n, seed=31))
    for _ in range(passes):
        orc.move_pass(e, om, threads=4)
    want_pairs = orc.collide_pass(e, om.world_w, om.world_h, radius, threads=4)
    want_flags = orc.collision_flags(e)
    g, order, sp, cell_start = sorted_view(msim, e["pos"].astype(f32), m.width, m.height, radius)
    flags_sorted = np.full(sp.shape[0], 7, np.uint8)
    stripes = np.zeros(64 * 16, np.uint64)
    emu.emu_query_paired(n, sp.ctypes.data, cell_start.ctypes.data, flags_sorted.ctypes.data, g["inv_cell"], g["hit_threshold"], f32(radius), g["cells_x"],
                         g["cells_y"], stripes.ctypes.data)
    got_flags = np.zeros(n, np.uint8)
    got_flags[order] = flags_sorted[:n]
    assert int(stripes[1::16].sum()) == want_pairs
    assert int(stripes[0::16].sum()) == int(want_flags.sum())
    assert np.array_equal(got_flags, want_flags)
    if case == "dense_untiled":  # 6000 entities on four roads: the 3-cell runs exceed the shared-memory window, the CTAs take the global-memory scans
        assert np.diff(cell_start.astype(np.int64)).max() * 2 > emu.emu_query_window()


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("fuse", [0, 1])
def test_whole_collision_tick_under_the_emulator(emu, msim, orc, small_city, fuse):
    """A complete sim tick of the counting-sort path, kernel by kernel as api.cu enqueues it: move_kernel<keys> with the rank atomics fused in
    (with and without the fused pass B) -> scan_tile_sums -> scan_tiles (both register variants, alternating) -> cell_scatter ->
    query_paired_kernel, every tick against the oracle's move and collision passes: positions, waypoints, roads, RNG states, colour flags and
    the unique-pair count."""
    vp, u32, u64, i32, f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
    emu.emu_move_keys.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, vp, u64, i32, i32, u32, vp, vp, vp, f, i32, i32]
    emu.emu_move_keys.restype = None
    emu.emu_scan_scatter.argtypes = [u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, i32, u32]
    emu.emu_scan_scatter.restype = None
    n, radius = 6000, 10.0
    e, om = initialised(orc, small_city, n, 77)
    for _ in range(60):  # leave the stacked start behind
        orc.move_pass(e, om, threads=4)
    s = SoA(e, small_city)
    g = msim.grid_params(small_city.width, small_city.height, radius)
    ncx, ncy = g["cells_x"], g["cells_y"]
    cells = ncx * ncy
    cap = s.pos[0].shape[0]
    keys, rank = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    cell_count, cell_start = np.zeros(cells + 1, np.uint32), np.zeros(cells + 1, np.uint32)
    tile_sums = np.zeros(cells // 4096 + 2, np.uint32)
    sorted_pos, sorted_idx = np.zeros((cap, 2), f32), np.zeros(cap, np.uint32)
    for t in range(8):
        emu.emu_move_keys(n, s.pos[s.cur].ctypes.data, s.pos[s.cur ^ 1].ctypes.data, s.target.ctypes.data, s.arrived.ctypes.data, s.road.ctypes.data,
                          s.rng.ctypes.data, s.roads.ctypes.data, s.conn.ctypes.data, s.conn.shape[0], fuse, 1 if (fuse and t) else 0, 2,
                          keys.ctypes.data, cell_count.ctypes.data, rank.ctypes.data, g["inv_cell"], ncx, ncy)
        s.cur ^= 1
        if not fuse:
            s.arrive(emu)
        orc.move_pass(e, om, threads=4)
        want_pairs = orc.collide_pass(e, om.world_w, om.world_h, radius, threads=4)
        emu.emu_scan_scatter(n, cells, cell_count.ctypes.data, tile_sums.ctypes.data, cell_start.ctypes.data, keys.ctypes.data, rank.ctypes.data,
                             s.pos[s.cur].ctypes.data, sorted_pos.ctypes.data, sorted_idx.ctypes.data, t % 2, 3)
        assert not cell_count[:cells].any(), "the scan leaves the counters zeroed for the next tick"
        assert int(cell_start[cells]) == n and np.all(np.diff(cell_start.astype(np.int64)) >= 0)
        assert np.array_equal(np.sort(sorted_idx[:n]), np.arange(n, dtype=np.uint32))  # a permutation
        assert np.array_equal(sorted_pos[:n], s.pos[s.cur][sorted_idx[:n]])
        assert np.all(np.diff(keys[sorted_idx[:n]].astype(np.int64)) >= 0), "slots are in cell order"
        flags_sorted = np.full(cap, 9, np.uint8)
        stripes = np.zeros(64 * 16, np.uint64)
        emu.emu_query_paired(n, sorted_pos.ctypes.data, cell_start.ctypes.data, flags_sorted.ctypes.data, g["inv_cell"], g["hit_threshold"], f32(radius), ncx,
                             ncy, stripes.ctypes.data)
        got_flags = np.zeros(n, np.uint8)
        got_flags[sorted_idx[:n]] = flags_sorted[:n]
        assert int(stripes[1::16].sum()) == want_pairs, f"tick {t}"
        assert np.array_equal(got_flags, orc.collision_flags(e)), f"tick {t}"
        if not fuse:
            s.check(e, f"tick {t}")
    if fuse:
        s.arrive(emu)  # complete the pending pass before looking
    s.check(e, "after 8 ticks")
