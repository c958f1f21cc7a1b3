"""The two numerical assumptions the CUDA collision path rests on, checked on the host (include/msim_shard.h: msim_grid_params,
msim_grid_rows compute exactly what the device computes):
  1. the query tests d2 < T instead of sqrtf(d2) < r (random_move.comp:551-562): T must make the two predicates identical for EVERY binary32 d2;
  2. candidates are taken from the 3 x 3 cells around an entity (and the paired kernel tests candidates of neighbouring runs without a range
     check): two points the predicate accepts must never be more than one cell apart in either axis, in spite of the binary32 rounding of
     pos * inv_cell at coordinates up to 82 km."""
import numpy as np
import pytest

f32 = np.float32
WORLDS = [(29007.4609, 16463.7656), (81900.0, 81900.0), (100.0, 100.0), (2200.0, 1600.0), (500000.0, 300.0)]
RADII = [10.0, 3.3, 0.1, 1.0, 25.0, 0.37, 123.456]


@pytest.mark.parametrize("radius", RADII + [1e-3, 12345.678, 2.0 ** -20])
def test_squared_threshold_is_exactly_the_sqrt_predicate(msim, radius):
    r = f32(radius)
    T = msim.grid_params(1000.0, 1000.0, radius)["hit_threshold"]
    assert T.dtype == np.float32
    # every binary32 value within 4096 ulps of T, T itself, and a spread of random magnitudes
    bits = np.arange(-4096, 4097, dtype=np.int64) + int(T.view(np.uint32))
    near = bits[(bits > 0)].astype(np.uint32).view(np.float32)
    rnd = np.random.default_rng(1)
    far = (f32(radius) ** 2 * rnd.uniform(0.0, 4.0, 200_000)).astype(np.float32)
    for d2 in (near, far, np.array([0.0, np.inf, T], dtype=np.float32)):
        assert np.array_equal(d2 < T, np.sqrt(d2) < r)


def device_cells(xy, inv_cell, ncx, ncy):
    """cell_key_of (csrc/msim_internal.h): one binary32 multiply per axis, floor, clamp."""
    cx = np.clip(np.floor(xy[:, 0].astype(np.float32) * inv_cell), 0, ncx - 1).astype(np.int64)
    cy = np.clip(np.floor(xy[:, 1].astype(np.float32) * inv_cell), 0, ncy - 1).astype(np.int64)
    return cx, cy


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("radius", RADII)
def test_points_in_range_are_in_adjacent_cells(msim, world, radius):
    w, h = world
    g = msim.grid_params(w, h, radius)
    inv, T, ncx, ncy = g["inv_cell"], g["hit_threshold"], g["cells_x"], g["cells_y"]
    cell = 1.0 / float(inv)
    assert cell > radius  # the edge lies above the radius
    rnd = np.random.default_rng(int(w) + int(radius * 1000))
    n = 400_000
    # first points: half of them uniform, half within a few ulps of a cell boundary (where rounding of pos * inv_cell can go either way)
    ax = rnd.uniform(0, w, n)
    ay = rnd.uniform(0, h, n)
    k = n // 2
    bx_cells = rnd.integers(1, max(2, ncx), k)
    by_cells = rnd.integers(1, max(2, ncy), k)
    ax[:k] = np.minimum(bx_cells * cell * (1.0 + rnd.uniform(-3e-7, 3e-7, k)), w)
    ay[:k] = np.minimum(by_cells * cell * (1.0 + rnd.uniform(-3e-7, 3e-7, k)), h)
    a = np.stack([ax, ay], axis=1).astype(np.float32)
    # second points: at distance just below / at / just above the radius, all directions, clipped to the world
    ang = rnd.uniform(0, 2 * np.pi, n)
    ang[: n // 8] = rnd.choice([0.0, np.pi / 2, np.pi, 3 * np.pi / 2], n // 8)  # axis-aligned pairs stress one axis fully
    dist = radius * (1.0 + rnd.uniform(-4e-7, 4e-7, n))
    dist[n // 2:] = radius * rnd.uniform(0.0, 1.0, n - n // 2)
    b = np.stack([np.clip(a[:, 0].astype(np.float64) + dist * np.cos(ang), 0, w), np.clip(a[:, 1].astype(np.float64) + dist * np.sin(ang), 0, h)],
                 axis=1).astype(np.float32)
    # the device's predicate: individually rounded binary32 operations (collide.cu: dist2)
    dx = (b[:, 0] - a[:, 0]).astype(np.float32)
    dy = (b[:, 1] - a[:, 1]).astype(np.float32)
    d2 = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32)
    hit = d2 < T
    assert hit.sum() > n // 4 and (~hit).sum() > 1000  # both sides of the radius are exercised
    acx, acy = device_cells(a, inv, ncx, ncy)
    bcx, bcy = device_cells(b, inv, ncx, ncy)
    rows, gx, gy = msim.grid_rows(w, h, radius, a)  # the library's own host copy of the device formula
    assert (gx, gy) == (ncx, ncy) and np.array_equal(rows.astype(np.int64), acy)
    far = hit & ((np.abs(acx - bcx) > 1) | (np.abs(acy - bcy) > 1))
    assert not far.any(), f"{int(far.sum())} accepted pairs lie more than one cell apart, e.g. {a[far][0]} {b[far][0]}"
