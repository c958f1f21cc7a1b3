import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_library_under_emulator():
    """A second copy of the binding whose library is tests/cuda_emu/libmsim_emu.so: the whole library (api.cu, kernels) compiled for the host
    SIMT emulator.  Test infrastructure: the product binding itself has no switch for it."""
    import importlib.util
    import subprocess

    emu_dir = os.path.join(ROOT, "tests", "cuda_emu")
    r = subprocess.run([sys.executable, os.path.join(emu_dir, "build_emu_lib.py")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("the emulator build of the library failed:\n" + r.stdout[-4000:] + r.stderr[-2000:])
    name = "movement_sim_b200_under_emulator"
    if name in sys.modules:
        return sys.modules[name]
    pkg = os.path.join(ROOT, "movement-sim_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    mod.LIB_PATH = os.path.join(emu_dir, "libmsim_emu.so")
    mod.lib()
    return mod


def emulate_torch_cuda():
    """MSIM_TEST_EMULATOR=1: the few torch.cuda pieces the single-process GPU tests use become host stand-ins ("device" memory of the emulated
    library is host memory): streams do nothing, device="cuda" tensors live on the CPU."""
    import contextlib
    import types

    import torch

    if getattr(torch, "_msim_emulated", False):
        return
    torch._msim_emulated = True
    torch.cuda.Stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=1, synchronize=lambda: None)
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: 1
    for name in ("zeros", "empty", "ones", "tensor"):
        real = getattr(torch, name)

        def on_host(*a, _real=real, **k):
            if str(k.get("device", "cpu")).startswith("cuda"):
                k["device"] = "cpu"
            k.pop("pin_memory", None)
            return _real(*a, **k)

        setattr(torch, name, on_host)


@pytest.fixture(scope="session")
def msim():
    """The product binding.  MSIM_TEST_EMULATOR=1 hands the tests the emulated copy instead, which lets `-m gpu` tests run on a box without a
    GPU (a development loop: slow, sizes permitting; it is never what the required GPU run uses)."""
    if os.environ.get("MSIM_TEST_EMULATOR") == "1":
        emulate_torch_cuda()
        return load_library_under_emulator()
    import movement_sim_b200 as M

    M.lib()  # fail loudly if the library was not built
    return M


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def test_map(msim):
    """The reference's 4-road fixture (tests/golden/test_map.json, see make_test_map.py)."""
    return msim.Map.load_json(os.path.join(GOLDEN, "test_map.json"))


@pytest.fixture(scope="session")
def small_city(msim):
    # ~2.2 x 1.6 km street graph, ~5k roads: big enough for every junction degree, small enough for CPU
    return msim.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7)


def to_oracle_entities(O, ents):
    """Product AoS -> oracle AoS (same bytes, different field names)."""
    return np.ascontiguousarray(ents).view(O.ENTITY_DTYPE).copy()


def oracle_map(O, m):
    return O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)


def assert_entities_equal(got, want, fields=None, what=""):
    """Bit-exact comparison of AoS arrays, field by field, with a useful message."""
    got = np.ascontiguousarray(got)
    want = np.ascontiguousarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    gb = got.view(np.uint32).reshape(got.shape[0], 16)
    wb = want.view(np.uint32).reshape(want.shape[0], 16)
    cols = {"color": (0, 4), "rng": (4, 8), "pos": (8, 10), "target": (10, 12), "dir": (12, 14), "road": (14, 15), "initialized": (15, 16)}
    for name, (a, b) in cols.items():
        if fields is not None and name not in fields:
            continue
        bad = np.nonzero((gb[:, a:b] != wb[:, a:b]).any(axis=1))[0]
        if bad.size:
            i = int(bad[0])
            raise AssertionError(
                f"{what}: field '{name}' differs for {bad.size}/{got.shape[0]} entities; first at {i}: "
                f"got {got[i]} want {want[i]}"
            )


ORACLE_THREADS = max(1, min(16, os.cpu_count() or 1))


def oracle_dispatch(O, e, omap, radius, tick, threads=ORACLE_THREADS):
    """One dispatch on the oracle (multi-threaded variant of O.dispatch); returns the unique pair count."""
    if tick % 2 == 0:
        O.move_pass(e, omap, threads=threads)
        return 0
    return O.collide_pass(e, omap.world_w, omap.world_h, radius, threads=threads)
