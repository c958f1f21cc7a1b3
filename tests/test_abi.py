"""The C-ABI library loads on a CPU-only box and exports every symbol include/msim*.h declares; the
PODs have the reference's sizes.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(msim_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(msim):
    L = msim.lib()
    declared = declared_symbols()
    assert declared, "no declarations found"
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"
    assert set(msim.ABI_SYMBOLS) <= declared
    assert L.msim_abi_version() == msim.MSIM_ABI_VERSION


def test_pod_sizes_match_reference_layouts(msim):
    # sim::Entity 64 B, sim::Road 32 B, gpu_quad_tree::Node 64 B, sim::PushConsts 28 B
    assert msim.ENTITY_DTYPE.itemsize == 64
    assert msim.ROAD_DTYPE.itemsize == 32
    assert msim.QUADTREE_NODE_DTYPE.itemsize == 64
    assert C.sizeof(msim.PushConsts) == 28
    off = {n: msim.ENTITY_DTYPE.fields[n][1] for n in msim.ENTITY_DTYPE.names}
    assert off == {"color": 0, "rand_state": 16, "pos": 32, "target": 40, "direction": 48, "road_index": 56, "initialized": 60}


def test_status_strings_and_null_handling(msim):
    L = msim.lib()
    assert L.msim_status_string(0) == b"ok"
    assert L.msim_status_string(msim.MSIM_ERR_CUDA) == b"CUDA error"
    L.msim_destroy(None)  # must be a no-op
    out = C.c_void_p()
    assert L.msim_create(None, C.byref(out)) == msim.MSIM_ERR_INVALID
    assert b"cfg is null" in L.msim_last_error(None)


def test_create_fails_loudly_without_gpu(msim, test_map):
    """There is no CPU fallback: on a box without a B200 msim_create must return MSIM_ERR_CUDA."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(msim.MsimError) as ei:
        msim.Simulation(test_map, test_map.init_entities(10))
    assert ei.value.status == msim.MSIM_ERR_CUDA
    assert "no CPU fallback" in ei.value.message


def test_create_validates_arguments_before_touching_cuda(msim, test_map):
    bad = msim.Map(test_map.width, test_map.height, test_map.roads, np.array([0, 1, 99], dtype=np.uint32))
    with pytest.raises(msim.MsimError) as ei:
        msim.Simulation(bad, bad.init_entities(4))
    assert ei.value.status == msim.MSIM_ERR_INVALID and "not a road index" in ei.value.message
