"""The default kernels must still be the machine code that last ran the GPU suite and the bench (commit VERIFIED below): everything
added since without a GPU is an extra template instantiation or an extra translation unit behind an opt-in flag / environment knob.
profiles/sass_identity.py recompiles that commit's CUDA sources and compares instruction streams; needs nvcc, cuobjdump and the git
history (skipped on the GPU box, where the repository travels without .git).  Update VERIFIED when a new build has been through a GPU."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

VERIFIED = "bf013ef"


@pytest.mark.timeout(900)
def test_default_kernels_are_the_gpu_verified_machine_code():
    if not (shutil.which("nvcc") and shutil.which("cuobjdump")):
        pytest.skip("CUDA toolchain not on PATH")
    if subprocess.run(["git", "-C", ROOT, "cat-file", "-e", VERIFIED], capture_output=True).returncode != 0:
        pytest.skip("git history not available")
    if not os.path.exists(os.path.join(ROOT, "movement-sim_b200", "csrc", "move.o")):
        pytest.skip("objects not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sass_identity.py"), VERIFIED], capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stderr[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.endswith(" 0 changed") and not last.startswith("0 kernels"), r.stdout[-3000:]
