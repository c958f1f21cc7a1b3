#!/usr/bin/env python
"""Is the machine code of the default kernels still the code that was last verified on a GPU?

    python profiles/sass_identity.py bf013ef          # the last commit whose build ran the GPU suite and the bench

Compiles every csrc/*.cu of that commit with the Makefile's flags into a scratch directory, disassembles old and current
objects (cuobjdump -sass) and looks, for every kernel of the old build, for a kernel of the current build with the same base
name and the same instruction stream (addresses and encodings stripped).  Kernels that gained template parameters (opt-in
variants) match through their default instantiation.  Needs no GPU.  Prints one line per old kernel."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "movement-sim_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xcompiler", "-fPIC"]


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            ln = re.sub(r"/\*[0-9a-f]{4}\*/", "", ln)
            ln = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln)
            funcs[cur].append(re.sub(r"\s+", " ", ln.strip()))
    return funcs


def base_name(mangled):
    d = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    m = re.search(r"(\w+_kernel)\b", d)
    return m.group(1) if m else d


def main():
    rev = sys.argv[1]
    files = subprocess.run(["git", "-C", ROOT, "ls-tree", "--name-only", rev, "movement-sim_b200/csrc/"], capture_output=True, text=True,
                           check=True).stdout.split()
    same = differ = 0
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "movement-sim_b200", "csrc"))
        os.makedirs(os.path.join(tmp, "include"))
        for f in files + subprocess.run(["git", "-C", ROOT, "ls-tree", "--name-only", rev, "include/"], capture_output=True, text=True).stdout.split():
            if f.endswith((".cu", ".h")):
                with open(os.path.join(tmp, f), "w") as out:
                    out.write(subprocess.run(["git", "-C", ROOT, "show", f"{rev}:{f}"], capture_output=True, text=True, check=True).stdout)
        from concurrent.futures import ThreadPoolExecutor

        def one(f):
            name = os.path.basename(f)[:-3]
            new_obj = os.path.join(CSRC, name + ".o")
            if not os.path.exists(new_obj):
                return name, None, None
            old_obj = os.path.join(tmp, name + ".o")
            subprocess.run(["nvcc", *FLAGS, "-c", os.path.join(tmp, f), "-o", old_obj], check=True, capture_output=True)
            return name, sass(old_obj), sass(new_obj)

        with ThreadPoolExecutor(max_workers=8) as pool:  # nvcc and cuobjdump are subprocesses: the files compile side by side
            results = list(pool.map(one, sorted(f for f in files if f.endswith(".cu"))))
        for name, old, new in results:
            if old is None:
                print(f"{name}: no current object (build first)")
                continue
            new_by_base = {}
            for k, v in new.items():
                new_by_base.setdefault(base_name(k), []).append(v)
            for k, v in old.items():
                b = base_name(k)
                ok = any(v == w for w in new_by_base.get(b, []))
                same += ok
                differ += not ok
                print(f"{name:9s} {b:34s} {len(v):5d} instructions  {'identical' if ok else 'CHANGED'}")
    print(f"{same} kernels identical to {rev}, {differ} changed")
    return 0


if __name__ == "__main__":
    sys.exit(main())
