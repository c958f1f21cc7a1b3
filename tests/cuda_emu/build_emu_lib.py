#!/usr/bin/env python
"""TEST INFRASTRUCTURE: builds the WHOLE library - api.cu, every kernel file, the host map code - for the host SIMT emulator
(tests/cuda_emu/cuda_runtime.h) into tests/cuda_emu/libmsim_emu.so, with the same C ABI as libmsim_cuda.so.

The CUDA sources are used as they are apart from three mechanical rewrites made on a scratch copy (tests/cuda_emu/_gen/):
  kernel<<<grid, block, smem, stream>>>(args)   ->  cuda_emu::cfg(kernel, grid, block, smem, stream)(args)
  extern __shared__ T name[];                    ->  T* name = static_cast<T*>(cuda_emu::dynamic_smem());
  three inline-PTX statements (relaxed load / store, %globaltimer) -> GCC atomics / a steady clock
Nothing of this is linked into or loaded by the product; tests load it explicitly (tests/test_library_under_emulator.py)."""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "movement-sim_b200", "csrc")
GEN = os.path.join(HERE, "_gen")
OUT = os.path.join(HERE, "libmsim_emu.so")

LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;{}()]*>)?)\s*<<<(.+?)>>>\s*\(", re.S)
DYN_SMEM = re.compile(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];")
PTX = {
    'asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");': "v = __atomic_load_n(p, __ATOMIC_RELAXED);",
    'asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");': "__atomic_store_n(p, v, __ATOMIC_RELAXED);",
    'asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));': "t = cuda_emu::now_ns();",
}


def transform(text):
    text = LAUNCH.sub(lambda m: f"cuda_emu::cfg({m.group(1)}, {m.group(2)})(", text)
    text = DYN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = static_cast<{m.group(1)}*>(cuda_emu::dynamic_smem());", text)
    for a, b in PTX.items():
        text = text.replace(a, b)
    return text


def up_to_date():
    outs = [OUT]
    if not all(os.path.exists(o) for o in outs):
        return False
    deps = [os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps += [os.path.join(HERE, n) for n in ("cuda_runtime.h", "build_emu_lib.py")]
    deps += [os.path.join(ROOT, "include", n) for n in os.listdir(os.path.join(ROOT, "include"))]
    return min(os.path.getmtime(o) for o in outs) > max(os.path.getmtime(d) for d in deps)


def main():
    if "--force" not in sys.argv and up_to_date():
        print("up to date")
        return 0
    cxx = os.environ.get("CXX", "g++")
    dst = os.path.join(GEN, "movement-sim_b200", "csrc")
    shutil.rmtree(GEN, ignore_errors=True)
    os.makedirs(dst)
    os.symlink(os.path.join(ROOT, "include"), os.path.join(GEN, "include"))
    sources = []
    for name in sorted(os.listdir(CSRC)):
        src = os.path.join(CSRC, name)
        if name.endswith(".cu"):
            with open(src) as f:
                text = transform(f.read())
            out = os.path.join(dst, name[:-3] + ".cpp")
            with open(out, "w") as f:
                f.write(text)
            sources.append(out)
        elif name.endswith((".h", ".cuh")):
            shutil.copy(src, os.path.join(dst, name))
        elif name.endswith(".cpp"):
            shutil.copy(src, os.path.join(dst, name))
            sources.append(os.path.join(dst, name))
    flags = ["-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fno-fast-math", "-DMSIM_HOST_EMU", "-I", HERE, "-fPIC", "-pthread", "-w"]
    objs = []
    procs = []
    for s in sources:
        o = s[:-4] + ".o"
        objs.append(o)
        procs.append((s, subprocess.Popen([cxx, *flags, "-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            print(f"== {os.path.basename(s)}\n{out[-6000:]}")
    if failed:
        return 1
    emu_globals = os.path.join(GEN, "emu_globals.cpp")
    with open(emu_globals, "w") as f:
        f.write('#include "cuda_runtime.h"\nthread_local uint3 threadIdx, blockIdx;\nthread_local dim3 blockDim, gridDim;\n'
                "namespace cuda_emu {\nthread_local Block* block = nullptr;\nthread_local unsigned lane = 0, warp = 0;\nthread_local void* dynamic_smem_ptr = nullptr;\n}\n")
    subprocess.check_call([cxx, *flags, "-shared", "-o", OUT, emu_globals, *objs])
    print(OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
