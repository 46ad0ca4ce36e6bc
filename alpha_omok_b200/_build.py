"""In-tree build of the CUDA C-ABI library (sm_100a only) and of the C oracle used by the tests.

`build()` is what `__graft_entry__.build()` calls. nvcc cross-compiles without a GPU.
The library is written next to this file so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libalpha_omok_b200.so")
PROBE_LIB_PATH = os.path.join(HERE, "libalpha_omok_b200_probe.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _sources(probe=False):
    src = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    if probe:
        pdir = os.path.join(CSRC, "probe")
        src += sorted(os.path.join(pdir, f) for f in os.listdir(pdir) if f.endswith(".cu"))
    return src


def _digest(probe=False):
    h = hashlib.sha256()
    dirs = [CSRC] + ([os.path.join(CSRC, "probe")] if probe else [])
    for d in dirs:
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    inc = os.path.join(os.path.dirname(HERE), "include")
    for f in sorted(os.listdir(inc)) if os.path.isdir(inc) else []:
        with open(os.path.join(inc, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(b"probe" if probe else b"product")
    return h.hexdigest()


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA C-ABI library cannot be built")
    return p


def build(force: bool = False, verbose: bool = False, probe: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library for sm_100a. Returns its path.

    probe=False: the product library libalpha_omok_b200.so (include/alpha_omok_b200.h).
    probe=True : libalpha_omok_b200_probe.so = the same sources with -DAO_PROBE (tower cycle counters, AO_TOWER_XFLAGS
                 timing experiments) + csrc/probe/*.cu (tcgen05 probes) - test / profiling tooling only
                 (include/alpha_omok_b200_probe.h)."""
    lib_path = PROBE_LIB_PATH if probe else LIB_PATH
    stamp = lib_path + ".stamp"
    dig = _digest(probe)
    if not force and os.path.exists(lib_path) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == dig:
                return lib_path
    inc = os.path.join(os.path.dirname(HERE), "include")
    objdir = os.path.join(HERE, "build", "probe" if probe else "product")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + (["-DAO_PROBE"] if probe else [])
    nvcc = nvcc_path()
    srcs = _sources(probe)
    procs, objs, log = [], [], ""
    for src in srcs:  # one nvcc per translation unit, all in parallel
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *flags, "-c", "-I", inc, "-I", CSRC, "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, pr in procs:
        out, _ = pr.communicate()
        log += " ".join(cmd) + "\n" + out
        failed |= pr.returncode != 0
    if not failed:
        cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(os.path.join(HERE, "build_probe.log" if probe else "build.log"), "w") as fh:
        fh.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building " + os.path.basename(lib_path))
    if verbose:
        print(log)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, probe="--probe" in sys.argv))
