"""Build recipe of libmicropp_b200.so (CUDA kernels + C++ host + C ABI) for sm_100a.

Run as ``python -m micropp_b200.build`` or through ``__graft_entry__.build()``.  The library is
built IN-TREE (micropp_b200/libmicropp_b200.so: git-ignored, but it travels to the GPU box with
the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libmicropp_b200.so"

SOURCES = ["mgpu_kernels.cu", "spmv_implicit.cu", "cg_resident.cu", "ell_generic.cu", "micropp_host.cpp", "micropp_geometry.cpp", "material_host.cpp", "ell_host.cpp", "slab_host.cpp",
           "micropp_c.cpp"]
HEADERS = list((ROOT / "include").glob("*.h*")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-fvisibility=default",
              "-I", str(ROOT / "include"), "-I", str(CSRC)]


def nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = objdir / (Path(s).stem + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-x", "cu", "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {s}")
        if verbose and out:
            print(out)
    link = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs,
            "-cudart", "static", "-Xlinker", "-Bsymbolic"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
