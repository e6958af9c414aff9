"""Builds repas_vision_b200/librepasvision.so from csrc/*.cu with nvcc for sm_100a.

In-tree build (the .so is git-ignored but travels with the working tree), explicit flags:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only, no other arch, no PTX fallback
  --fmad=false                              one rounding per float op (bit-exact parity with numpy / C)
  -lineinfo                                 so ncu's source page maps to csrc/
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.path.join(HERE, "librepasvision.so")
SOURCES = ["rv_abi.cu", "rv_deproject.cu", "rv_deproject_tma.cu", "rv_register.cu", "rv_cloud.cu", "rv_neighbors.cu",
           "rv_misc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a extension")


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "repas_vision.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra: str | None = None, tag: str = "obj") -> str:
    """out / extra / tag: a variant build (other -D tunables, other output file, its own object directory)."""
    if out is None and not force and not needs_build():
        return SO_PATH
    nvcc = _nvcc()
    so_path = out or SO_PATH
    objdir = os.path.join(HERE, "..", "build", tag)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(extra if extra is not None else os.environ.get("RV_NVCC_EXTRA", "")).split(), "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
    tmp = so_path + ".tmp"
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    subprocess.check_call(link)
    os.replace(tmp, so_path)
    return so_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
