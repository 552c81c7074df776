#!/usr/bin/env python3
"""Compile csrc/*.cu into lib/libnbody_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python nbody-gnn-hpc_b200/build.py [--force] [--verbose]

The library is built in-tree so that it travels with the repository snapshot to the GPU box; it
is git-ignored.  ptxas resource usage of every kernel is written to lib/ptxas.log.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIBDIR = HERE / "lib"
LIB = LIBDIR / "libnbody_b200.so"
HEADER = HERE.parent / "include" / "nbody_b200.h"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def stale() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [HEADER, Path(__file__)])
    return LIB.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: Path = None) -> Path:
    """extra_flags / out: experiment builds (e.g. -DNB_F64_PAIR_FIRST_ORDER) into another file; the default build is
    what everything loads (NBODY_B200_LIB selects another library at run time)."""
    if out is None and not extra_flags and not force and not stale():
        return LIB
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    lib_out = Path(out) if out is not None else LIB
    objdir = LIBDIR / ("obj" if out is None else "obj_" + lib_out.stem)
    objdir.mkdir(exist_ok=True)
    log = []
    objs = []
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc (a wrapper); let nvcc use the distro g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if Path("/usr/bin/g++").exists() else []
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, *ccbin, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
    for src, obj, cmd, p in procs:
        text, _ = p.communicate()
        log.append("$ " + " ".join(cmd) + "\n" + text)
        if p.returncode != 0:
            sys.stderr.write(text)
            raise RuntimeError(f"nvcc failed on {src.name}")
        objs.append(str(obj))
    cmd = [nvcc, "-shared", *ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib_out), *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    (LIBDIR / ("ptxas.log" if out is None else f"ptxas_{lib_out.stem}.log")).write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib_out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--out", default=None, help="experiment build: write the library here")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="experiment build: extra -D macro")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, extra_flags=[f"-D{d}" for d in a.defines], out=a.out))
