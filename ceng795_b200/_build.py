"""Build recipes for the native pieces (run by __graft_entry__.build()).

* ``libb200rt.so``  — the product: host builder (g++) + sm_100a CUDA core (nvcc).
* ``oracle/liboracle.so`` — the CPU restatement used ONLY by tests / smoke / the
  bench's cpu_baseline leg.
* ``oracle/_ref/*`` — the untouched reference trees, when /root/reference exists.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ceng795_b200", "csrc")
LIB = os.path.join(ROOT, "ceng795_b200", "libb200rt.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # accept/reject geometry must not contract mul+add into FMA (SURVEY.md §7 hard part 2/3);
    # kernels opt back in to FMA explicitly (__fmaf_rn) where the result is not compared.
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {cmd[0]}")


def build_lib(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    hdrs = [os.path.join(ROOT, "include", "b200rt.h")] + [
        os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".inc"))]
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cpp = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cpp")]
    if not force and not _newer(LIB, hdrs + cu + cpp):
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libb200rt.so")
    objs = []
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    for src in cpp:
        o = os.path.join(ROOT, "build", os.path.basename(src) + ".o")
        if force or _newer(o, [src] + hdrs):
            _run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-c", src, "-o", o])
        objs.append(o)
    for src in cu:
        o = os.path.join(ROOT, "build", os.path.basename(src) + ".o")
        if force or _newer(o, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose:
                sys.stderr.write(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                raise RuntimeError("nvcc failed")
        objs.append(o)
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ROOT, "oracle", "oracle.c")
    deps = [src, os.path.join(ROOT, "include", "b200rt.h")]
    if force or _newer(ORACLE_LIB, deps):
        _run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
              src, "-o", ORACLE_LIB, "-lm"])
    return ORACLE_LIB


def build_ref() -> None:
    """Compile the untouched reference trees into oracle/_ref/ (only where
    /root/reference exists — the GPU box uses the prebuilt files)."""
    ref = os.environ.get("CENG795_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        return
    _run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle(force="--force" in sys.argv)
    print(LIB, ORACLE_LIB)
