"""Builds the clodb200 native libraries in-tree.

  product : basicrenderer_b200/libclodb200.so   nvcc, sm_100a only, -lineinfo (needs a CUDA device at run time)
  emu     : tests/emu/libclodb200_emu.so         g++ -DCLODB_EMU: development-only serial emulation of the same kernel
                                                  sources for debugging stage logic without a GPU (never shipped/loaded
                                                  by the product package)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "basicrenderer_b200", "csrc")
SOURCES = ["rt.cu", "remap.cu", "mikk.cu", "clusterize.cu", "bounds.cu", "groups.cu", "partition.cu", "simplify.cu", "output.cu", "dag.cu", "artifacts.cu", "capi.cu", "comm.cu", "cachenames.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
PRODUCT_LIB = os.path.join(ROOT, "basicrenderer_b200", "libclodb200.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libclodb200_emu.so")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "clodb200.h"))
    return hs


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("compile failed: " + cmd[-1])
    return r.stdout + r.stderr


def build_product(force: bool = False, verbose: bool = False, extra_flags=(), lib_path: str | None = None, tag: str = "product") -> str:
    """extra_flags / lib_path / tag build tuning variants next to the product library (tools/tune_*.py)."""
    objdir = os.path.join(ROOT, "build", tag)
    target_lib = lib_path or PRODUCT_LIB
    os.makedirs(objdir, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            # -fmad=false: the reference is IEEE x86-64 code without FMA contraction; keeping mul/add separate is what
            # makes the float stages reproduce its results bit for bit
            jobs.append([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
                         "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"] + list(extra_flags) + ["-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        print("\n".join(outs))
    if jobs or force or not os.path.exists(target_lib):
        _run([NVCC, "-shared", "-o", target_lib] + objs + ["-lcudart", "-ldl"])
    return target_lib


def build_emu(force: bool = False) -> str:
    objdir = os.path.join(ROOT, "build", "emu")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append(["g++", "-x", "c++", "-DCLODB_EMU", "-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-ffp-contract=off", "-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, jobs))
    if jobs or force or not os.path.exists(EMU_LIB):
        _run(["g++", "-shared", "-o", EMU_LIB] + objs + ["-ldl"])
    return EMU_LIB


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("emu", "all"):
        print(build_emu(force="--force" in sys.argv))
    if what in ("product", "all"):
        print(build_product(force="--force" in sys.argv, verbose="-v" in sys.argv))
