"""In-tree build of the sm_100a shared library behind include/seir_b200.h.

    python -m covid19uk_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting ``covid19uk_b200/lib/libseir_b200.so`` is git-ignored
but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libseir_b200.so"
SOURCES = ["seir_api.cu", "ingest.cu", "contract.cu", "contract_i8.cu", "loglik.cu", "delta.cu", "hmc.cu", "hmc_traj.cu", "propose.cu", "sweep.cu", "analytics.cu", "simulate.cu", "host_pack.cpp"]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread"]
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libseir_b200.so")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "seir_b200.h"))
    srcs = _sources()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if not force and not _stale(obj, [path] + headers):
            return obj, ""
        if src.endswith(".cpp"):  # plain host code: g++ directly (function multiversioning attributes are not nvcc-friendly)
            cmd = [os.environ.get("CXX", "g++")] + CXX_FLAGS + ["-c", path, "-o", obj]
        else:
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get("SEIR_NVCC_EXTRA", "").split() + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return obj, res.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    with open(os.path.join(LIBDIR, "ptxas.log"), "a" if not force else "w") as f:
        f.write(log)
    if verbose and log:
        print(log, file=sys.stderr)
    target = lib_path()
    if force or _stale(target, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-pthread", "-o", target] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return target


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
