"""Build libsqe_b200.so in-tree with nvcc for sm_100a (no network, no torch needed).

    python semantic-query-engine_b200/build.py [--force] [--verbose]

The library links cudart statically and resolves the one driver entry point it
needs (cuTensorMapEncodeTiled) at run time, so it loads on a machine without
libcuda (e.g. the CPU build container) and exports every symbol of
include/sqe_b200.h there.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsqe_b200.so")

SOURCES = ["api.cu", "normalize_cast.cu", "topk_gemv.cu", "topk_prefilter.cu", "topk_batched.cu", "topk_batched_i8.cu", "exchange.cu",
           "encoder_gemm.cu", "encoder_gemm_small.cu", "encoder_attn.cu", "encoder_rows.cu"]
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]
LINK_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(nvcc: str, src: str, obj: str, extra, verbose: bool):
    cmd = [nvcc] + COMPILE_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return proc.returncode, proc.stdout, cmd


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per source, compiled in parallel and only when the source (or any header) is
    newer than its object; then one link.  Objects live in lib/obj/ (git-ignored)."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = find_nvcc()
    extra = os.environ.get("SQE_NVCC_EXTRA", "").split()      # extra nvcc flags for experiments
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + \
              [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)] + [os.path.abspath(__file__)]
    t_hdr = max(os.path.getmtime(h) for h in headers)
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s[:-3] + ".o")
        if force or extra or verbose or not os.path.isfile(obj) or \
                os.path.getmtime(obj) < max(t_hdr, os.path.getmtime(src)):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1], extra, verbose), jobs))
    for rc, out, cmd in results:
        if verbose or rc != 0:
            sys.stderr.write(out)
        if rc != 0:
            raise RuntimeError("nvcc failed (exit %d): %s" % (rc, " ".join(cmd)))
    cmd = [nvcc] + LINK_FLAGS + ["-o", LIB_PATH] + [os.path.join(obj_dir, s[:-3] + ".o") for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed (exit %d): %s" % (proc.returncode, " ".join(cmd)))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
