"""Build the CUDA library (sm_100a only) in-tree: ``python -m mm_b200.build``.

nvcc cross-compiles without a GPU; the resulting ``libmm_b200.so`` sits next to this file so
that it travels with a repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(HERE, "..", "csrc"))
INCLUDE = os.path.normpath(os.path.join(HERE, "..", "..", "include"))
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libmm_b200.so")

SOURCES = ["stages.cu", "capi.cu", "nccl_shim.cu", "bandcomp.cu", "deesser.cu", "followers.cu", "reverb.cu", "spectral.cu", "denoise.cu", "bigfft.cu", "export.cu", "analyzers.cu", "context.cu", "design.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # every rounding is the one written (numpy float32 steps are reproduced)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, defs=(), tag: str = "") -> str:
    """defs / tag: an experiment build (extra -D flags) into libmm_b200_<tag>.so with its own object directory; the loader picks
    it up through MM_B200_LIB (A/B timing of kernel variants in one GPU call).  The product build has neither."""
    global BUILD, LIB
    if tag:
        BUILD_, LIB_ = os.path.join(CSRC, "build_" + tag), os.path.join(HERE, f"libmm_b200_{tag}.so")
        saved = (BUILD, LIB, list(NVCC_FLAGS))
        BUILD, LIB = BUILD_, LIB_
        NVCC_FLAGS.extend(defs)
        try:
            return build(force, verbose)
        finally:
            BUILD, LIB = saved[0], saved[1]
            NVCC_FLAGS[:] = saved[2]
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(INCLUDE, "mm_b200.h"))
    objs, jobs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        stamp = obj + ".sha"
        dig = _digest([sp] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", sp, "-o", obj]
        jobs.append((cmd, stamp, dig, src))

    def run(job):
        cmd, stamp, dig, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dig)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defs=[a for a in sys.argv if a.startswith("-D")], tag=tag))
