"""In-tree build of libfdb200.so (sm_100a only) with nvcc.

    python -m flashdeconv_b200.build [--force] [--verbose]

The shared library is written next to this file so that it travels with the
source tree (it is git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libfdb200.so")
SOURCES = ["common.cu", "sketch.cu", "graph.cu", "bcd.cu", "comm.cu", "peer.cu", "tile.cu", "wide.cu"]
# the production sweep kernel is fully unrolled per row width: one translation unit per Kp, widest (slowest) first
SWEEP_KP = [64, 56, 48, 40, 32, 24, 16, 8]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libfdb200.so cannot be built")
    return cand


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "fdb200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    common = [nvcc, "-O3", "-std=c++17", *ARCH, "-lineinfo", "--extended-lambda", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden"]
    if verbose:
        common += ["-Xptxas", "-v"]

    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(os.path.dirname(HERE), "include", "fdb200.h"), os.path.abspath(__file__)]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def one(job):
        src, defs, tag = job
        obj = os.path.join(OBJ, src.replace(".cu", tag + ".o"))
        # incremental: an object is reused while it is newer than its source and every header
        if (not force and not verbose and os.path.exists(obj)
                and os.path.getmtime(obj) > max(newest_header, os.path.getmtime(os.path.join(CSRC, src)))):
            return obj
        cmd = common + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    # widest rows first: they take longest
    jobs = [("bcd_p_inst.cu", [f"-DFDB_P_KP={kp}", f"-DFDB_P_COMM={c}"], f"_{kp}_{c}") for kp in sorted(SWEEP_KP, reverse=True)
            for c in (0, 1)] + [(s, [], "") for s in SOURCES]
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(one, jobs))
    link = [nvcc, "-shared", *ARCH, "-o", LIB + ".tmp", *objs, "-cudart", "static", "-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
