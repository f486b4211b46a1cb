"""Build libmfar_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python multifield-adaptive-retrieval_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("MFAR_OUT") or os.path.join(HERE, "mfar_b200", "libmfar_b200.so")   # MFAR_OUT: instrumented builds
BUILD_DIR = os.path.join(HERE, "build" + ("_" + os.path.basename(OUT).replace(".", "_") if os.environ.get("MFAR_OUT") else ""))
SOURCES = ["capi.cu", "aux_kernels.cu", "score_simt.cu", "score_tc.cu", "score_qs.cu", "bm25.cu", "train.cu", "topk_rows.cu", "sparse_coo.cu", "union_rescore.cu"]
HEADERS = ["common.cuh", "kernels.h", "tc_ptx.cuh", os.path.join("..", "..", "include", "mfar_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


FLAGS += [f"-D{d}" for d in os.environ.get("MFAR_DEFINES", "").split(",") if d]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    objs = []
    procs = []
    os.makedirs(BUILD_DIR, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(BUILD_DIR, s.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode):
            print(f"--- {s}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        print(r.stdout)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
