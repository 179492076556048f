"""Build libfatesplat.so (sm_100a only) in-tree with plain nvcc.

    python -m fateavatar_b200.build [--force] [--verbose]

The library has no torch / pybind dependency: it is a C-ABI shared object loaded with ctypes
(fateavatar_b200/_lib.py).  Kept in-tree (fateavatar_b200/lib/) so it travels with the source snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
# FATESPLAT_BUILD_OUT: developer knob -- build an experiment variant next to the product library (see FATESPLAT_LIB)
OUT = os.environ.get("FATESPLAT_BUILD_OUT") or os.path.join(OUT_DIR, "libfatesplat.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "blend_forward.cu", "backward.cu", "backward_pipe.cu", "knn.cu", "pose.cu", "flame.cu", "stats.cu", "exchange.cu", "optim.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "fatesplat.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # no -use_fast_math: the parity contract needs IEEE div/sqrt and the accurate expf
]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps += [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("FATESPLAT_NVCC_DEFINES", "").split()  # experiment knobs, e.g. -DFS_FWD_GROUP=8
    cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", OUT]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
