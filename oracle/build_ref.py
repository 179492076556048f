#!/usr/bin/env python
"""Build recipe for the GPU reference oracle (TEST INFRASTRUCTURE, not product).

Compiles the reference's own rasterizer / simple-knn sources *where they lie*
under /root/reference (nothing is copied into this repo) into

    oracle/_ref/ref_dgr/_C.so      (diff-gaussian-rasterization, pybind module `_C`)
    oracle/_ref/ref_knn/_C.so      (simple-knn, pybind module `_C`)

with plain nvcc command lines (the reference's setup.py / CMake are not run).
`oracle/_ref/` is git-ignored but travels to the GPU box with `gpurun`, where
tests/ and bench.py use it as the *checker* / baseline only.

The only deviation from a stock build: `-include cstdint` because
cuda_rasterizer/rasterizer_impl.h:24,40 use std::uintptr_t / uint32_t without
including the header (gcc 13 no longer pulls it in transitively).
"""
import os
import subprocess
import sys
import sysconfig

REF = os.environ.get("FATE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce

    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = ce.library_paths()[0]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = inc + [
        f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
        "-DTORCH_EXTENSION_NAME=_C",
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-std=c++17",
        "-O3",
    ]
    link = [
        f"-L{libdir}",
        "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
        f"-Xlinker", f"-rpath={libdir}",
    ]
    return common, link


def _nvcc(srcs, out, extra):
    common, link = _torch_flags()
    cmd = (
        ["nvcc", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100,code=sm_100",
         "-lineinfo", "-w", "--expt-relaxed-constexpr"]
        + common + extra + srcs + ["-o", out] + link
    )
    print(" ".join(cmd))
    subprocess.check_call(cmd)


def build(force=False):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: using prebuilt oracle/_ref if any")
        return False
    dgr = os.path.join(REF, "submodules/diff-gaussian-rasterization")
    knn = os.path.join(REF, "submodules/simple-knn")
    ok = True
    tgt = os.path.join(OUT, "ref_dgr", "_C.so")
    if force or not os.path.exists(tgt):
        os.makedirs(os.path.dirname(tgt), exist_ok=True)
        _nvcc(
            [os.path.join(dgr, f) for f in (
                "cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu",
                "cuda_rasterizer/backward.cu", "rasterize_points.cu", "ext.cpp")],
            tgt,
            [f"-I{os.path.join(dgr, 'third_party/glm')}", f"-I{dgr}", "-include", "cstdint"],
        )
    tgt = os.path.join(OUT, "ref_knn", "_C.so")
    if force or not os.path.exists(tgt):
        os.makedirs(os.path.dirname(tgt), exist_ok=True)
        _nvcc(
            [os.path.join(knn, f) for f in ("spatial.cu", "simple_knn.cu", "ext.cpp")],
            tgt, [f"-I{knn}"],
        )
    return ok


if __name__ == "__main__":
    build(force="--force" in sys.argv)
