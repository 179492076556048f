"""The reference-side binding shown in INTEGRATION.md section 2 must at least compile against torch's headers and
include/fatesplat.h: the block is extracted from the document and type-checked with g++ -fsyntax-only (no GPU)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_integration_md_pybind_shim_compiles(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from torch.utils import cpp_extension as ce

    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```cpp\n(.*?)```", md, flags=re.S)
    shim = [b for b in blocks if "RasterizeGaussiansCUDA" in b]
    assert len(shim) == 1, "INTEGRATION.md must hold exactly one RasterizeGaussiansCUDA shim"
    src = tmp_path / "shim.cpp"
    src.write_text('#include <tuple>\n#include <torch/extension.h>\n#include <ATen/cuda/CUDAContext.h>\n'
                   '#include <cuda_runtime.h>\n#include "fatesplat.h"\n' + shim[0])
    inc = ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True)
    import sysconfig

    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-I", os.path.join(ROOT, "include"),
           "-I", sysconfig.get_paths()["include"]] + [a for p in inc for a in ("-isystem", p)] + [str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
