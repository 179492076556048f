#!/usr/bin/env python
"""Stage the reference's UNCHANGED Python callers for the GPU box (TEST INFRASTRUCTURE, not product).

/root/reference does not exist on the GPU box, so the handful of upstream files that sit directly above the operator
boundary are copied -- byte for byte, nothing edited -- into the git-ignored oracle/_ref/pyref/ (it travels with the
snapshot exactly like the compiled oracle/_ref/*.so; nothing under oracle/_ref is ever committed):

    volume_rendering/{render_3dgs,gaussian_model,camera_3dgs,mesh_compute,mesh_sampling}.py
    tools/gs_utils/{general_utils,system_utils,sh_utils,graphics_utils}.py
    flame/{FLAME,lbs}.py        model/fateavatar.py        weights/head_template_mouth_close.obj
    submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py -> oracle/_ref/ref_dgr/__init__.py
        (next to the compiled reference _C.so, so `import ref_dgr` is the reference's own operator API on its own kernels)

tests/test_dropin_reference_gpu.py imports them verbatim on the B200 with fateavatar_b200.install() providing
`diff_gaussian_rasterization` / `simple_knn`, and compares against the same files driving the compiled reference.
"""
import hashlib
import os
import shutil

REF = os.environ.get("FATE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "pyref")

FILES = [
    "volume_rendering/render_3dgs.py", "volume_rendering/gaussian_model.py", "volume_rendering/camera_3dgs.py",
    "volume_rendering/mesh_compute.py", "volume_rendering/mesh_sampling.py",
    "tools/gs_utils/general_utils.py", "tools/gs_utils/system_utils.py", "tools/gs_utils/sh_utils.py",
    "tools/gs_utils/graphics_utils.py",
    "flame/FLAME.py", "flame/lbs.py", "model/fateavatar.py",
    "weights/head_template_mouth_close.obj",
]
DGR_INIT = "submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py"


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found")
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{_sha(dst)}  {rel}")
    dst = os.path.join(HERE, "_ref", "ref_dgr", "__init__.py")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(os.path.join(REF, DGR_INIT), dst)
    manifest.append(f"{_sha(dst)}  {DGR_INIT}")
    with open(os.path.join(OUT, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    print(f"[stage_ref_py] {len(manifest)} files -> {OUT}")
    return OUT


def verify():
    """True when every staged file still equals its upstream source (checked where /root/reference is mounted)."""
    for line in open(os.path.join(OUT, "MANIFEST.sha256")):
        sha, rel = line.split()
        src = os.path.join(REF, rel)
        if os.path.exists(src) and _sha(src) != sha:
            return False
    return True


if __name__ == "__main__":
    stage()
