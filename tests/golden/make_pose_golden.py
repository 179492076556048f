#!/usr/bin/env python
"""Golden fixture for the mesh part of the pose stage, made by importing the REFERENCE's own
volume_rendering/mesh_compute.py on the CPU (this container only; /root/reference does not travel):

    python tests/golden/make_pose_golden.py        # writes tests/golden/pose_mesh_small.npz

Outputs only (face frames, face scales, un-normalised normals of a seeded posed mesh, and the barycentric splat
positions of volume_rendering/mesh_sampling.py:171-200); the inputs are regenerated
at test time from fateavatar_b200.scenes.pose_inputs(N=10, seed=31).  The quaternion helpers the stage also uses
come from pytorch3d, which the reference does not vendor: nothing to record for them (parity unpinned there).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fateavatar_b200 import scenes  # noqa: E402

REF_MESH = "/root/reference/volume_rendering/mesh_compute.py"
CASE = dict(N=10, seed=31)


def load_ref_mesh_sampling():
    """volume_rendering/mesh_sampling.py imports four pytorch3d symbols at module level that the function used here
    (reweight_verts_by_barycoords, :171-200) never touches; pytorch3d is not installed, so they are stubbed."""
    import types

    for name in ("pytorch3d", "pytorch3d.structures", "pytorch3d.io", "pytorch3d.renderer", "pytorch3d.renderer.mesh",
                 "pytorch3d.ops"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pytorch3d.structures"].Meshes = sys.modules["pytorch3d.io"].load_obj = None
    sys.modules["pytorch3d.renderer.mesh"].rasterize_meshes = sys.modules["pytorch3d.ops"].mesh_face_areas_normals = None
    spec = importlib.util.spec_from_file_location("ref_mesh_sampling", "/root/reference/volume_rendering/mesh_sampling.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    spec = importlib.util.spec_from_file_location("ref_mesh_compute", REF_MESH)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    p = scenes.pose_inputs(**CASE)
    verts, faces = torch.from_numpy(p["verts"])[None], torch.from_numpy(p["faces"])
    orient, scale = ref.compute_face_orientation(verts, faces, return_scale=True)
    normals = ref.compute_face_normals(verts, faces)
    _, canon = ref.compute_face_orientation(torch.from_numpy(p["canon_verts"])[None], faces, return_scale=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose_mesh_small.npz")
    # every 7th face keeps the fixture small
    pos = load_ref_mesh_sampling().reweight_verts_by_barycoords(verts, faces, torch.from_numpy(p["face_index"]),
                                                                torch.from_numpy(p["bary"]))
    np.savez_compressed(path, orient=orient[0, ::7].numpy(), scale=scale[0, ::7].numpy(), normals=normals[0, ::7].numpy(),
                        canon_scale=canon[0, ::7].numpy(), pos=pos[0].numpy())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
