"""Harness that executes the reference's UNCHANGED `FateAvatar.forward` (model/fateavatar.py:196-298) -- on the CPU
(rasterizer = the C oracle drop-in) or, on the GPU box, from the staged copies under oracle/_ref/pyref with the real
drop-in (or the compiled reference) as `diff_gaussian_rasterization` (tests/test_dropin_reference_gpu.py).

Used by tests/test_avatar_host.py and tests/golden/make_frame_golden.py.  The reference's own Camera, FLAME methods,
mesh_compute / mesh_sampling functions, GaussianModel and render() run as they are; what is substituted:
  * the CUDA rasterizer   -> oracle/cpu_dropin.py (the C oracle behind the operator API),
  * pytorch3d (absent)    -> this repo's restatements of its three quaternion functions (parity unpinned there),
  * plyfile / simple_knn / tools.util (imported, unused on this path) -> stubs,
  * Tensor.cuda()         -> no-op;  constructors that need the licensed FLAME pickle / template assets are bypassed
                             (objects are made with __new__ and given the synthetic model's buffers).
`patch` is pytest's monkeypatch or the `Patch` shim below (setitem / setattr / undo).
"""
import importlib.util
import sys
import types

import numpy as np
import torch

REF = "/root/reference"


class Patch:
    def __init__(self):
        self._undo = []

    def setitem(self, d, k, v):
        self._undo.append((d, k, d.get(k, None), k in d))
        d[k] = v

    def setattr(self, obj, name, v):
        old = getattr(obj, name)
        self._undo.append((obj, name, old, None))
        setattr(obj, name, v)

    def undo(self):
        for d, k, old, had in reversed(self._undo):
            if had is None:
                setattr(d, k, old)
            elif had:
                d[k] = old
            else:
                d.pop(k, None)
        self._undo.clear()


STAGED = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))),
                                    "oracle", "_ref", "pyref")


def load_reference(patch, root=None, rasterizer="cpu", knn=None):
    """Returns (FateAvatar class, FLAME class, mesh_compute module) loaded from the reference tree `root`.
    rasterizer="cpu": the C-oracle drop-in, Tensor.cuda() patched to a no-op (CPU run).  Otherwise a module object that
    provides the `diff_gaussian_rasterization` operator API (fateavatar_b200's drop-in or the compiled reference) and
    nothing about devices is patched (GPU run); `knn` then provides simple_knn._C.distCUDA2."""
    from fateavatar_b200 import avatar
    from oracle import cpu_dropin, oracle as orc, pose_oracle as po

    REF = root or globals()["REF"]

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        patch.setitem(sys.modules, name, m)

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        patch.setitem(sys.modules, name, m)
        spec.loader.exec_module(m)
        return m

    stub("pytorch3d")
    stub("pytorch3d.io", load_obj=None)
    stub("pytorch3d.ops", knn_points=None, mesh_face_areas_normals=None)
    stub("pytorch3d.structures", Meshes=None)
    stub("pytorch3d.renderer")
    stub("pytorch3d.renderer.mesh", rasterize_meshes=None)
    stub("pytorch3d.transforms", quaternion_to_axis_angle=avatar.quaternion_to_axis_angle,
         matrix_to_quaternion=po.matrix_to_quaternion, quaternion_multiply=po.quaternion_multiply)
    stub("plyfile", PlyData=object, PlyElement=object)
    stub("simple_knn")
    stub("simple_knn._C", distCUDA2=knn if knn is not None else
         (lambda p: torch.from_numpy(orc.knn_mean_dist2(p.detach().cpu().numpy()))))
    stub("tools")
    stub("tools.gs_utils")
    stub("tools.util", get_bg_color=lambda c: torch.ones(3) if c == "white" else torch.zeros(3))
    for mod in ("general_utils", "system_utils", "sh_utils", "graphics_utils"):
        load(f"tools.gs_utils.{mod}", f"{REF}/tools/gs_utils/{mod}.py")
    stub("flame")
    load("flame.lbs", f"{REF}/flame/lbs.py")
    FLAME = load("flame.FLAME", f"{REF}/flame/FLAME.py").FLAME
    patch.setitem(sys.modules, "diff_gaussian_rasterization", cpu_dropin if rasterizer == "cpu" else rasterizer)
    stub("volume_rendering")
    for mod in ("camera_3dgs", "gaussian_model", "render_3dgs", "mesh_sampling", "mesh_compute"):
        load(f"volume_rendering.{mod}", f"{REF}/volume_rendering/{mod}.py")
    FateAvatar = load("ref_model_fateavatar", f"{REF}/model/fateavatar.py").FateAvatar
    if rasterizer == "cpu":
        patch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return FateAvatar, FLAME, sys.modules["volume_rendering.mesh_compute"]


PARAMS = ("_scaling", "_rotation", "_offset", "_opacity", "_features_dc", "delta_vertex", "delta_posedirs", "delta_shapedirs")


def build_reference_model(FateAvatar, FLAME, mesh_compute, a, img_res, device="cpu"):
    """A reference FateAvatar instance whose state is the synthetic avatar `a` (scenes.small_avatar / equivalents)."""
    t = lambda x: torch.from_numpy(np.asarray(x)).to(device)
    flame_obj = FLAME.__new__(FLAME)
    torch.nn.Module.__init__(flame_obj)
    flame_obj.dtype, flame_obj.n_shape, flame_obj.n_exp = torch.float32, a["n_shape"], a["n_exp"]
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
        flame_obj.register_buffer(k, t(a[k]))
    flame_obj.register_buffer("parents", t(a["parents"]))
    faces = t(a["faces"])
    _, canon = mesh_compute.compute_face_orientation(t(a["v_template"])[None], faces, return_scale=True)
    ref = FateAvatar.__new__(FateAvatar)
    torch.nn.Module.__init__(ref)
    par = lambda x: torch.nn.Parameter(t(x).clone())
    N = a["face_index"].shape[0]
    ref.flame, ref.device, ref.img_res, ref.shell_len = flame_obj, str(device), tuple(img_res), a["shell_len"]
    ref.cfg_model = types.SimpleNamespace(delta_blendshape=True, delta_vertex=True, resize_scale=True)
    ref.bg_color = torch.ones(3, device=device)
    ref.faces, ref.face_index, ref.bary_coords, ref.face_scaling_canonical = faces, t(a["face_index"]), t(a["bary"]), canon
    ref.delta_shapedirs, ref.delta_posedirs, ref.delta_vertex = par(a["delta_shapedirs"]), par(a["delta_posedirs"]), par(a["delta_vertex"])
    ref._features_dc, ref._features_rest = par(a["features_dc"]), torch.zeros(N, 0, 3, device=device)
    ref._scaling, ref._rotation, ref._offset, ref._opacity = par(a["scaling_raw"]), par(a["rotation_raw"]), par(a["offset_raw"]), par(a["opacity_raw"])
    return ref


def frame_input(a, fovx=0.35, fovy=0.3, T=(0.02, -0.01, 1.25), device="cpu"):
    t = lambda x: torch.from_numpy(np.asarray(x)).to(device)
    cam_pose = np.eye(4, dtype=np.float32)
    cam_pose[:3, :3], cam_pose[:3, 3] = np.diag([1.0, -1.0, -1.0]), T
    return dict(cam_pose=t(cam_pose)[None], fovx=torch.tensor([fovx]), fovy=torch.tensor([fovy]),
                flame_pose=t(a["pose"])[None], expression=t(a["betas"][a["n_shape"]:])[None])
