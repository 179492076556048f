"""fateavatar_b200 -- sm_100a 3D-Gaussian-splatting operators behind FateAvatar's rasterizer API.

The package holds only the hot path: the CUDA kernels + C ABI (csrc/, lib/) and the Python
mirrors of the reference's interfaces for that path:

    rasterizer.py, knn.py   diff_gaussian_rasterization / simple_knn operator API          (fs_forward, fs_backward, ...)
    render.py               volume_rendering/render_3dgs.py:render
    flame.py                flame/FLAME.py forward / forward_with_delta_blendshape          (fs_flame_*)
    pose.py                 the per-splat placement block of model/fateavatar.py:225-258     (fs_pose_*)
    densify.py              _add_densification_stats                                         (fs_densify_stats)
    avatar.py               FateAvatar.forward as one call (forward_frame / attach)
    graph.py                whole-frame CUDA-graph capture / replay (CapturedStep)
    optimizer.py, losses.py   optimise loop: fused Adam, in-place densify / prune / reset, OptimiseLoop; fused L1 image loss
    exchange.py, parallel.py  frame-sharded multi-GPU: peer-memory gradient exchange, sampler, densify sync
    io.py                   on-disk formats of the path: the splat PLY (gaussian_model.py:190-269) and the checkpoint
                            layout (trainer.py:396-435 / deserialize.py:7-40) over the capacity-allocated splat store

    import fateavatar_b200; fateavatar_b200.install()
    # from here on `import diff_gaussian_rasterization` / `from simple_knn._C import distCUDA2`
    # resolve to the B200 implementation, so the reference's volume_rendering/*.py run unchanged.
"""
import os
import sys

__version__ = "0.1.0"

_DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install():
    """Make the drop-in operator modules importable under the reference's module names."""
    if _DROPIN not in sys.path:
        sys.path.insert(0, _DROPIN)
    for name in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C"):
        mod = sys.modules.get(name)
        if mod is not None and not (getattr(mod, "__file__", None) or "").startswith(_DROPIN):
            del sys.modules[name]
    import diff_gaussian_rasterization  # noqa: F401
    import simple_knn._C  # noqa: F401
    return _DROPIN
