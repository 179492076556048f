"""ctypes binding of libfatesplat.so -- the only native code the product loads.

There is NO fallback: if the shared object is missing or cannot be loaded, importing a drop-in operator
raises.  The oracle under oracle/ is never consulted from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FATESPLAT_LIB: developer knob to load another build of the same library (kernel experiments); never a fallback
LIB_PATH = os.environ.get("FATESPLAT_LIB") or os.path.join(_HERE, "lib", "libfatesplat.so")

FS_OK = 0


class FsFrameInfo(C.Structure):
    _fields_ = [
        ("num_rendered", C.c_uint32),
        ("overflow", C.c_uint32),
        ("num_visible", C.c_uint32),
        ("max_tile_instances", C.c_uint32),
        ("reserved", C.c_uint32 * 4),
    ]


_LAYOUT_FIELDS = [
    "total_bytes", "info", "depths", "cov3D", "splat", "clamped", "rect", "tiles_touched", "tile_count",
    "tile_cursor", "ranges", "big_tiles", "work_order", "tile_meta", "seg_base", "seg_info", "ckpt", "final_C", "inst_keys", "inst_keys_alt", "point_list", "inst_splat", "final_T",
    "n_contrib", "bwd_counter", "grad_acc", "instance_capacity", "pair_mask",
]


class FsWorkspaceLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in _LAYOUT_FIELDS]


# every symbol include/fatesplat.h declares
EXPORTS = [
    "fs_workspace_bytes", "fs_get_workspace_layout", "fs_forward", "fs_backward", "fs_mark_visible",
    "fs_knn_workspace_bytes", "fs_knn_mean_dist2", "fs_last_launch_count", "fs_last_error", "fs_version",
    "fs_profile_enable", "fs_profile_read", "fs_pose_forward", "fs_pose_backward", "fs_set_tile_hint",
    "fs_flame_workspace_bytes", "fs_flame_forward", "fs_flame_backward", "fs_flame_backward_coeffs", "fs_densify_stats", "fs_flame_expand_grads", "fs_p2p_allreduce", "fs_p2p_reduce_scatter_bcast",
    "fs_p2p_exchange", "fs_p2p_wait", "fs_p2p_exchange_flag_floats", "fs_densify_stats_inc", "fs_p2p_exchange_timing", "fs_set_early_notify", "fs_frame_camera", "fs_l1_loss", "fs_l1_loss_workspace_bytes",
    "fs_adam_step", "fs_splat_append", "fs_splat_prune_workspace_bytes", "fs_splat_prune", "fs_opacity_reset",
]

STAGES = ["preprocess", "tile_scan", "scatter", "tile_sort", "big_tile_sort", "blend_forward", "blend_backward",
          "preprocess_backward", "knn", "pose_forward", "pose_backward", "flame_forward", "flame_backward", "exchange"]

_lib = None


class FateSplatError(RuntimeError):
    pass


def load():
    """Load (once) and type the library.  Raises FateSplatError when it is absent: no CPU path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FateSplatError(
            f"{LIB_PATH} not found: build it with `python -m fateavatar_b200.build` (needs nvcc, sm_100a). "
            "fateavatar_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, f, i, sz = C.c_void_p, C.c_float, C.c_int, C.c_size_t
    lib.fs_workspace_bytes.restype = sz
    lib.fs_workspace_bytes.argtypes = [i, i, i, sz]
    lib.fs_get_workspace_layout.restype = i
    lib.fs_get_workspace_layout.argtypes = [i, i, i, sz, C.POINTER(FsWorkspaceLayout)]
    lib.fs_forward.restype = i
    lib.fs_forward.argtypes = [i, i, i, vp, i, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, i, vp, vp, vp, sz,
                               sz, vp, vp]
    lib.fs_backward.restype = i
    lib.fs_backward.argtypes = [i, i, i, vp, i, i, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, vp, vp, sz, sz, vp,
                                vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fs_mark_visible.restype = i
    lib.fs_mark_visible.argtypes = [i, vp, vp, vp, vp, vp]
    lib.fs_knn_workspace_bytes.restype = sz
    lib.fs_knn_workspace_bytes.argtypes = [i]
    lib.fs_knn_mean_dist2.restype = i
    lib.fs_knn_mean_dist2.argtypes = [i, vp, vp, vp, sz, vp]
    lib.fs_pose_forward.restype = i
    lib.fs_pose_forward.argtypes = [i, i, i] + [vp] * 9 + [f, i] + [vp] * 4 + [vp]
    lib.fs_pose_backward.restype = i
    lib.fs_pose_backward.argtypes = [i, i, i] + [vp] * 9 + [f, i] + [vp] * 4 + [vp] * 5 + [vp]
    lib.fs_flame_workspace_bytes.restype = sz
    lib.fs_flame_workspace_bytes.argtypes = [i]
    lib.fs_flame_forward.restype = i
    lib.fs_flame_forward.argtypes = [i, i, i, i, C.POINTER(C.c_int)] + [vp] * 10 + [vp] * 5 + [vp, sz, vp]
    lib.fs_flame_backward.restype = i
    lib.fs_flame_backward.argtypes = [i, i, i, i, C.POINTER(C.c_int)] + [vp] * 4 + [vp, sz] + [vp] * 6 + [vp]
    lib.fs_flame_backward_coeffs.restype = i
    lib.fs_flame_backward_coeffs.argtypes = [i, i, i, i, C.POINTER(C.c_int)] + [vp] * 6 + [vp, sz, vp, vp, vp]
    lib.fs_flame_expand_grads.restype = i
    lib.fs_flame_expand_grads.argtypes = [i, i, i, i, i, vp, sz, f, vp, vp, vp, vp]
    lib.fs_p2p_reduce_scatter_bcast.restype = i
    lib.fs_p2p_reduce_scatter_bcast.argtypes = [i, i, vp, vp, sz, vp]
    lib.fs_p2p_allreduce.restype = i
    lib.fs_p2p_allreduce.argtypes = [i, vp, vp, sz, sz, vp, vp]
    lib.fs_p2p_exchange_flag_floats.restype = sz
    lib.fs_p2p_exchange_flag_floats.argtypes = []
    lib.fs_p2p_exchange.restype = i
    lib.fs_p2p_exchange.argtypes = [i, i, i, vp, vp, vp, sz, sz, sz, sz, sz, sz, sz, vp, i, i, i, i, f, vp, vp, vp, vp]
    lib.fs_p2p_exchange_timing.restype = i
    lib.fs_p2p_exchange_timing.argtypes = [vp, sz, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), i]
    lib.fs_p2p_wait.restype = i
    lib.fs_p2p_wait.argtypes = [i, vp, sz, vp]
    lib.fs_densify_stats_inc.restype = i
    lib.fs_densify_stats_inc.argtypes = [i, vp, vp, vp, vp, vp]
    lib.fs_densify_stats.restype = i
    lib.fs_densify_stats.argtypes = [i, vp, vp, vp, vp, vp]
    lib.fs_l1_loss_workspace_bytes.restype = sz
    lib.fs_l1_loss_workspace_bytes.argtypes = []
    lib.fs_l1_loss.restype = i
    lib.fs_l1_loss.argtypes = [sz, vp, vp, vp, vp, vp, vp]
    lib.fs_frame_camera.restype = i
    lib.fs_frame_camera.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.fs_set_early_notify.restype = None
    lib.fs_set_early_notify.argtypes = [i]
    lib.fs_set_tile_hint.restype = None
    lib.fs_set_tile_hint.argtypes = [C.c_uint32]
    lib.fs_profile_enable.restype = None
    lib.fs_profile_enable.argtypes = [i]
    lib.fs_profile_read.restype = i
    lib.fs_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), i]
    lib.fs_last_launch_count.restype = i
    lib.fs_last_error.restype = C.c_char_p
    lib.fs_version.restype = C.c_char_p
    _lib = lib
    return lib


def stream_ptr(dev):
    """Raw cudaStream_t of torch's current stream on `dev` (what every launch in this package goes to).  The
    private accessor is ~30x cheaper than torch.cuda.current_stream(dev).cuda_stream, which matters when a frame is
    a dozen ctypes calls."""
    import torch

    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class on_device:
    """`with on_device(dev):` -- make `dev` the current CUDA device for the launches inside; free when it already is."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index if dev.index is not None else -1
        self.prev = -1

    def __enter__(self):
        import torch

        if self.idx >= 0:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            import torch

            torch.cuda.set_device(self.prev)
            self.prev = -1
        return False


def check(rc, what):
    if rc != FS_OK:
        raise FateSplatError(f"{what} failed ({rc}): {load().fs_last_error().decode()}")


def workspace_layout(P, W, H, capacity):
    L = FsWorkspaceLayout()
    check(load().fs_get_workspace_layout(int(P), int(W), int(H), int(capacity), C.byref(L)), "fs_get_workspace_layout")
    return L


def profile_read():
    """{stage: (total_ms, launches)} accumulated since the last read (see fs_profile_enable)."""
    n = len(STAGES)
    ms = (C.c_float * n)()
    cnt = (C.c_int * n)()
    check(load().fs_profile_read(ms, cnt, n), "fs_profile_read")
    return {STAGES[k]: (float(ms[k]), int(cnt[k])) for k in range(n)}
