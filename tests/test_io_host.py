"""On-disk formats of the path (second half): the splat PLY of gaussian_model.py:190-269 and the checkpoint
layout of trainer.py:396-435 / deserialize.py:7-40.  CPU only."""
import numpy as np
import torch

from fateavatar_b200 import io as fio


def _splats(P=37, K=15, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(xyz=r(P, 3), features_dc=r(P, 1, 3), features_rest=r(P, K, 3), opacity=r(P, 1), scaling=r(P, 3),
                rotation=r(P, 4))


def test_ply_layout_is_the_reference_one_and_round_trips(tmp_path):
    s = _splats()
    path = str(tmp_path / "point_cloud.ply")
    fio.write_ply(path, **s)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode("ascii").split("\n")
    # the header plyfile writes for PlyElement.describe(elements, 'vertex') with all-'f4' columns (gaussian_model.py:215-221)
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    names = [l.split()[2] for l in lines[3:] if l]
    assert all(l.startswith("property float ") for l in lines[3:] if l)
    assert names == (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] +
                     [f"f_rest_{i}" for i in range(45)] + ["opacity"] + [f"scale_{i}" for i in range(3)] +
                     [f"rot_{i}" for i in range(4)])
    rows = np.frombuffer(body, dtype="<f4").reshape(37, len(names))
    assert len(body) == 37 * len(names) * 4
    assert np.array_equal(rows[:, 0:3], s["xyz"].numpy()) and not rows[:, 3:6].any()
    # channel-major feature columns: f_dc_c = features_dc[:, 0, c]; f_rest_{c*K+k} = features_rest[:, k, c]
    assert np.array_equal(rows[:, 6:9], s["features_dc"][:, 0, :].numpy())
    assert np.array_equal(rows[:, 9 + 1 * 15 + 4], s["features_rest"][:, 4, 1].numpy())
    back = fio.read_ply(path, max_sh_degree=3)
    for k, v in s.items():
        assert back[k].shape == v.shape and torch.equal(back[k], v), k


def test_ply_sh0_and_ascii(tmp_path):
    s = _splats(P=5, K=0)
    path = str(tmp_path / "a.ply")
    fio.write_ply(path, **s)
    back = fio.read_ply(path, max_sh_degree=0)
    assert back["features_rest"].shape == (5, 0, 3) and torch.equal(back["rotation"], s["rotation"])
    names = fio.ply_attributes(3, 0)
    cols = torch.cat((s["xyz"], torch.zeros(5, 3), s["features_dc"][:, 0], s["opacity"], s["scaling"], s["rotation"]), 1)
    with open(tmp_path / "b.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 5\n")
        f.write("".join(f"property float {n}\n" for n in names) + "end_header\n")
        for r in cols.tolist():
            f.write(" ".join(repr(float(np.float32(x))) for x in r) + "\n")
    b2 = fio.read_ply(str(tmp_path / "b.ply"))
    assert torch.allclose(b2["scaling"], s["scaling"]) and torch.allclose(b2["features_dc"], s["features_dc"])


class _Model(torch.nn.Module):  # the attributes of model/fateavatar.py that a checkpoint carries
    def __init__(self, P):
        super().__init__()
        g = torch.Generator().manual_seed(P)
        for n, shape in (("_offset", (P, 1)), ("_features_dc", (P, 1, 3)), ("_features_rest", (P, 0, 3)),
                         ("_scaling", (P, 3)), ("_rotation", (P, 4)), ("_opacity", (P, 1))):
            setattr(self, n, torch.nn.Parameter(torch.randn(*shape, generator=g)))
        self.register_buffer("face_index", torch.randint(0, 100, (P,), generator=g))
        self.register_buffer("bary_coords", torch.rand(P, 3, generator=g))
        self.delta_vertex = torch.nn.Parameter(torch.randn(11, 3, generator=g))
        self.num_points = P


def test_checkpoint_layout_round_trip_with_a_changed_splat_count(tmp_path):
    big = _Model(50)
    big._scaling = torch.nn.Parameter(torch.randn(200, 3)[:50])  # a view of a larger allocation, as SplatStore binds them
    state = {"epoch": 3, "global_step": 1234, "model": fio.model_state(big)}  # trainer.py:404-410
    assert state["model"]["_scaling"].untyped_storage().nbytes() == 50 * 3 * 4  # compact, not the 200-row allocation
    path = str(tmp_path / "ckpt_ep0003.pth")
    torch.save(state, path)
    small = _Model(20)  # a freshly constructed model has the initial splat count
    ck = torch.load(path)
    assert set(fio.GAUSSIAN_ATTRIBUTES) <= set(ck["model"])
    missing, unexpected = fio.restore_splats(small, ck["model"])
    assert missing == [] and unexpected == []
    assert small.num_points == 50 and small.xyz_gradient_accum.shape == (50, 1) and small.max_radii2D.shape == (50,)
    for k in fio.GAUSSIAN_ATTRIBUTES:
        assert torch.equal(getattr(small, k).detach(), getattr(big, k).detach()), k
    assert isinstance(small._offset, torch.nn.Parameter) and small._offset.requires_grad
    assert not isinstance(small.face_index, torch.nn.Parameter)
    assert torch.equal(small.delta_vertex.detach(), big.delta_vertex.detach())


def test_restore_splats_equals_the_reference_deserializer(tmp_path):
    """The reference's unchanged train/deserialize.py:deserialize_checkpoints_fateavatar on the same checkpoint."""
    import importlib.util
    import os
    import types

    import pytest

    src = "/root/reference/train/deserialize.py"
    if not os.path.exists(src):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location("ref_deserialize", src)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    big = _Model(64)
    ck = {"model": fio.model_state(big)}
    mine, theirs = _Model(20), _Model(20)
    fio.restore_splats(mine, ck["model"])
    trainer = types.SimpleNamespace(model=theirs, device=torch.device("cpu"), log=lambda *a, **k: None)
    ref.deserialize_checkpoints_fateavatar(trainer, {"model": dict(ck["model"])})
    for k in fio.GAUSSIAN_ATTRIBUTES + ("max_radii2D", "xyz_gradient_accum", "denom", "sample_flag", "delta_vertex"):
        a, b = getattr(mine, k), getattr(theirs, k)
        assert type(a) is type(b) and a.shape == b.shape and torch.equal(a.detach(), b.detach()), k
    assert mine.num_points == theirs.num_points == 64
