"""On-disk formats either side of the path (second half): the splat PLY and the checkpoint layout.

    write_ply / read_ply           volume_rendering/gaussian_model.py:190-269 (save_ply / load_ply): the vanilla-3DGS PLY
                                   -- binary little-endian, one float property per column in the order x y z nx ny nz
                                   f_dc_* f_rest_* opacity scale_* rot_* -- that downstream viewers / bakers read.  The
                                   reference writes it with the `plyfile` package; this is the same byte layout without
                                   the dependency.
    model_state / restore_splats   train/trainer.py:396-435 (save_checkpoint: state['model'] = model.state_dict()) and
                                   train/deserialize.py:7-40 (deserialize_checkpoints_fateavatar): the splat attributes
                                   have a variable number of rows, so they are taken out of the state dict and re-created;
                                   the densification statistics restart from zero.

With optimizer.SplatStore the model's per-splat tensors are [:P] views of capacity-sized arrays; torch.save() of such a
view would write the whole capacity.  model_state() therefore clones every entry into compact storage, and
restore_splats(..., store=) refills the store's rows in place (no reallocation) instead of re-creating tensors.
"""
import numpy as np
import torch

GAUSSIAN_ATTRIBUTES = ("_offset", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "face_index",
                       "bary_coords")  # train/deserialize.py:10-12


# ---- PLY ----------------------------------------------------------------------------------------------------------------
def ply_attributes(n_dc, n_rest, n_scale=3, n_rot=4):
    """Column names in file order (gaussian_model.py:190-202)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    names += [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)]
    return names


def write_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation):
    """xyz [P,3], features_dc [P,1,3], features_rest [P,K,3] (K may be 0), opacity [P,1], scaling [P,3], rotation [P,4]:
    the *raw* (pre-activation) tensors, exactly what GaussianModel.save_ply stores; normals are written as zeros."""
    t = lambda a: a.detach().cpu().float() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float32)
    xyz, opacity, scaling, rotation = t(xyz), t(opacity).reshape(-1, 1), t(scaling), t(rotation)
    P = xyz.shape[0]
    f_dc = t(features_dc).reshape(P, -1, 3).transpose(1, 2).flatten(start_dim=1)      # channel-major, like upstream
    f_rest = t(features_rest).reshape(P, -1, 3).transpose(1, 2).flatten(start_dim=1)
    cols = torch.cat((xyz, torch.zeros_like(xyz), f_dc, f_rest, opacity, scaling, rotation), dim=1).contiguous()
    names = ply_attributes(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    assert cols.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {P}\n"
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(cols.numpy().astype("<f4").tobytes())


def read_ply(path, max_sh_degree=None):
    """Inverse of write_ply / reader for files written by the reference: returns the dict of raw tensors in the shapes
    GaussianModel.load_ply builds (features_dc [P,1,3], features_rest [P,K,3]).  Accepts binary_little_endian and ascii
    vertex elements made of scalar properties (float / double / integer types are converted to float32)."""
    np_types = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
                "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
                "double": "f8", "float64": "f8"}
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            w = line.decode("ascii").split()
            if not w or w[0] == "comment":
                continue
            if w[0] == "format":
                fmt = w[1]
            elif w[0] == "element":
                in_vertex = w[1] == "vertex"
                if in_vertex:
                    count = int(w[2])
                elif count == 0:
                    raise ValueError(f"{path}: the vertex element must come first")
            elif w[0] == "property" and in_vertex:
                if w[1] == "list":
                    raise ValueError(f"{path}: list properties are not part of the splat format")
                props.append((w[2], np_types[w[1]]))
            elif w[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            rec = np.dtype([(n, "<" + ty) for n, ty in props])
            data = np.frombuffer(f.read(rec.itemsize * count), dtype=rec, count=count)
            col = {n: data[n].astype(np.float32) for n, _ in props}
        elif fmt == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2)
            col = {n: rows[:, i].astype(np.float32) for i, (n, _) in enumerate(props)}
        else:
            raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
    by_index = lambda prefix: sorted((n for n in col if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    stack = lambda names: torch.from_numpy(np.stack([col[n] for n in names], axis=1)) if names else torch.zeros(count, 0)
    rest = by_index("f_rest_")
    if max_sh_degree is not None and len(rest) != 3 * (max_sh_degree + 1) ** 2 - 3:  # gaussian_model.py:242
        raise ValueError(f"{path}: {len(rest)} f_rest columns do not match max_sh_degree {max_sh_degree}")
    P = count
    return {"xyz": stack(["x", "y", "z"]),
            "features_dc": stack(by_index("f_dc_")).reshape(P, 3, -1).transpose(1, 2).contiguous(),
            "features_rest": stack(rest).reshape(P, 3, -1).transpose(1, 2).contiguous(),
            "opacity": stack(["opacity"]), "scaling": stack(by_index("scale_")), "rotation": stack(by_index("rot_"))}


# ---- checkpoint ---------------------------------------------------------------------------------------------------------
def model_state(model):
    """What Trainer.save_checkpoint stores under 'model' (trainer.py:410), every entry cloned into compact storage."""
    return {k: v.detach().clone().contiguous() for k, v in model.state_dict().items()}


def restore_splats(model, model_state_dict, store=None, device=None):
    """deserialize_checkpoints_fateavatar (train/deserialize.py:7-40) for `model`: the eight variable-length splat
    attributes are taken out of the dict, everything else goes through load_state_dict(strict=False), the splat
    attributes are re-created (Parameters for the six trained ones), num_points is overwritten and the densification
    statistics restart.  With `store` (optimizer.SplatStore) the rows are written into the store instead and the model
    is re-bound to its views.  Returns (missing_keys, unexpected_keys) like upstream logs them."""
    sd = dict(model_state_dict)
    g = {k: sd.pop(k) for k in GAUSSIAN_ATTRIBUTES}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [k for k in missing if k not in GAUSSIAN_ATTRIBUTES]
    dev = torch.device(device) if device is not None else (store.device if store is not None else g["_offset"].device)
    P = int(g["_offset"].shape[0])
    if store is not None:  # rows go into the store; store.model (normally `model` itself) is re-bound to its views
        store.load_rows({k: v.to(dev) for k, v in g.items()})
        if hasattr(store.model, "_features_rest"):
            store.model._features_rest = torch.nn.Parameter(g["_features_rest"].to(dev).requires_grad_(True))
    else:
        for k, v in g.items():
            v = v.to(dev)
            setattr(model, k, torch.nn.Parameter(v.requires_grad_(True)) if k not in ("face_index", "bary_coords") else v)
        model.num_points = P
        model.max_radii2D = torch.zeros(P, device=dev)
        model.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
        model.denom = torch.zeros(P, 1, device=dev)
        model.sample_flag = torch.zeros(P, device=dev)
    return missing, list(unexpected)
