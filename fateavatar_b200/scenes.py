"""Deterministic synthetic workloads for tests and bench.py (BASELINE.json configs 1, 2, 5).

Everything here is *input generation*: numpy float64 math cast to float32.  Camera matrices follow the
reference's conventions (tools/gs_utils/graphics_utils.py:38-84, volume_rendering/camera_3dgs.py:53-72):
`viewmatrix` = world-to-view transposed, `projmatrix` = viewmatrix @ projection^T, points are row vectors.
"""
import math

import numpy as np

SH_C0 = 0.28209479177387814


def make_camera(W, H, fovx, fovy, R=None, T=None, znear=0.01, zfar=100.0):
    """R: camera-to-world rotation (3x3), T: world-to-camera translation (3,), as `cam_pose` holds them
    (SURVEY Appendix A).  Returns the tensors GaussianRasterizationSettings needs."""
    R = np.diag([1.0, -1.0, -1.0]) if R is None else np.asarray(R, np.float64)
    T = np.array([0.0, 0.0, 2.5]) if T is None else np.asarray(T, np.float64)
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    view_t = Rt.T  # world_view_transform
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    full = view_t @ Pm.T
    campos = np.linalg.inv(view_t)[3, :3]
    return dict(W=int(W), H=int(H), fovx=float(fovx), fovy=float(fovy), tanfovx=math.tan(fovx * 0.5),
                tanfovy=math.tan(fovy * 0.5), viewmatrix=view_t.astype(np.float32),
                projmatrix=full.astype(np.float32), campos=campos.astype(np.float32))


def orbit_camera(W, H, fov, k, n, radius=2.5):
    """k-th of n views on a horizontal circle looking at the origin (the 360 degree sweep of config 5; same
    parametrisation as LookAtPoseSampler.sample(pi/2 + 2 pi k/n, pi/2, 0, radius), camera_eg3d.py:36-54)."""
    theta = math.pi / 2 + 2 * math.pi * k / n
    origin = np.array([radius * math.cos(math.pi - theta), 0.0, radius * math.sin(math.pi - theta)])
    fwd = -origin / np.linalg.norm(origin)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    up2 = np.cross(fwd, right)
    c2w_R = np.stack([right, -up2, fwd], axis=1)  # camera x right, y down, z forward
    w2c_R = c2w_R.T
    T = -w2c_R @ origin
    return make_camera(W, H, fov, fov, R=c2w_R, T=T)


def _quat(rng, n):
    q = rng.standard_normal((n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def config1_scene(seed=0, P=10000, W=256, H=256):
    """BASELINE config 1: 10k random Gaussians, one static 256x256 camera, SH degree 0, white background."""
    rng = np.random.default_rng(seed)
    s = dict(
        means3D=rng.uniform(-0.5, 0.5, (P, 3)),
        scales=np.exp(rng.normal(math.log(0.01), 0.3, (P, 3))),
        rotations=_quat(rng, P),
        opacities=_sigmoid(rng.standard_normal((P, 1))),
        shs=((rng.uniform(0, 1, (P, 1, 3)) - 0.5) / SH_C0),
        sh_degree=0,
        bg=np.ones(3),
        camera=make_camera(W, H, 0.35, 0.35, T=[0, 0, 2.5]),
        name=f"config1: {P} random Gaussians, {W}x{H}, SH0",
    )
    return _f32(s)


def head_points(rng, P):
    """Points on a head-sized ellipsoid shell (extents ~0.21 x 0.31 x 0.22 like the FLAME template)."""
    d = rng.standard_normal((P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    semi = np.array([0.105, 0.155, 0.11])
    return d * semi * (1.0 + 0.01 * rng.standard_normal((P, 1)))


def head_scene(seed=0, P=100000, W=512, H=512, sh_degree=0, scale_mult=1.0):
    """BASELINE config 2 geometry: ~100k splats on a head-sized surface filling ~70 % of a 512x512 frame
    (camera distance 1.25, fov 0.35 rad).  Scale = mean nearest-neighbour spacing (FateAvatar's init rule,
    model/fateavatar.py:596-608) with log-normal anisotropy; opacity = sigmoid(N(0,1))."""
    rng = np.random.default_rng(seed)
    xyz = head_points(rng, P)
    area = 4 * math.pi * (((0.105 * 0.155) ** 1.6075 + (0.105 * 0.11) ** 1.6075 + (0.155 * 0.11) ** 1.6075) / 3) ** (
        1 / 1.6075)
    spacing = 0.5 * math.sqrt(area / P) * scale_mult
    M = (sh_degree + 1) ** 2
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0, 1, (P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0, 0.3, (P, M - 1, 3))
    s = dict(
        means3D=xyz,
        scales=np.exp(rng.normal(math.log(spacing), 0.3, (P, 3))),
        rotations=_quat(rng, P),
        opacities=_sigmoid(rng.standard_normal((P, 1))),
        shs=shs, sh_degree=sh_degree, bg=np.ones(3),
        camera=make_camera(W, H, 0.35, 0.35, T=[0, 0, 1.25]),
        name=f"config2: {P} head-surface Gaussians, {W}x{H}, SH{sh_degree}",
    )
    return _f32(s)


def stress_scene(seed=0, P=500000, W=1024, H=1024, view=0, n_views=72):
    """BASELINE config 5: 500k Gaussians, SH degree 3, 1024x1024, orbit view `view` of `n_views`; 5 % of the
    splats are large (stresses tile rectangles, per-tile list length and the large-tile sort path)."""
    rng = np.random.default_rng(seed)
    xyz = head_points(rng, P) * 4.0  # object ~0.9 m wide seen from 2.5 m
    base = 0.5 * math.sqrt(4 * math.pi * 0.45 ** 2 / P)
    log_s = rng.normal(math.log(base), 0.4, (P, 3))
    big = rng.uniform(size=P) < 0.05
    log_s[big] += math.log(8.0)
    shs = np.zeros((P, 16, 3))
    shs[:, 0, :] = (rng.uniform(0, 1, (P, 3)) - 0.5) / SH_C0
    shs[:, 1:, :] = rng.normal(0, 0.3, (P, 15, 3))
    s = dict(
        means3D=xyz, scales=np.exp(log_s), rotations=_quat(rng, P),
        opacities=_sigmoid(rng.standard_normal((P, 1))), shs=shs, sh_degree=3, bg=np.zeros(3),
        camera=orbit_camera(W, H, 0.35, view, n_views, radius=2.5),
        name=f"config5: {P} Gaussians, {W}x{H}, SH3, orbit view {view}/{n_views}",
    )
    return _f32(s)


def _f32(s):
    for k, v in list(s.items()):
        if isinstance(v, np.ndarray):
            s[k] = np.ascontiguousarray(v, dtype=np.float32)
    return s


def to_torch(scene, device):
    """Scene dict -> torch tensors on `device` + a GaussianRasterizationSettings-ready camera dict."""
    import torch

    out = {}
    for k, v in scene.items():
        if isinstance(v, np.ndarray):
            out[k] = torch.from_numpy(v).to(device)
        elif isinstance(v, dict):
            out[k] = {kk: (torch.from_numpy(vv).to(device) if isinstance(vv, np.ndarray) else vv) for kk, vv in v.items()}
        else:
            out[k] = v
    return out


def ellipsoid_mesh(n_lat=51, n_lon=100, semi=(0.105, 0.155, 0.11)):
    """Closed triangle mesh with FLAME-like size (V = (n_lat-1)*n_lon + 2 = 5002, F = 2*n_lon*(n_lat-1) = 10000
    by default): stand-in for the licensed FLAME template in the pose-stage tests and benchmarks."""
    verts = [(0.0, 1.0, 0.0)]
    for i in range(1, n_lat):
        th = math.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * math.pi * j / n_lon
            verts.append((math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)))
    verts.append((0.0, -1.0, 0.0))
    verts = np.asarray(verts, np.float64) * np.asarray(semi)
    faces = []
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)
    for j in range(n_lon):
        faces.append((0, ring(1, j + 1), ring(1, j)))
        faces.append((len(verts) - 1, ring(n_lat - 1, j), ring(n_lat - 1, j + 1)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces.append((a, b, c))
            faces.append((b, d, c))
    return verts.astype(np.float32), np.asarray(faces, np.int64)


def pose_inputs(N=100000, seed=0):
    """Synthetic inputs of the per-splat pose stage: a posed mesh (canonical ellipsoid + smooth deformation),
    N splat sites (face_index, barycentrics) and raw per-splat parameters as FateAvatar stores them."""
    rng = np.random.default_rng(seed)
    canon_verts, faces = ellipsoid_mesh()
    bend = 0.08 * np.sin(6.0 * canon_verts[:, [1, 2, 0]]) * canon_verts + 0.002 * rng.standard_normal(canon_verts.shape)
    verts = (canon_verts + bend).astype(np.float32)
    # splat sites are spread uniformly over the surface (FateAvatar samples the template uniformly in UV space,
    # volume_rendering/mesh_sampling.py:86-138): faces are drawn proportionally to their area
    tri = canon_verts[faces]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1).astype(np.float64)
    face_index = rng.choice(faces.shape[0], size=N, p=area / area.sum()).astype(np.int64)
    bary = rng.dirichlet(np.ones(3), N).astype(np.float32)
    return dict(
        verts=verts, canon_verts=canon_verts, faces=faces, face_index=face_index, bary=bary,
        scaling_raw=np.log(np.full((N, 3), 8e-4)).astype(np.float32) + 0.3 * rng.standard_normal((N, 3)).astype(np.float32),
        rotation_raw=(np.array([1, 0, 0, 0], np.float32) + 0.5 * rng.standard_normal((N, 4))).astype(np.float32),
        offset_raw=(0.5 * rng.standard_normal((N, 1))).astype(np.float32),
        opacity_raw=rng.standard_normal((N, 1)).astype(np.float32),
        shell_len=0.05,
    )


def _smooth_fields(rng, x, n, amp):
    """n smooth vector fields sampled at the points x [V,3]: amp_l * sin(f_l <x, u_l> + phi_l), f in [5, 30] rad/m --
    blendshape-like (neighbouring vertices move together), so posed triangles stay well shaped.  -> [V, 3, n]"""
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    f, phi = rng.uniform(5.0, 30.0, n), rng.uniform(0, 2 * np.pi, n)
    a = amp * rng.standard_normal((n, 3))
    s = np.sin((x @ u.T) * f + phi)  # [V, n]
    return s[:, None, :] * a.T[None, :, :]


def _smooth_skin_weights(rng, x, J):
    """Skinning weights that vary smoothly over the surface (softmax of distances to J random centres)."""
    c = x[rng.choice(x.shape[0], J, replace=False)] * 0.7
    d2 = ((x[:, None, :] - c[None, :, :]) ** 2).sum(-1)
    return np.exp(-d2 / 0.08 ** 2) + 1e-3


def flame_inputs(seed=0, V=None, n_shape=300, n_exp=100, J=5, with_deltas=True, v_template=None):
    """Synthetic FLAME-shaped model + one frame's coefficients (SURVEY Appendix C recipe): the licensed FLAME
    pickle cannot be shipped, so the buffers flame/FLAME.py:72-107 registers are drawn at FLAME's sizes and
    magnitudes (smooth blendshape fields of a few millimetres).  V=None uses the 5002-vertex ellipsoid template of
    `ellipsoid_mesh` (FLAME: 5023)."""
    rng = np.random.default_rng(seed)
    if v_template is not None:  # e.g. the vertices of weights/head_template_mouth_close.obj
        v_template = np.asarray(v_template, np.float64)
        V = v_template.shape[0]
    elif V is None:
        v_template, _ = ellipsoid_mesh()
        V = v_template.shape[0]
    else:
        d = rng.standard_normal((V, 3))
        v_template = d / np.linalg.norm(d, axis=1, keepdims=True) * np.array([0.105, 0.155, 0.11])
    v_template = v_template - v_template.mean(0, keepdims=True)
    L, NP = n_shape + n_exp, (J - 1) * 9
    jr = rng.uniform(size=(J, V)) * (rng.uniform(size=(J, V)) < 0.02)
    jr[:, 0] += 1e-3
    parents = np.array([-1, 0, 1, 1, 1, 1, 1, 1][:J], np.int64)
    pose = 0.05 * rng.standard_normal(J * 3)
    if J >= 3:
        pose[6] = 0.2  # jaw, cf. canonical_pose in config/fateavatar.yaml
    betas = np.concatenate([np.zeros(n_shape), 0.5 * rng.standard_normal(n_exp)])
    posed = lambda amp: _smooth_fields(rng, v_template, NP, amp).reshape(V * 3, NP).T  # [NP, 3V]
    out = dict(
        v_template=v_template, shapedirs=_smooth_fields(rng, v_template, L, 1e-3), posedirs=posed(1e-3),
        J_regressor=jr / jr.sum(1, keepdims=True),
        lbs_weights=_smooth_skin_weights(rng, v_template, J), parents=parents, betas=betas, pose=pose,
        n_shape=n_shape, n_exp=n_exp,
    )
    out["lbs_weights"] = out["lbs_weights"] / out["lbs_weights"].sum(1, keepdims=True)
    if with_deltas:
        out["delta_shapedirs"] = _smooth_fields(rng, v_template, L, 2e-4)
        out["delta_posedirs"] = posed(2e-4)
        out["delta_vertex"] = _smooth_fields(rng, v_template, 1, 1e-3)[:, :, 0]
    out = _f32(out)
    out["parents"] = parents
    return out


def small_avatar(seed=0, n_lat=9, n_lon=16, N=900):
    """A complete miniature avatar for whole-frame tests: FLAME-shaped model on a small ellipsoid mesh (V = (n_lat-1) *
    n_lon + 2 vertices), N splat sites on it and raw splat parameters sized for a 64x80 render at distance 1.25."""
    rng = np.random.default_rng(seed)
    verts, faces = ellipsoid_mesh(n_lat=n_lat, n_lon=n_lon)
    f = flame_inputs(seed=seed, V=verts.shape[0])
    f["v_template"] = (verts - verts.mean(0, keepdims=True)).astype(np.float32)
    tri = verts[faces]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1).astype(np.float64)
    f.update(
        faces=faces, face_index=rng.choice(faces.shape[0], size=N, p=area / area.sum()).astype(np.int64),
        bary=rng.dirichlet(np.ones(3), N).astype(np.float32),
        scaling_raw=(np.log(6e-3) + 0.3 * rng.standard_normal((N, 3))).astype(np.float32),
        rotation_raw=(np.array([1, 0, 0, 0], np.float32) + 0.5 * rng.standard_normal((N, 4))).astype(np.float32),
        offset_raw=(0.5 * rng.standard_normal((N, 1))).astype(np.float32), opacity_raw=rng.standard_normal((N, 1)).astype(np.float32),
        features_dc=((rng.uniform(0, 1, (N, 1, 3)) - 0.5) / SH_C0).astype(np.float32), shell_len=0.05)
    return f


def read_obj(path):
    """(verts [V,3] float32, faces [F,3] int64, 0-based) of a triangulated Wavefront OBJ; only `v` and `f` lines are
    read (pytorch3d.io.load_obj's verts / faces.verts_idx for the template mesh, model/fateavatar.py:123-127)."""
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
                for k in range(1, len(idx) - 1):  # fan-triangulate (the template is already triangles)
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, np.float32), np.asarray(faces, np.int64)


def template_avatar(verts, faces, N=100000, seed=0, scale=8e-4):
    """SURVEY 8d's config-2 avatar on a given template mesh (the reference's head_template_mouth_close.obj: 5023
    vertices, 10006 faces): FLAME-shaped model whose v_template is the re-centred mesh, N area-uniform splat sites,
    raw splat parameters as FateAvatar stores them."""
    rng = np.random.default_rng(seed)
    f = flame_inputs(seed=seed, v_template=verts)
    tri = f["v_template"][faces].astype(np.float64)
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    f.update(
        faces=faces, face_index=rng.choice(faces.shape[0], size=N, p=area / area.sum()).astype(np.int64),
        bary=rng.dirichlet(np.ones(3), N).astype(np.float32),
        scaling_raw=(np.log(scale) + 0.3 * rng.standard_normal((N, 3))).astype(np.float32),
        rotation_raw=(np.array([1, 0, 0, 0], np.float32) + 0.5 * rng.standard_normal((N, 4))).astype(np.float32),
        offset_raw=(0.5 * rng.standard_normal((N, 1))).astype(np.float32), opacity_raw=rng.standard_normal((N, 1)).astype(np.float32),
        features_dc=((rng.uniform(0, 1, (N, 1, 3)) - 0.5) / SH_C0).astype(np.float32), shell_len=0.05)
    return f
