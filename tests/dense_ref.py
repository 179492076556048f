"""Dense float64 PyTorch restatement of the splat render, differentiable by autograd.

Independent cross-check of the *hand-derived* backward (oracle C code and CUDA kernels): the forward is written
as plain tensor algebra over a [pixels, Gaussians] grid, with the reference's discrete decisions (tile-rect
inclusion, power>0 / alpha<1/255 skips, early termination) applied as constant masks and the 0.99 alpha clamp
as a straight-through op (the reference does not zero the gradient there, backward.cu:499).  Autograd of
this function is then the gradient the reference's backward implements, except inside the frustum clamp
region (backward.cu:175-176), which the test scenes avoid.

Small scenes only: memory is O(H*W*P).
"""
import math

import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def eval_sh(deg, sh, d):
    """sh [P,M,3], d [P,3] unit directions -> [P,3] (same basis as forward.cu:20-71)."""
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
               + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
               + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def dense_render(means3D, scales, rotations, opacities, shs, sh_degree, cam, bg, radii, scale_modifier=1.0,
                 colors_precomp=None, means2D_delta=None):
    """All tensor inputs float64 torch.  `radii` (int array from a forward pass) fixes which Gaussians are
    alive and their tile rectangles.  Returns color [3,H,W]."""
    dt = torch.float64
    W, H = cam["W"], cam["H"]
    V = torch.as_tensor(np.asarray(cam["viewmatrix"], np.float64), dtype=dt).reshape(4, 4)
    PM = torch.as_tensor(np.asarray(cam["projmatrix"], np.float64), dtype=dt).reshape(4, 4)
    campos = torch.as_tensor(np.asarray(cam["campos"], np.float64), dtype=dt)
    tanx, tany = cam["tanfovx"], cam["tanfovy"]
    fx, fy = W / (2 * tanx), H / (2 * tany)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    t = ph @ V
    hom = ph @ PM
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    if means2D_delta is not None:
        ndc = ndc + means2D_delta[:, :2]
    pix = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], 1)

    r, x, y, z = rotations[:, 0], rotations[:, 1], rotations[:, 2], rotations[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
    S = scale_modifier * scales
    RS = R * S[:, None, :]
    Sigma3 = RS @ RS.transpose(1, 2)

    tz = t[:, 2]
    limx, limy = 1.3 * tanx, 1.3 * tany
    txc = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    tyc = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * txc / (tz * tz), zero, fy / tz, -fy * tyc / (tz * tz)], 1).reshape(P, 2, 3)
    Rw2c = V[:3, :3].t()
    JW = J @ Rw2c
    cov2 = JW @ Sigma3 @ JW.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    conx, cony, conz = c / det, -b / det, a / det

    if colors_precomp is None:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp(eval_sh(sh_degree, shs, d) + 0.5, min=0.0)
    else:
        rgb = colors_precomp

    radii_t = torch.as_tensor(np.asarray(radii), dtype=torch.int64)
    alive = radii_t > 0
    order = torch.argsort(t[:, 2].detach().double() + 0.0, stable=True)
    # exact tie-break on float32 depth bits then index, as the reference
    depth32 = t[:, 2].detach().to(torch.float32)
    order = torch.from_numpy(np.lexsort((np.arange(P), depth32.numpy()))).long()

    gx, gy = (W + 15) // 16, (H + 15) // 16
    pix32 = pix.detach().to(torch.float32)
    rad = radii_t.to(torch.float32)
    x0 = torch.clamp(((pix32[:, 0] - rad) / 16).to(torch.int32), 0, gx)
    y0 = torch.clamp(((pix32[:, 1] - rad) / 16).to(torch.int32), 0, gy)
    x1 = torch.clamp(((pix32[:, 0] + rad + 15) / 16).to(torch.int32), 0, gx)
    y1 = torch.clamp(((pix32[:, 1] + rad + 15) / 16).to(torch.int32), 0, gy)

    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxf, pyf = xs.reshape(-1).to(dt), ys.reshape(-1).to(dt)
    tx_, ty_ = (xs.reshape(-1) // 16), (ys.reshape(-1) // 16)

    o = order
    dx = pix[o, 0][None, :] - pxf[:, None]
    dy = pix[o, 1][None, :] - pyf[:, None]
    power = -0.5 * (conx[o][None] * dx * dx + conz[o][None] * dy * dy) - cony[o][None] * dx * dy
    incl = (alive[o][None] & (tx_[:, None] >= x0[o][None]) & (tx_[:, None] < x1[o][None])
            & (ty_[:, None] >= y0[o][None]) & (ty_[:, None] < y1[o][None]))
    G = torch.exp(torch.clamp(power, max=0.0))
    alpha_raw = opacities.reshape(-1)[o][None] * G
    alpha = alpha_raw + (torch.clamp(alpha_raw, max=0.99) - alpha_raw).detach()
    valid = incl & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
    one_m = 1 - a_eff
    T_incl = torch.cumprod(one_m, dim=1)
    T_excl = torch.cat([torch.ones(T_incl.shape[0], 1, dtype=dt), T_incl[:, :-1]], 1)
    # termination: first contributing j with T_excl*(1-alpha) < 1e-4 stops the pixel (not applied)
    stop = valid & (T_incl.detach() < 1e-4)
    stopped = torch.cumsum(stop.to(torch.int32), dim=1) > 0
    w = torch.where(stopped, torch.zeros_like(a_eff), a_eff * T_excl)
    applied = valid & ~stopped
    T_final = torch.prod(torch.where(applied, one_m, torch.ones_like(one_m)), dim=1)
    col = w @ rgb[o]
    bg_t = torch.as_tensor(np.asarray(bg, np.float64), dtype=dt)
    out = col + T_final[:, None] * bg_t[None]
    return out.t().reshape(3, H, W)


def dense_gradients(scene, cam, radii, dL_dpix):
    """Autograd gradients of sum(color * dL_dpix) for a scene dict of numpy arrays (scales/rotations/opacities
    are the *activated* values the rasterizer receives)."""
    dt = torch.float64
    leaf = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=dt, requires_grad=True)
    m3, sc, ro, op, sh = leaf(scene["means3D"]), leaf(scene["scales"]), leaf(scene["rotations"]), \
        leaf(scene["opacities"]), leaf(scene["shs"])
    d2 = torch.zeros(m3.shape[0], 3, dtype=dt, requires_grad=True)
    color = dense_render(m3, sc, ro, op, sh, int(scene["sh_degree"]), cam, scene["bg"], radii, means2D_delta=d2)
    loss = (color * torch.as_tensor(np.asarray(dL_dpix, np.float64))).sum()
    loss.backward()
    return dict(color=color.detach().numpy(), dL_dmeans3D=m3.grad.numpy(), dL_dscales=sc.grad.numpy(),
                dL_drotations=ro.grad.numpy(), dL_dopacity=op.grad.numpy(), dL_dsh=sh.grad.numpy(),
                dL_dmeans2D=d2.grad.numpy())
