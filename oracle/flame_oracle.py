"""Oracle of the FLAME skinning stage.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch restatement of

  flame/lbs.py:24-100      lbs(): v_shaped = template + einsum('bl,mkl->bmk', betas, shapedirs); joints;
                           pose_feature = (R[1:] - I); pose_offsets = pose_feature @ posedirs; kinematic chain;
                           T = W @ A; verts = (T @ [v_posed; 1])[:3]
  flame/lbs.py:232-270     batch_rodrigues (angle = ||r + 1e-8||)
  flame/lbs.py:285-342     batch_rigid_transform (rel_transforms = transforms - pad(transforms @ [J; 0]))
  flame/FLAME.py:131-204   forward / forward_with_delta_blendshape (betas = zeros(n_shape) ++ expression;
                           deltas added to template / shapedirs / posedirs before lbs)

Gradients come from autograd of this restatement (float64 for a clean reference).  Pinned against the reference's
own flame/lbs.py (imported on CPU where /root/reference is mounted: tests/test_flame.py, and through the committed
fixture tests/golden/flame_small.npz made by tests/golden/make_flame_golden.py).
"""
import torch


def batch_rodrigues(rot_vecs):  # lbs.py:232-270
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros_like(rx)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def rigid_transforms(rot_mats, joints, parents):  # lbs.py:285-342, batch of one
    Jn = joints.shape[0]
    rel = joints.clone()
    rel[1:] = joints[1:] - joints[parents[1:]]
    M = torch.zeros(Jn, 4, 4, dtype=joints.dtype)
    M[:, :3, :3] = rot_mats
    M[:, :3, 3] = rel
    M[:, 3, 3] = 1.0
    chain = [M[0]]
    for i in range(1, Jn):
        chain.append(chain[int(parents[i])] @ M[i])
    G = torch.stack(chain)
    jh = torch.cat([joints, torch.zeros(Jn, 1, dtype=joints.dtype)], dim=1)[..., None]  # [J,4,1], w = 0
    A = G - torch.nn.functional.pad(G @ jh, [3, 0])
    return G[:, :3, 3], A


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """betas [L], pose [J*3], v_template [V,3], shapedirs [V,3,L], posedirs [(J-1)*9, 3V], J_regressor [J,V],
    parents [J] long, lbs_weights [V,J] -> verts [V,3], pose_feature [(J-1)*9], A [J,4,4]."""
    v_shaped = v_template + torch.einsum("l,mkl->mk", betas, shapedirs)
    J = J_regressor @ v_shaped
    R = batch_rodrigues(pose.view(-1, 3))
    pose_feature = (R[1:] - torch.eye(3, dtype=R.dtype)).reshape(-1)
    v_posed = (pose_feature @ posedirs).view(-1, 3) + v_shaped
    _, A = rigid_transforms(R, J, parents)
    T = (lbs_weights @ A.view(-1, 16)).view(-1, 4, 4)
    vh = torch.cat([v_posed, torch.ones_like(v_posed[:, :1])], dim=1)
    verts = (T @ vh[..., None])[:, :3, 0]
    return verts, pose_feature, A


def forward_with_delta_blendshape(m, betas, pose, delta_shapedirs=None, delta_posedirs=None, delta_vertex=None):
    """m: dict of model tensors (v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights)."""
    vt = m["v_template"] if delta_vertex is None else m["v_template"] + delta_vertex
    sd = m["shapedirs"] if delta_shapedirs is None else m["shapedirs"] + delta_shapedirs
    pd = m["posedirs"] if delta_posedirs is None else m["posedirs"] + delta_posedirs
    return lbs(betas, pose, vt, sd, pd, m["J_regressor"], m["parents"], m["lbs_weights"])
