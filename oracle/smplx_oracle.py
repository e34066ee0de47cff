"""CPU oracle for the batched SMPL-X evaluation.  TEST INFRASTRUCTURE ONLY.

The reference evaluates the body model with the third-party package ``smplx`` (requirements.txt:5 pins
smplx~=0.1.28), one Python call PER FRAME in float64 (scripts/demo.py:122-152 ``get_vertices``;
data_utils/get_j.py:20-51 ``get_joints``).  The package is not vendored in /root/reference and not installed here, so
this file restates its published algorithm — ``smplx.body_models.SMPLX.forward`` + ``smplx.lbs.lbs`` /
``batch_rodrigues`` / ``batch_rigid_transform`` / ``vertices2landmarks`` / ``VertexJointSelector`` of release 0.1.28 —
in float64 torch, with the constructor arguments of the reference's call site (scripts/demo.py:272-291):
``use_pca=False, flat_hand_mean=False, num_betas=300, num_expression_coeffs=100, use_face_contour=False``.
Parity is "unpinned" in the sense of the task statement: neither the package nor the licensed model file
(SMPLX_NEUTRAL_2020.npz) is available, so there is no golden vector; the CUDA path is compared with this restatement on
synthetic SMPL-X-shaped tensors (talkshow_b200/smplx_lbs.py:synthetic_model).

Model dict (float64 unless noted), shapes of the published model file:
  v_template [V,3], shapedirs [V,3,400] (300 shape + 100 expression components), posedirs [486, V*3]
  (= reshape(posedirs_file [V,3,486], [V*3,486]).T), J_regressor [55,V], parents [55] int64 (parents[0] = -1),
  lbs_weights [V,55], pose_mean [165] (zeros except the two 45-dim hand means), faces [Fc,3] int64,
  lmk_faces_idx [51] int64, lmk_bary_coords [51,3], extra_joint_idx [21] int64 (vertex ids appended as joints).
"""
from __future__ import annotations

import torch


def batch_rodrigues(rot_vecs):
    """smplx/lbs.py batch_rodrigues: [N,3] axis-angle -> [N,3,3]; note the 1e-8 added to the VECTOR before the norm."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle)[:, None]
    sin = torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros_like(rx)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """smplx/lbs.py batch_rigid_transform: -> (posed joints [B,J,3], relative transforms A [B,J,4,4])."""
    B, J = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    tm = torch.zeros(B, J, 4, 4, dtype=joints.dtype)
    tm[:, :, :3, :3] = rot_mats
    tm[:, :, :3, 3:] = rel
    tm[:, :, 3, 3] = 1
    chain = [tm[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    tr = torch.stack(chain, dim=1)
    posed = tr[:, :, :3, 3]
    jh = torch.nn.functional.pad(joints, [0, 0, 0, 1])
    rel_tr = tr - torch.nn.functional.pad(torch.matmul(tr, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, rel_tr


def full_pose_from_265(pose265):
    """Argument order of the reference's call (demo.py:129-138 / get_j.py:21-29) -> SMPLX.forward's
    full_pose = [global_orient, body_pose(21), jaw, leye, reye, left_hand(15), right_hand(15)] (165)."""
    p = pose265
    return torch.cat([p[:, 9:12], p[:, 12:75], p[:, 0:3], p[:, 3:6], p[:, 6:9], p[:, 75:120], p[:, 120:165]], 1)


def smplx_forward(model, pose265, betas=None, use_expression=True):
    """pose265 [F,265] (jaw | leye | reye | global | body | lhand | rhand | expression) -> (vertices [F,V,3],
    joints [F,127,3]) as smplx.SMPLX.forward(return_verts=True) returns them (55 LBS joints + 21 vertex joints +
    51 static landmarks).  betas [1,300] or None (zeros, demo.py:159); use_expression False = zero expression
    (get_vertices(exp=False) passes a 50-dim zero vector, i.e. no expression offset)."""
    m = {k: (v.double() if torch.is_floating_point(v) else v) for k, v in model.items()}
    pose265 = pose265.double()
    F = pose265.shape[0]
    full_pose = full_pose_from_265(pose265) + m["pose_mean"][None]
    bet = torch.zeros(1, 300, dtype=torch.float64) if betas is None else betas.double().reshape(1, 300)
    expr = pose265[:, 165:265] if use_expression else torch.zeros(F, 100, dtype=torch.float64)
    shape_components = torch.cat([bet.expand(F, -1), expr], 1)                                  # [F,400]
    v_shaped = m["v_template"][None] + torch.einsum("bl,mkl->bmk", shape_components, m["shapedirs"])
    J = torch.einsum("bik,ji->bjk", v_shaped, m["J_regressor"])
    rot = batch_rodrigues(full_pose.reshape(-1, 3)).view(F, -1, 3, 3)
    pose_feature = (rot[:, 1:] - torch.eye(3, dtype=torch.float64)).reshape(F, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, m["posedirs"]).view(F, -1, 3)
    J_tr, A = batch_rigid_transform(rot, J, m["parents"])
    T = torch.matmul(m["lbs_weights"][None].expand(F, -1, -1), A.view(F, -1, 16)).view(F, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(F, v_posed.shape[1], 1, dtype=torch.float64)], 2)
    verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
    # VertexJointSelector + static landmarks (vertices2landmarks: barycentric interpolation on the landmark faces)
    extra = verts[:, m["extra_joint_idx"]]
    tri = m["faces"][m["lmk_faces_idx"]]                                                       # [51,3]
    lmk = torch.einsum("blfi,lf->bli", verts[:, tri], m["lmk_bary_coords"])
    return verts, torch.cat([J_tr, extra, lmk], 1)
