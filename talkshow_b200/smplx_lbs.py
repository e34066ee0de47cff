"""Batched SMPL-X evaluation on the device: the step right after the hot path in the reference's
``scripts/demo.py`` (``get_vertices``, :122-152 — one ``smplx_model(...)`` call PER FRAME in float64 on the CPU) and in
``data_utils/get_j.py`` (``get_joints``, :20-51).  Here all frames of all samples go through ONE engine call
(``ts_smplx_forward``: blend-shape GEMM + kinematic chain + skinning kernels, fp32).

The licensed model file (SMPLX_NEUTRAL_2020.npz) is not redistributable; ``load_smplx_npz`` reads it when the user has it,
``synthetic_model`` builds SMPL-X-shaped random tensors for tests and benchmarks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

NUM_BETAS, NUM_EXPR, NUM_JOINTS = 300, 100, 55
# vertex ids smplx appends as joints (smplx/vertex_ids.py, 'smplx' table, in VertexJointSelector order:
# nose, reye, leye, rear, lear, LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel, l/r thumb, index, middle, ring, pinky tips)
SMPLX_EXTRA_JOINT_VERTS = [9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
                           5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022]


def load_smplx_npz(path, device=None, engine=None):
    """SMPLX_*.npz of the SMPL-X release -> SmplxModel, with the reference's constructor arguments
    (scripts/demo.py:272-291: num_betas=300, num_expression_coeffs=100, use_pca=False, flat_hand_mean=False)."""
    d = np.load(path, allow_pickle=True, encoding="latin1")
    V = d["v_template"].shape[0]
    shapedirs = np.asarray(d["shapedirs"], dtype=np.float64)
    sd = np.concatenate([shapedirs[:, :, :NUM_BETAS], shapedirs[:, :, 300:300 + NUM_EXPR]], 2)
    posedirs = np.asarray(d["posedirs"], dtype=np.float64).reshape(V * 3, -1).T
    parents = np.asarray(d["kintree_table"][0], dtype=np.int64).copy()
    parents[0] = -1
    pose_mean = np.zeros(165)
    pose_mean[75:120] = d["hands_meanl"]
    pose_mean[120:165] = d["hands_meanr"]
    t = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a), dtype=dt)
    model = {"v_template": t(d["v_template"]), "shapedirs": t(sd), "posedirs": t(posedirs), "J_regressor": t(d["J_regressor"]),
             "parents": t(parents, torch.int64), "lbs_weights": t(d["weights"]), "pose_mean": t(pose_mean),
             "faces": t(d["f"].astype(np.int64), torch.int64), "lmk_faces_idx": t(d["lmk_faces_idx"].astype(np.int64), torch.int64),
             "lmk_bary_coords": t(d["lmk_bary_coords"]), "extra_joint_idx": torch.tensor(SMPLX_EXTRA_JOINT_VERTS)}
    return SmplxModel(model, device=device, engine=engine)


def synthetic_model(V=10475, seed=0, nfaces=20908):
    """SMPL-X-shaped random model tensors (float64): realistic magnitudes, a valid kinematic tree (the published SMPL-X
    parents table), skinning weights with 4 non-zero joints per vertex summing to one, a sparse joint regressor."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64) * 2 - 1
    parents = torch.tensor([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                            20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                            21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53])
    v_template = r(V, 3) * torch.tensor([0.4, 0.9, 0.15], dtype=torch.float64)
    shapedirs = r(V, 3, NUM_BETAS + NUM_EXPR) * 0.01
    posedirs = r(486, V * 3) * 0.005
    w = torch.zeros(V, NUM_JOINTS, dtype=torch.float64)
    idx = torch.randint(0, NUM_JOINTS, (V, 4), generator=g)
    w.scatter_(1, idx, torch.rand(V, 4, generator=g, dtype=torch.float64) + 0.05)
    w = w / w.sum(1, keepdim=True)
    Jr = torch.zeros(NUM_JOINTS, V, dtype=torch.float64)
    ji = torch.randint(0, V, (NUM_JOINTS, 32), generator=g)
    Jr.scatter_(1, ji, torch.rand(NUM_JOINTS, 32, generator=g, dtype=torch.float64))
    Jr = Jr / Jr.sum(1, keepdim=True)
    pose_mean = torch.zeros(165, dtype=torch.float64)
    pose_mean[75:165] = r(90) * 0.3
    faces = torch.randint(0, V, (nfaces, 3), generator=g)
    bary = torch.rand(51, 3, generator=g, dtype=torch.float64)
    bary = bary / bary.sum(1, keepdim=True)
    return {"v_template": v_template, "shapedirs": shapedirs, "posedirs": posedirs, "J_regressor": Jr, "parents": parents,
            "lbs_weights": w, "pose_mean": pose_mean, "faces": faces, "lmk_faces_idx": torch.randint(0, nfaces, (51,), generator=g),
            "lmk_bary_coords": bary, "extra_joint_idx": torch.randint(0, V, (21,), generator=g)}


class SmplxModel:
    """Callable like the reference's ``smplx_model`` (keyword arguments of demo.py:129-138 / get_j.py:21-29), any number of
    frames per call; returns an object with ``.vertices`` [F,V,3], ``.joints`` [F,127,3] and ``.body_pose`` on the device."""

    def __init__(self, model, device=None, engine=None):
        from .engine import Engine
        from .nets.base import shared_engine

        if engine is None:
            dev = torch.device("cuda", 0) if device is None else torch.device(device)
            engine = shared_engine(dev)
        self.e = engine
        self.device = engine.device
        self.batch_size = 1
        arr, keep = _lib.pack_tensors(model)
        engine._check(engine.L.ts_load_smplx(engine.h, arr, len(model)), "ts_load_smplx")
        v, j = C.c_int(0), C.c_int(0)
        engine._check(engine.L.ts_smplx_dims(engine.h, C.byref(v), C.byref(j)), "ts_smplx_dims")
        self.V, self.J = v.value, j.value

    def forward_pose265(self, pose265, betas=None, expression=True, want_vertices=True):
        """pose265 [F,265] -> (vertices [F,V,3] or None, joints [F,J,3]) on the device."""
        e = self.e
        p = e._dev(pose265, torch.float32)
        F = p.shape[0]
        b = None if betas is None else e._dev(betas.reshape(-1), torch.float32)
        verts = torch.empty(F, self.V, 3, device=self.device) if want_vertices else None
        joints = torch.empty(F, self.J, 3, device=self.device)
        e._check(e.L.ts_smplx_forward(e.h, _lib.ptr(p), _lib.ptr(b), int(bool(expression)), _lib.ptr(verts), _lib.ptr(joints), F,
                                      e._s()), "ts_smplx_forward")
        return verts, joints

    def __call__(self, betas=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None, global_orient=None,
                 body_pose=None, left_hand_pose=None, right_hand_pose=None, return_verts=True, **kw):
        F = body_pose.shape[0]
        z = lambda n: torch.zeros(F, n, device=body_pose.device, dtype=body_pose.dtype)
        use_expr = expression is not None and expression.shape[-1] == NUM_EXPR
        parts = [jaw_pose if jaw_pose is not None else z(3), leye_pose if leye_pose is not None else z(3),
                 reye_pose if reye_pose is not None else z(3), global_orient if global_orient is not None else z(3), body_pose,
                 left_hand_pose if left_hand_pose is not None else z(45), right_hand_pose if right_hand_pose is not None else z(45),
                 expression if use_expr else z(NUM_EXPR)]
        pose265 = torch.cat([x.reshape(F, -1).to(torch.float32) for x in parts], 1)
        b = None if betas is None else betas[:1]
        verts, joints = self.forward_pose265(pose265, b, expression=use_expr)
        out = {"vertices": verts, "joints": joints, "body_pose": body_pose}
        return _Out(out)


class _Out(dict):
    __getattr__ = dict.__getitem__


def get_vertices(smplx_model, betas, result_list, exp, require_pose=False, engine=None):
    """scripts/demo.py:122-152 with ONE batched call per sample instead of one model call per frame.
    result_list: tensors [F,265] -> list of numpy vertices (F, V, 3) (and body poses when require_pose)."""
    vertices_list, poses_list = [], []
    for res in result_list:
        verts, _ = smplx_model.forward_pose265(res, betas, expression=bool(exp))
        vertices_list.append(verts.cpu().numpy())
        poses_list.append(res[:, 12:75].detach().cpu())
    return (vertices_list, poses_list) if require_pose else (vertices_list, None)


def get_joints(smplx_model, betas, pred):
    """data_utils/get_j.py:32-51: pred [B,T,265] or [N,265] -> joints [B,T,J,3] / [N,J,3] (device tensor)."""
    flat = pred.reshape(-1, 265)
    _, joints = smplx_model.forward_pose265(flat, betas, expression=True, want_vertices=False)
    return joints.reshape(tuple(pred.shape[:-1]) + (joints.shape[1], 3))
