"""SURVEY.md §8 f4 — batched SMPL-X evaluation.  CPU: the float64 restatement of smplx 0.1.28 (oracle/smplx_oracle.py)
satisfies the algorithm's invariants.  GPU (-m gpu): ts_smplx_forward vs that restatement on SMPL-X-shaped synthetic
tensors (the licensed model file and the smplx package are absent: parity is against the restatement, see its header)."""
import numpy as np
import pytest
import torch

import smplx_oracle as SO
from talkshow_b200 import smplx_lbs


def _poses(F, seed, scale=0.4):
    g = torch.Generator().manual_seed(seed)
    p = (torch.rand(F, 265, generator=g, dtype=torch.float64) * 2 - 1) * scale
    p[:, 165:] *= 2.0            # expression coefficients
    return p


def test_oracle_invariants():
    m = smplx_lbs.synthetic_model(V=500, seed=3, nfaces=900)
    F = 4
    p = _poses(F, 1)
    # rest pose: every rotation is the identity when the pose cancels the hand mean -> vertices = shaped template,
    # joints = J_regressor . vertices, landmarks on the un-posed mesh
    rest = torch.zeros(F, 265, dtype=torch.float64)
    rest[:, 75:165] = -m["pose_mean"][75:165]
    rest[:, 165:] = p[:, 165:]
    verts, joints = SO.smplx_forward(m, rest)
    shaped = m["v_template"][None] + torch.einsum("bl,mkl->bmk", torch.cat([torch.zeros(F, 300, dtype=torch.float64), rest[:, 165:]], 1),
                                                  m["shapedirs"])
    assert (verts - shaped).abs().max() < 1e-6          # 1e-8 epsilon of batch_rodrigues leaves ~1e-8 rotations
    assert (joints[:, :55] - torch.einsum("bik,ji->bjk", shaped, m["J_regressor"])).abs().max() < 1e-6
    assert joints.shape == (F, 55 + 21 + 51, 3)
    # a global rotation rotates the whole posed body about the root joint
    a = p.clone()
    b = p.clone()
    b[:, 9:12] = torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64)
    a[:, 9:12] = 0
    va, ja = SO.smplx_forward(m, a)
    vb, jb = SO.smplx_forward(m, b)
    R = SO.batch_rodrigues(b[:1, 9:12])[0]
    root = ja[:, :1, :]                                   # root joint is fixed by the global rotation
    assert ((va - root) @ R.T + root - vb).abs().max() < 1e-9
    assert ((ja - root) @ R.T + root - jb).abs().max() < 1e-9
    # frames are independent
    v2, _ = SO.smplx_forward(m, p[1:3])
    vall, _ = SO.smplx_forward(m, p)
    assert (v2 - vall[1:3]).abs().max() < 1e-12


@pytest.mark.gpu
def test_smplx_forward_matches_oracle():
    from talkshow_b200.engine import Engine

    m = smplx_lbs.synthetic_model(V=10475, seed=0)
    e = Engine(0)
    try:
        sm = smplx_lbs.SmplxModel(m, engine=e)
        assert (sm.V, sm.J) == (10475, 127)
        F = 37
        p = _poses(F, 7)
        betas = (torch.rand(1, 300, generator=torch.Generator().manual_seed(2), dtype=torch.float64) * 2 - 1) * 0.5
        for bt, use_expr in ((None, True), (betas, True), (None, False)):
            ref_v, ref_j = SO.smplx_forward(m, p, bt, use_expr)
            v, j = sm.forward_pose265(p, bt, expression=use_expr)
            ev = (v.cpu().double() - ref_v).abs().max().item()
            ej = (j.cpu().double() - ref_j).abs().max().item()
            print("smplx LBS max-abs err: vertices %.2e, joints %.2e (|v| max %.2f)" % (ev, ej, ref_v.abs().max().item()))
            assert ev <= 1e-4 and ej <= 1e-4
        # the reference's keyword interface (demo.py:129-138), one frame and many frames
        out = sm(betas=torch.zeros(1, 300), expression=p[:, 165:265], jaw_pose=p[:, 0:3], leye_pose=p[:, 3:6], reye_pose=p[:, 6:9],
                 global_orient=p[:, 9:12], body_pose=p[:, 12:75], left_hand_pose=p[:, 75:120], right_hand_pose=p[:, 120:165],
                 return_verts=True)
        ref_v, _ = SO.smplx_forward(m, p, None, True)
        assert (out.vertices.cpu().double() - ref_v).abs().max().item() <= 1e-4
        # demo.py get_vertices / get_j.py get_joints surfaces
        vl, _ = smplx_lbs.get_vertices(sm, None, [p[:5].float().cuda(), p[5:9].float().cuda()], True)
        assert vl[0].shape == (5, 10475, 3) and np.abs(vl[1] - ref_v[5:9].numpy()).max() <= 1e-4
        jj = smplx_lbs.get_joints(sm, None, p[:36].float().reshape(3, 12, 265).cuda())
        assert jj.shape == (3, 12, 127, 3)
        # more frames than one GEMM chunk (4096)
        big = _poses(4100, 11)
        vb, _ = sm.forward_pose265(big)
        rv, _ = SO.smplx_forward(m, big[4090:])
        assert (vb[4090:].cpu().double() - rv).abs().max().item() <= 1e-4
    finally:
        torch.cuda.synchronize()
        e.close()


def test_host_mirror_matches_reference_call_semantics():
    """get_vertices / get_joints / the keyword call of talkshow_b200.smplx_lbs against the reference's own get_vertices
    (scripts/demo.py:122-152) and get_joints (data_utils/get_j.py:20-51), both answered by the same float64 restatement of the body
    model (fixture: tests/golden/make_golden.py --only smplx_calls): which columns of the 265-vector feed which argument, one
    batched call instead of one per frame, output layouts."""
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "smplx_calls.npz"))
    model = smplx_lbs.synthetic_model(V=int(g["V"]), seed=int(g["model_seed"]), nfaces=int(g["nfaces"]))

    class Model(smplx_lbs.SmplxModel):                     # the engine call replaced by the oracle; everything else is the product's
        def __init__(self):
            self.device, self.batch_size, self.calls = torch.device("cpu"), 1, 0

        def forward_pose265(self, pose265, betas=None, expression=True, want_vertices=True):
            self.calls += 1
            v, j = SO.smplx_forward(model, pose265, betas, use_expression=expression)
            return (v if want_vertices else None), j

    gen = torch.Generator().manual_seed(int(g["pose_seed"]))
    res = [((torch.rand(5, 265, generator=gen, dtype=torch.float64) * 2 - 1) * 0.4) for _ in range(2)]
    betas = (torch.rand(1, 300, generator=gen, dtype=torch.float64) - 0.5) * 0.2
    m = Model()
    verts, poses = smplx_lbs.get_vertices(m, betas, res, True, require_pose=True)
    assert m.calls == 2                                    # one call per sample, not per frame
    assert np.abs(np.stack(verts) - g["verts"]).max() <= 1e-6 and np.abs(torch.stack(poses).numpy() - g["poses"]).max() <= 1e-7
    j3 = smplx_lbs.get_joints(m, betas, torch.stack(res))
    j2 = smplx_lbs.get_joints(m, betas, res[0])
    assert j3.shape == (2, 5, 127, 3) and j2.shape == (5, 127, 3)
    assert np.abs(j3.numpy() - g["joints3"]).max() <= 1e-6 and np.abs(j2.numpy() - g["joints2"]).max() <= 1e-6
    # the keyword call of the reference's smplx_model (demo.py:129-138), many frames at once
    p = res[1]
    out = m(betas=betas, expression=p[:, 165:265], jaw_pose=p[:, 0:3], leye_pose=p[:, 3:6], reye_pose=p[:, 6:9], global_orient=p[:, 9:12],
            body_pose=p[:, 12:75], left_hand_pose=p[:, 75:120], right_hand_pose=p[:, 120:165], return_verts=True)
    assert np.abs(out.vertices.numpy() - g["verts"][1]).max() <= 1e-5 and torch.equal(out.body_pose, p[:, 12:75])
    assert np.abs(out["joints"].numpy() - g["joints3"][1]).max() <= 1e-5
