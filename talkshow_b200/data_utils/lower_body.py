"""SMPL-X pose layout helpers of the reference (data_utils/lower_body.py): which of the 165
axis-angle dims the body model generates, and the fixed lower-body block that ``part2full``
re-inserts.  The device version of part2full is csrc/api.cu:assemble_kernel (ts_assemble_pose)."""
import numpy as np
import torch

# values the reference hard-codes for the seated lower body (lower_body.py:4-8)
LOWER_POSE = torch.tensor(
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0747, -0.0158, -0.0152, -1.1826512813568115, 0.23866955935955048,
     0.15146760642528534, -1.2604516744613647, -0.3160211145877838, -0.1603458970785141, 1.1654603481292725, 0.0, 0.0,
     1.2521806955337524, 0.041598282754421234, -0.06312154978513718] + [0.0] * 12)

_FIXED_3D = set(range(18)) | set(range(21, 27)) | set(range(30, 36)) | set(range(45, 51))
c_index_3d = np.asarray([i for i in range(165) if i not in _FIXED_3D])      # 129 generated dims (:44-56)


def part2full(pred, stand=False):
    """[F,232] (jaw3 | body+hands129 | expr100) -> [F,265] (lower_body.py:68-87)."""
    lp = LOWER_POSE.to(pred)
    if stand:
        lp = torch.zeros_like(lp)
        lp[6:9] = torch.tensor([3.0747, -0.0158, -0.0152]).to(pred)
    lp = lp[None].expand(pred.shape[0], -1)
    return torch.cat([pred[:, :3], lp[:, :15], pred[:, 3:6], lp[:, 15:21], pred[:, 6:9], lp[:, 21:27], pred[:, 9:12],
                      lp[:, 27:], pred[:, 12:]], dim=1)


c_index_6d = np.asarray([v for i in c_index_3d for v in (2 * i, 2 * i + 1)])   # the same joints in the 6-D layout (:58-65)

# the standing lower body the reference's poses2pred(stand=True) inserts (lower_body.py:9-16)
LOWER_POSE_STAND = torch.tensor(
    [8.9759e-04, 7.1074e-04, -5.9163e-06, 8.9759e-04, 7.1074e-04, -5.9163e-06, 3.0747, -0.0158, -0.0152,
     -3.6665e-01, -8.8455e-03, 1.6113e-01, -3.6665e-01, -8.8455e-03, 1.6113e-01, -3.9716e-01, -4.0229e-02, -1.2637e-01,
     7.9163e-01, 6.8519e-02, -1.5091e-01, 7.9163e-01, 6.8519e-02, -1.5091e-01, 7.8632e-01, -4.3810e-02, 1.4375e-02,
     -1.0675e-01, 1.2635e-01, 1.6711e-02, -1.0675e-01, 1.2635e-01, 1.6711e-02])

# column blocks of the 265-vector: kept (generated) and replaced (lower body) ranges, shared by the three helpers below
_KEEP_FULL = ((0, 3), (18, 21), (27, 30), (36, 39), (45, 265))
_KEEP_PRED = ((0, 3), (3, 6), (6, 9), (9, 12), (12, 232))
_LOWER_FULL = ((3, 18), (21, 27), (30, 36), (39, 45))
_LOWER_TABLE = ((0, 15), (15, 21), (21, 27), (27, 33))


def _interleave(x, keep, fill):
    """keep[0] | fill[0] | keep[1] | fill[1] | ... | keep[4] along dim 1."""
    parts = []
    for i, (a, b) in enumerate(keep):
        parts.append(x[:, a:b])
        if i < len(fill):
            parts.append(fill[i])
    return torch.cat(parts, dim=1)


def pred2poses(pred, gt):
    """[F,232] prediction + the lower body of ground truth frame 0 -> [F,265] (lower_body.py:90-101)."""
    fill = [gt[0:1, a:b].repeat(pred.shape[0], 1) for a, b in _LOWER_FULL]
    return _interleave(pred, _KEEP_PRED, fill)


def poses2poses(poses, gt):
    """[F,265] with its lower body replaced by ground truth frame 0's (lower_body.py:104-115)."""
    fill = [gt[0:1, a:b].repeat(poses.shape[0], 1) for a, b in _LOWER_FULL]
    return _interleave(poses, _KEEP_FULL, fill)


def poses2pred(poses, stand=False):
    """[F,265] with its lower body replaced by the fixed seated / standing block (lower_body.py:117-134)."""
    lp = (LOWER_POSE_STAND if stand else LOWER_POSE).to(poses)[None].repeat(poses.shape[0], 1)
    return _interleave(poses, _KEEP_FULL, [lp[:, a:b] for a, b in _LOWER_TABLE])
