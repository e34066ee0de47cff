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
