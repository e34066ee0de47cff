from .lower_body import c_index_3d, c_index_6d, part2full, poses2poses, poses2pred, pred2poses  # noqa: F401
from .utils import get_mfcc_sepa, get_mfcc_ta  # noqa: F401
