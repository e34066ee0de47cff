from .lower_body import c_index_3d, part2full  # noqa: F401
from .utils import get_mfcc_ta  # noqa: F401
