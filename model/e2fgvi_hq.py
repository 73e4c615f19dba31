from e2fgvi_b200.model.e2fgvi_hq import *  # noqa: F401,F403
from e2fgvi_b200.model.e2fgvi_hq import InpaintGenerator  # noqa: F401
