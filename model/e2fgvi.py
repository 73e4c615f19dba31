from e2fgvi_b200.model.e2fgvi import *  # noqa: F401,F403
from e2fgvi_b200.model.e2fgvi import InpaintGenerator  # noqa: F401
