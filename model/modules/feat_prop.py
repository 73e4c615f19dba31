from e2fgvi_b200.model.modules.feat_prop import *  # noqa: F401,F403
