from e2fgvi_b200.model.modules.flow_comp import *  # noqa: F401,F403
