from e2fgvi_b200.model.modules.tfocal_transformer_hq import *  # noqa: F401,F403
