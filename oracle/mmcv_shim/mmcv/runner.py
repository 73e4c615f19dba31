"""Shim of mmcv.runner.load_checkpoint: the reference's SPyNet ctor downloads weights (flow_comp.py:59-72);
offline this is a no-op (E2FGVI checkpoints carry update_spynet.* themselves)."""


def load_checkpoint(model, filename, strict=True, **kwargs):
    return {}
