"""HQ flavour of the temporal focal transformer (reference: model/modules/tfocal_transformer_hq.py).
The implementation is shared with the base variant; these wrappers only pin ``hq=True``."""
from functools import partial

from .tfocal_transformer import (FusionFeedForward, SoftSplit, WindowAttention, window_partition,  # noqa: F401
                                 window_partition_noreshape, window_reverse)
from .tfocal_transformer import SoftComp as _SoftComp
from .tfocal_transformer import TemporalFocalTransformerBlock as _Block


def SoftComp(channel, hidden, kernel_size, stride, padding):
    return _SoftComp(channel, hidden, None, kernel_size, stride, padding, hq=True)


TemporalFocalTransformerBlock = partial(_Block, hq=True)
