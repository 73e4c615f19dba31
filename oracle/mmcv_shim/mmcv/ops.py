"""Shim of the two mmcv.ops names the reference imports (model/modules/feat_prop.py:7).
``modulated_deform_conv2d`` is served by torchvision's DCNv2 CPU kernel (same offset/mask channel layout)."""
import math

import torch
import torch.nn as nn
from torchvision.ops import deform_conv2d as _tv_deform_conv2d


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def modulated_deform_conv2d(x, offset, mask, weight, bias, stride, padding, dilation, groups, deform_groups):
    del groups, deform_groups  # inferred by torchvision from the tensor shapes
    return _tv_deform_conv2d(x, offset, weight, bias, _pair(stride), _pair(padding), _pair(dilation), mask=mask)


class ModulatedDeformConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        n = in_channels * self.kernel_size[0] * self.kernel_size[1]
        self.weight.data.uniform_(-1.0 / math.sqrt(n), 1.0 / math.sqrt(n))
