"""TEST helper: route the three kernel entry points of ``e2fgvi_b200.ops`` to the CPU oracle so the model's HOST
logic (orchestration, state-dict plumbing, layouts) can be checked on a CPU-only box.  Never used by the product."""
import contextlib

import torch

from e2fgvi_b200 import ops
from e2fgvi_b200.model.modules.tfocal_transformer import rolled_valid_indices
from oracle import restate


def _flow_warp(x, flow, interpolation="bilinear", padding_mode="zeros", align_corners=True):
    return restate.flow_warp(x.contiguous(), flow, interpolation, padding_mode, align_corners)


def _pack(weight, deform_groups):
    return weight.detach()  # the oracle consumes the unpacked weight


def _fused(x, head, flow_1, flow_2, w_packed, bias, deform_groups, max_residue_magnitude=10.0,
           out_dtype=torch.float32, out_split=False):
    y = _fused_plain(x, head, flow_1, flow_2, w_packed, bias, deform_groups, max_residue_magnitude, out_dtype)
    return (y, _as_split(y)) if out_split else y


def _fused_plain(x, head, flow_1, flow_2, w_packed, bias, deform_groups, max_residue_magnitude=10.0,
                 out_dtype=torch.float32):
    o1, o2, mask = torch.chunk(head, 3, dim=1)
    off = max_residue_magnitude * torch.tanh(torch.cat((o1, o2), 1))
    a, b = torch.chunk(off, 2, dim=1)
    half = a.size(1) // 2
    off = torch.cat([a + flow_1.flip(1).repeat(1, half, 1, 1), b + flow_2.flip(1).repeat(1, half, 1, 1)], 1)
    return restate.modulated_deform_conv2d(x.contiguous(), off, torch.sigmoid(mask), w_packed, bias, 1, 1, 1, 1,
                                           deform_groups).to(out_dtype)


def _mdcn(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deform_groups=1,
          out_dtype=torch.float32):
    return restate.modulated_deform_conv2d(x, offset, mask, weight, bias, 1, 1, 1, 1, deform_groups).to(out_dtype)


def _attention(qkv, qkv_pooled, num_heads, window_size, expand_size, focal_window, scale, out_dtype=torch.float32):
    return restate.focal_window_attention(qkv.float(), None if qkv_pooled is None else qkv_pooled.float(), num_heads,
                                          window_size, expand_size, focal_window, scale,
                                          rolled_valid_indices(window_size, expand_size)).to(
                                              torch.float32 if isinstance(out_dtype, str) else out_dtype)


def _unfold(img, kernel_size, stride, padding, gelu=False, out="f32"):
    t = torch.nn.functional.unfold(img, kernel_size, padding=padding, stride=stride).permute(0, 2, 1).contiguous()
    return torch.nn.functional.gelu(t) if gelu else t


def _fold(tokens, output_size, kernel_size, stride, padding, normalize=False, bias=None, residual=None,
          channels_last=False):
    F = torch.nn.functional
    img = F.fold(tokens.permute(0, 2, 1), output_size, kernel_size, padding=padding, stride=stride)
    if normalize:
        ones = torch.ones(1, kernel_size[0] * kernel_size[1], tokens.shape[1], dtype=tokens.dtype)
        img = img / F.fold(ones, output_size, kernel_size, padding=padding, stride=stride)
    img = img if bias is None else img + bias[None]
    return img if residual is None else img + residual


def _fold_unfold(tokens, output_size, kernel_size, stride, padding, gelu=False, out="f32", pitch=None):
    return _unfold(_fold(tokens, output_size, kernel_size, stride, padding, normalize=True), kernel_size, stride, padding,
                   gelu=gelu)


def _dcn_pack_input(a, b):
    return torch.cat([a, b], 1)


def _prop_prologue(prop, feat_n2, flow_n1, flow_prev):
    """feat_prop.py:106-126 operator by operator (the sequence ops.prop_prologue fuses)."""
    grid_n1 = flow_n1.permute(0, 2, 3, 1)
    cond_n1 = _flow_warp(prop, grid_n1)
    if feat_n2 is not None:
        flow_n2 = flow_n1 + _flow_warp(flow_prev, grid_n1)
        cond_n2 = _flow_warp(feat_n2, flow_n2.permute(0, 2, 3, 1))
    else:
        feat_n2, flow_n2, cond_n2 = torch.zeros_like(prop), torch.zeros_like(flow_n1), torch.zeros_like(cond_n1)
    return torch.cat([prop, feat_n2], 1), cond_n1, cond_n2, torch.cat([flow_n1, flow_n2], 1), flow_n1, flow_n2


def _upsample(x):
    return torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def _layer_norm(x, weight, bias, eps=1e-5, out="f32"):
    y = torch.nn.functional.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    return (y, y) if out == "both" else y


def _window_pool(x, weight, bias, window_size, out="split"):
    """Reference formulation (tfocal_transformer.py:508-516), returned in the reference layout (B,nWh,nWw,T,C)."""
    B, T, H, W, C = x.shape
    wh, ww = window_size
    xw = x.view(B, T, H // wh, wh, W // ww, ww, C).permute(0, 2, 4, 1, 3, 5, 6).reshape(B, H // wh, W // ww, T, wh * ww, C)
    return torch.nn.functional.linear(xw.transpose(4, 5), weight, bias).flatten(-2)


def _layer_norm_pool(x, weight, bias, eps, pool_weight, pool_bias, window_size):
    B, T, H, W, C = x.shape
    y = torch.nn.functional.layer_norm(x, (C,), weight, bias, eps)
    pooled = _window_pool(y, pool_weight, pool_bias, window_size).permute(0, 3, 1, 2, 4)      # (B,T,nWh,nWw,C)
    return torch.cat([y.reshape(-1, C), pooled.reshape(-1, C)]), B * T * H * W


def _pack_rows(x, lead, cin=None):
    return x


def _dense(s):
    """fp32 (n,c,h,w) of a conv source: a tensor, or a SplitNHWC whose hi / lo are (n,h,w,c) tensors (any dtype)."""
    if isinstance(s, ops.SplitNHWC):
        return (s.hi.float() + s.lo.float()).permute(0, 3, 1, 2)
    return s


def _as_split(y):
    """(n,c,h,w) fp32 -> SplitNHWC stand-in: hi = the fp32 NHWC values, lo = zeros."""
    nhwc = y.permute(0, 2, 3, 1).contiguous()
    return ops.SplitNHWC(nhwc, torch.zeros_like(nhwc), tuple(y.shape))


def _split_bf16(x):
    return x.clone(), torch.zeros_like(x)


def _conv_frames(sources, weight, bias=None, negative_slope=1.0, residual=None, out="f32", into=None):
    F = torch.nn.functional
    srcs = [_dense(s) for s in (sources if isinstance(sources, (list, tuple)) else [sources])]
    y = F.leaky_relu(F.conv2d(torch.cat(srcs, 1), weight, bias, 1, weight.shape[2] // 2), negative_slope)
    y = y if residual is None else y + residual
    if into is not None:
        o32, ohi, olo = into
        nhwc = y.permute(0, 2, 3, 1)
        for t_ in (o32, ohi):
            if t_ is not None:
                t_.copy_(nhwc)
        if olo is not None:
            olo.zero_()
    sp = _as_split(y)
    return y if out == "f32" else sp if out == "split" else (y, sp)


def _conv3x3(sources, weight, bias=None, groups=1, negative_slope=1.0, residual=None, out="f32", stride=1,
             padding=None, out_lead=0):
    F = torch.nn.functional
    srcs = [_dense(s_) for s_ in (sources if isinstance(sources, (list, tuple)) else [sources])]
    if groups == 1:
        x = torch.cat(srcs, 1)
    else:  # group-wise concatenation (e2fgvi.py:103-108)
        n, _, h, w = srcs[0].shape
        x = torch.cat([s.reshape(n, groups, -1, h, w) for s in srcs], 2).reshape(n, -1, h, w)
    pad = weight.shape[2] // 2 if padding is None else padding
    y = F.leaky_relu(F.conv2d(x, weight, bias, stride, pad, 1, groups), negative_slope)
    y = y if residual is None else y + residual
    return (y, _as_split(y)) if out == "both" else y


def _conv_kxn(x, weight, bias=None, negative_slope=1.0, residual=None, out="f32", tanh_nchw=False, groups=1):
    y = _conv3x3(x if isinstance(x, (list, tuple)) else [x], weight, bias, groups=groups, negative_slope=negative_slope,
                 residual=residual, out="f32")
    if tanh_nchw:
        return torch.tanh(y).contiguous()
    return y if out == "f32" else _as_split(y) if out == "split" else (y, _as_split(y))


def _conv3x3_tanh_nchw(x, weight, bias):
    return torch.tanh(torch.nn.functional.conv2d(_dense(x), weight, bias, 1, 1)).contiguous()


def _split_nhwc(x):
    return x


def _linear(x, weight, bias=None, residual=None, out_dtype=torch.float32, tile_hint=0):
    y = torch.nn.functional.linear(x, weight.reshape(weight.shape[0], -1), bias)
    return (y if residual is None else y + residual.reshape(y.shape)).to(out_dtype if x.is_cuda else torch.float32)


def _soft_split(x, weight, bias, kernel_size, stride, padding):
    F = torch.nn.functional
    x = _dense(x)
    return F.linear(F.unfold(x, kernel_size, padding=padding, stride=stride).permute(0, 2, 1), weight, bias)


def _soft_comp(tokens, weight, bias, output_size, kernel_size, stride, padding, bias_map_extra=None, residual=None,
               out="f32"):
    F = torch.nn.functional
    n = tokens.shape[0]
    feat = F.linear(tokens.reshape(n, -1, tokens.shape[-1]), weight, bias)
    y = F.fold(feat.permute(0, 2, 1), output_size, kernel_size, padding=padding, stride=stride)
    if bias_map_extra is not None:
        y = y + bias_map_extra
    if residual is not None:
        y = y + residual
    return (y, y) if out == "both" else y


@contextlib.contextmanager
def oracle_ops():
    saved = {n: getattr(ops, n) for n in ("flow_warp", "pack_dcn_weight", "deform_align_fused",
                                          "modulated_deform_conv2d", "focal_window_attention", "t2t_unfold",
                                          "t2t_fold", "linear", "conv3x3", "split_nhwc", "upsample2x_split",
                                          "layer_norm", "dcn_pack_input", "t2t_fold_unfold", "pack_rows", "window_pool",
                                          "prop_prologue", "soft_split", "soft_comp", "layer_norm_pool", "conv_frames", "split_bf16", "conv3x3_tanh_nchw", "conv_kxn")}
    ops.flow_warp, ops.pack_dcn_weight, ops.deform_align_fused = _flow_warp, _pack, _fused
    ops.modulated_deform_conv2d, ops.focal_window_attention = _mdcn, _attention
    ops.t2t_unfold, ops.t2t_fold, ops.linear, ops.t2t_fold_unfold = _unfold, _fold, _linear, _fold_unfold
    ops.conv3x3, ops.split_nhwc, ops.pack_rows, ops.window_pool = _conv3x3, _split_nhwc, _pack_rows, _window_pool
    ops.upsample2x_split, ops.layer_norm, ops.dcn_pack_input = _upsample, _layer_norm, _dcn_pack_input
    ops.prop_prologue = _prop_prologue
    ops.soft_split, ops.soft_comp, ops.layer_norm_pool = _soft_split, _soft_comp, _layer_norm_pool
    ops.conv_frames, ops.split_bf16, ops.conv3x3_tanh_nchw = _conv_frames, _split_bf16, _conv3x3_tanh_nchw
    ops.conv_kxn = _conv_kxn
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
