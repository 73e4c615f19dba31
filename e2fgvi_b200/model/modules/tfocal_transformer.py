"""Temporal focal transformer (reference: model/modules/tfocal_transformer.py:19-536 and _hq.py).

One implementation serves both variants: the base model fixes ``output_size`` at construction
(tfocal_transformer.py:30-37,56-59,83-87), the HQ model threads it through at run time (_hq.py:32-46,92-119).
Parameter / buffer names follow the reference so released ``state_dict``s load strictly.

The attention core (window partition, 4 rolled ring key sets, pooled-window keys with the -100 mask, softmax,
P·V, window reverse) is ONE kernel, ``ops.focal_window_attention``; rolled K/V copies and the logits matrix are
never materialised.
"""
import math
from functools import reduce

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops


def _token_grid(output_size, kernel_size, stride, padding):
    return tuple((output_size[i] + 2 * padding[i] - (kernel_size[i] - 1) - 1) // stride[i] + 1 for i in range(2))


class SoftSplit(nn.Module):
    """unfold(7x7, s3, p3) + Linear(49*C -> hidden) (tfocal_transformer.py:19-46)."""

    def __init__(self, channel, hidden, kernel_size, stride, padding, t2t_param=None):
        super().__init__()
        self.kernel_size, self.stride, self.padding = tuple(kernel_size), tuple(stride), tuple(padding)
        self.embedding = nn.Linear(reduce(lambda a, b: a * b, kernel_size) * channel, hidden)
        self.t2t_param = t2t_param
        self.output_size = None if t2t_param is None else t2t_param.get("output_size")

    def forward(self, x, b, output_size=None):
        output_size = output_size or self.output_size
        f_h, f_w = _token_grid(output_size, self.kernel_size, self.stride, self.padding)
        # unfold + Linear == a 7x7 / stride-3 conv: one implicit-GEMM launch, the 49x unfolded operand never exists
        feat = ops.soft_split(x, self.embedding.weight, self.embedding.bias, self.kernel_size, self.stride, self.padding)
        return feat.view(b, -1, f_h, f_w, feat.size(2))


class SoftComp(nn.Module):
    """Linear(hidden -> 49*C) + fold + bias map (base, tfocal_transformer.py:49-72) or 3x3 conv (HQ, _hq.py:49-79)."""

    def __init__(self, channel, hidden, output_size=None, kernel_size=(7, 7), stride=(3, 3), padding=(3, 3), hq=False):
        super().__init__()
        self.kernel_size, self.stride, self.padding = tuple(kernel_size), tuple(stride), tuple(padding)
        self.embedding = nn.Linear(hidden, reduce(lambda a, b: a * b, kernel_size) * channel)
        self.output_size = output_size
        self.hq = hq
        if hq:
            self.bias_conv = nn.Conv2d(channel, channel, kernel_size=3, stride=1, padding=1)
        else:
            self.bias = nn.Parameter(torch.zeros((channel, output_size[0], output_size[1]), dtype=torch.float32))

    def forward(self, x, t, output_size=None, residual=None):
        """``residual`` (b*t, C, H, W), optional: added to the result by the fold kernel (base model) or the conv
        epilogue (HQ) — the ``enc_feat + trans_feat`` of e2fgvi.py:263; the result is then channels_last."""
        output_size = output_size or self.output_size
        b_, t_, f_h, f_w, c_ = x.shape
        # Linear + fold == the transposed 7x7 / stride-3 conv: one implicit-GEMM launch over nine output phases; the
        # 6272-wide token matrix and the fold pass never exist
        tokens = x.view(b_ * t_, f_h, f_w, c_)
        if self.hq:
            feat = ops.soft_comp(tokens, self.embedding.weight, self.embedding.bias, output_size, self.kernel_size,
                                 self.stride, self.padding, out="split")
            return ops.conv3x3([feat], self.bias_conv.weight, self.bias_conv.bias, residual=residual)
        return ops.soft_comp(tokens, self.embedding.weight, self.embedding.bias, output_size, self.kernel_size,
                             self.stride, self.padding, bias_map_extra=self.bias, residual=residual)


class FusionFeedForward(nn.Module):
    """Linear(512->1960), overlap-average through fold/normalise/unfold, GELU, Linear(1960->512)
    (tfocal_transformer.py:75-98; _hq.py:82-119)."""

    def __init__(self, d_model, n_vecs=None, t2t_params=None):
        super().__init__()
        hd = 1960
        self.conv1 = nn.Sequential(nn.Linear(d_model, hd))
        self.conv2 = nn.Sequential(nn.GELU(), nn.Linear(hd, d_model))
        assert t2t_params is not None
        self.t2t_params = dict(t2t_params)
        self.n_vecs = n_vecs

    def forward(self, x, output_size=None, residual=None):
        """conv1 -> [fold / fold(ones) -> unfold -> GELU] (one fused kernel on the token-major layout) -> conv2
        (+ residual, fused into the GEMM epilogue)."""
        p = self.t2t_params
        output_size = output_size or p.get("output_size")
        f_h, f_w = _token_grid(output_size, p["kernel_size"], p["stride"], p["padding"])
        n_vecs = f_h * f_w
        x = ops.linear(x, self.conv1[0].weight, self.conv1[0].bias)
        b, n, c = x.size()
        # fold / fold(ones) -> unfold -> conv2[0] (GELU): one kernel, the folded image stays in shared memory
        # rows padded 1960 -> 1984 (zero columns) so that the A operand of conv2 starts every row on a 128-byte line
        x = ops.t2t_fold_unfold(x.view(-1, n_vecs, c), output_size, p["kernel_size"], p["stride"], p["padding"],
                                gelu=True, out="split", pitch=(c + 63) // 64 * 64)
        x = x.view(b, n, x.shape[-1])
        return ops.linear(x, self.conv2[1].weight, self.conv2[1].bias, residual=residual)


def window_partition(x, window_size):
    """(B,T,H,W,C) -> (B*nW, T*wh*ww, C) (tfocal_transformer.py:101-114)."""
    B, T, H, W, C = x.shape
    wh, ww = window_size
    x = x.view(B, T, H // wh, wh, W // ww, ww, C)
    return x.permute(0, 2, 4, 1, 3, 5, 6).reshape(-1, T * wh * ww, C)


def window_partition_noreshape(x, window_size):
    """(B,T,H,W,C) -> (B, nWh, nWw, T, wh, ww, C) (tfocal_transformer.py:117-129)."""
    B, T, H, W, C = x.shape
    wh, ww = window_size
    return x.view(B, T, H // wh, wh, W // ww, ww, C).permute(0, 2, 4, 1, 3, 5, 6).contiguous()


def window_reverse(windows, window_size, T, H, W):
    """(B*nW, T, wh, ww, C) -> (B,T,H,W,C) (tfocal_transformer.py:132-147)."""
    wh, ww = window_size
    B = windows.shape[0] // ((H // wh) * (W // ww))
    x = windows.view(B, H // wh, W // ww, T, wh, ww, -1)
    return x.permute(0, 3, 1, 4, 2, 5, 6).reshape(B, T, H, W, -1)


def rolled_valid_indices(window_size, expand_size):
    """Flat indices kept from the 4 rolled window copies (tfocal_transformer.py:166-179): positions of the
    (tl, tr, bl, br) rolled windows that fall OUTSIDE the query window."""
    wh, ww = window_size
    eh, ew = expand_size
    keep = []
    for quad, (row_from_end, col_from_end) in enumerate(((True, True), (True, False), (False, True), (False, False))):
        for r in range(wh):
            for c in range(ww):
                row_ok = r >= wh - eh if row_from_end else r < eh
                col_ok = c >= ww - ew if col_from_end else c < ew
                if row_ok or col_ok:
                    keep.append(quad * wh * ww + r * ww + c)
    return torch.tensor(keep, dtype=torch.int64)


class WindowAttention(nn.Module):
    """Temporal focal window attention (tfocal_transformer.py:150-399)."""

    def __init__(self, dim, expand_size, window_size, focal_window, focal_level, num_heads, qkv_bias, pool_method):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.expand_size, self.window_size = tuple(expand_size), tuple(window_size)
        self.focal_window, self.focal_level, self.pool_method = tuple(focal_window), focal_level, pool_method
        self.scale = (dim // num_heads) ** -0.5
        if focal_level > 2:
            raise NotImplementedError("E2FGVI uses focal_level=2 (one pooled level)")
        if any(i > 0 for i in self.expand_size) and focal_level > 0:
            self.register_buffer("valid_ind_rolled", rolled_valid_indices(self.window_size, self.expand_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    @property
    def uses_pooled(self):
        return self.pool_method != "none" and self.focal_level > 1

    def pooled_kernel(self):
        """Neighbourhood of pooled windows each query window attends (unfold kernel, tfocal_transformer.py:186-196)."""
        return tuple(2 * (i // 2) + 1 for i in self.focal_window)

    def attend(self, x, pooled, residual=None, joint_shape=None):
        """x (B,T,H,W,C) normed tokens (tensor or ops.SplitMat), pooled (B,nWh,nWw,T,C) tensor or the SplitMat of
        ops.window_pool (already (B,T,nWh,nWw,C)) -> (B,T,H,W,C) after proj (+ residual).
        ``joint_shape`` = (B,T,H,W): x is the SplitMat of ``ops.layer_norm_pool`` (token rows followed by pooled rows) and
        ONE qkv GEMM serves both."""
        if joint_shape is not None:
            B, T, H, W = joint_shape
            wh, ww = self.window_size
            n_tok = B * T * H * W
            both = ops.linear(x, self.qkv.weight, self.qkv.bias, out_dtype=torch.float16)
            qkv = both[:n_tok].view(B, T, H, W, -1)
            qkv_pooled = both[n_tok:].view(B, T, H // wh, W // ww, -1)
        else:
            qkv = ops.linear(x, self.qkv.weight, self.qkv.bias, out_dtype=torch.float16)
            qkv_pooled = None
            if self.uses_pooled:
                if not isinstance(pooled, ops.SplitMat):      # reference layout (B,nWh,nWw,T,C) -> (B,T,nWh,nWw,C)
                    pooled = pooled.permute(0, 3, 1, 2, 4).contiguous()
                qkv_pooled = ops.linear(pooled, self.qkv.weight, self.qkv.bias, out_dtype=torch.float16)
        out = ops.focal_window_attention(qkv, qkv_pooled, self.num_heads, self.window_size, self.expand_size,
                                         self.pooled_kernel(), self.scale, out_dtype="split")
        return ops.linear(out, self.proj.weight, self.proj.bias, residual=residual)

    def forward(self, x_all, mask_all=None):
        """Reference boundary: x_all = [x (B,T,H,W,C), pooled (B,nWh,nWw,T,C)] -> (B*nW, T*wh*ww, C)."""
        del mask_all  # always [None, None] on the path (tfocal_transformer.py:475)
        out = self.attend(x_all[0], x_all[1] if len(x_all) > 1 else None)
        return window_partition(out, self.window_size)


class TemporalFocalTransformerBlock(nn.Module):
    """LN -> window pool -> focal attention -> +res -> LN -> fusion FFN -> +res (tfocal_transformer.py:402-536)."""

    def __init__(self, dim, num_heads, window_size=(5, 9), mlp_ratio=4., qkv_bias=True, pool_method="fc",
                 focal_level=2, focal_window=(5, 9), norm_layer=nn.LayerNorm, n_vecs=None, t2t_params=None,
                 hq=False):
        super().__init__()
        self.dim, self.num_heads, self.window_size = dim, num_heads, tuple(window_size)
        self.expand_size = tuple(i // 2 for i in window_size)
        self.mlp_ratio, self.pool_method = mlp_ratio, pool_method
        self.focal_level, self.focal_window = focal_level, tuple(focal_window)
        self.hq = hq
        self.pool_layers = nn.ModuleList()
        if pool_method != "none":
            for k in range(focal_level - 1):
                ws = tuple(math.floor(i / (2 ** k)) for i in self.window_size)
                layer = nn.Linear(ws[0] * ws[1], 1)
                layer.weight.data.fill_(1.0 / (ws[0] * ws[1]))
                layer.bias.data.fill_(0)
                self.pool_layers.append(layer)
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, self.expand_size, self.window_size, focal_window, focal_level, num_heads,
                                    qkv_bias, pool_method)
        self.norm2 = norm_layer(dim)
        self.mlp = FusionFeedForward(dim, n_vecs=n_vecs, t2t_params=t2t_params)

    def _forward(self, x, output_size):
        shortcut = x
        B, T, H, W, C = x.shape
        if self.attn.uses_pooled:
            if H % self.window_size[0] or W % self.window_size[1]:
                raise ValueError(f"token grid {H}x{W} must be a multiple of the window {self.window_size}")
            # norm1 + focal window pooling in one kernel (pooled tokens accumulated in registers next to the LayerNorm);
            # token rows and pooled rows share one operand buffer, so one qkv GEMM serves both
            lin = self.pool_layers[0]
            rows, _ = ops.layer_norm_pool(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, lin.weight, lin.bias,
                                          self.window_size)
            x = self.attn.attend(rows, None, residual=shortcut, joint_shape=(B, T, H, W))
        else:
            xn_split = ops.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, out="split")
            x = self.attn.attend(xn_split, None, residual=shortcut)
        y = ops.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, out="split")
        return self.mlp(y.view(B, T * H * W, C), output_size, residual=x.view(B, T * H * W, C)).view(B, T, H, W, C)

    def forward(self, x):
        if self.hq:  # x = [tokens, (h, w)] -> (tokens, (h, w))   (_hq.py:492-495,562-565)
            tokens, output_size = x[0], x[1]
            return self._forward(tokens, output_size), output_size
        return self._forward(x, None)
