"""Operator-level mirror of the three reference boundaries the CUDA kernels sit behind:

* ``flow_warp``                   — model/modules/flow_comp.py:345-383
* ``modulated_deform_conv2d``     — mmcv.ops.modulated_deform_conv2d as called at model/modules/feat_prop.py:55-58
* ``deform_align_fused``          — the tail of SecondOrderDeformableAlignment.forward, feat_prop.py:41-58
* ``focal_window_attention``      — WindowAttention.forward between qkv and proj, tfocal_transformer.py:226-396

Same names / argument meaning / error behaviour as the reference operators; every call goes through the C ABI of
``libe2fgvi_b200.so`` on the current CUDA stream.  There is no CPU path: CPU tensors raise.
"""
import os
import weakref

import torch

from . import _lib

_PAD = {"zeros": 0, "border": 1}
_DT = {torch.float32: 0, torch.float16: 1}


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional live kernel timing for bench.py's roofline block: name -> [(start_event, end_event, work)], recorded on
# the launching stream around the C-ABI call (the call is one kernel launch).
_PROFILE = None


def profile_kernels(enable=True):
    """Start (returns the live dict) or stop (enable=False) recording CUDA events around kernel launches."""
    global _PROFILE
    _PROFILE = {} if enable else None
    return _PROFILE


class _timed:
    def __init__(self, name, work=0.0):
        self.name, self.work = name, work

    def __enter__(self):
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.e1.record()
            _PROFILE.setdefault(self.name, []).append((self.e0, self.e1, self.work))
        return False


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("e2fgvi_b200 kernels need CUDA tensors on a B200 (sm_100a); there is no CPU fallback")


def _is_cl(x):
    """True if the 4-D tensor is dense NHWC in memory (channels_last or a permuted NHWC view)."""
    n, c, h, w = x.shape
    return x.stride() == (h * w * c, 1, w * c, c)


def flow_warp(x, flow, interpolation="bilinear", padding_mode="zeros", align_corners=True):
    """Warp ``x`` (n,c,h,w) with ``flow`` (n,h,w,2; pixel units, [...,0] = x-displacement).

    Mirrors flow_comp.py:345-383, including the ValueError on a spatial mismatch (:364-366).  Output keeps the
    memory format of ``x`` (NCHW-contiguous or channels_last) and its dtype (fp32 / fp16).
    """
    if x.size()[-2:] != flow.size()[1:3]:
        raise ValueError(f"The spatial sizes of input ({x.size()[-2:]}) and "
                         f"flow ({flow.size()[1:3]}) are not the same.")
    if interpolation != "bilinear" or not align_corners:
        raise NotImplementedError("only bilinear / align_corners=True is on the E2FGVI path")
    if padding_mode not in _PAD:
        raise NotImplementedError(f"padding_mode={padding_mode!r} is not on the E2FGVI path")
    _need_cuda(x, flow)
    lib = _lib.load()
    n, c, h, w = x.shape
    flow = flow.contiguous().float()
    vec = 8 if x.dtype == torch.float16 else 4
    if x.dtype in _DT and _is_cl(x) and c % vec == 0:
        out = torch.empty_like(x)  # preserves strides (dense)
        st = lib.e2f_flow_warp(x.data_ptr(), flow.data_ptr(), out.data_ptr(), n, h, w, c, _DT[x.dtype],
                               _PAD[padding_mode], _stream())
        _lib.check(st, "e2f_flow_warp")
        return out
    xc = x.contiguous()
    if xc.dtype != torch.float32:
        xc = xc.float()
    out = torch.empty_like(xc)
    st = lib.e2f_flow_warp_nchw(xc.data_ptr(), flow.data_ptr(), out.data_ptr(), n, c, h, w, _PAD[padding_mode],
                                _stream())
    _lib.check(st, "e2f_flow_warp_nchw")
    return out if out.dtype == x.dtype else out.to(x.dtype)


def pack_dcn_weight(weight, deform_groups):
    """fp32 [Cout,Cin,3,3] -> fp16 [Cout, 9*Cin] GEMM operand in sampler K-order (k = (g*9+tap)*cpg + c).

    Not cached here (a data_ptr-keyed cache is unsound once the allocator reuses addresses); modules that own a
    long-lived weight cache the result themselves (see SecondOrderDeformableAlignment.packed_weight)."""
    _need_cuda(weight)
    cout, cin, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise ValueError("only 3x3 deformable kernels are on the E2FGVI path")
    w32 = weight.detach().contiguous().float()
    packed = torch.empty(cout, 9 * cin, dtype=torch.float16, device=weight.device)
    st = _lib.load().e2f_dcn_pack_weight(w32.data_ptr(), packed.data_ptr(), cout, cin, deform_groups, _stream())
    _lib.check(st, "e2f_dcn_pack_weight")
    return packed


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def modulated_deform_conv2d(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                            deform_groups=1, out_dtype=torch.float32):
    """Drop-in for ``mmcv.ops.modulated_deform_conv2d`` (call site feat_prop.py:55-58).

    x (n,cin,h,w), offset (n,2*9*dg,h,w) with channel (g*9+tap)*2+{dy,dx}, mask (n,9*dg,h,w), weight
    (cout,cin,3,3).  Returns (n,cout,h,w) in channels_last memory format.
    """
    if _pair(stride) != (1, 1) or _pair(padding) != (1, 1) or _pair(dilation) != (1, 1) or groups != 1:
        raise NotImplementedError("E2FGVI uses 3x3 / stride 1 / padding 1 / dilation 1 / groups 1 only")
    _need_cuda(x, offset, mask, weight, bias)
    n, cin, h, w = x.shape
    cout = weight.shape[0]
    if offset.shape != (n, 18 * deform_groups, h, w) or mask.shape != (n, 9 * deform_groups, h, w):
        raise ValueError(f"offset/mask shapes {tuple(offset.shape)}/{tuple(mask.shape)} do not match x {tuple(x.shape)}")
    wp = pack_dcn_weight(weight, deform_groups)
    x_cl = x.to(dtype=torch.float16, memory_format=torch.channels_last)
    off_cl = offset.to(dtype=torch.float32, memory_format=torch.channels_last)
    msk_cl = mask.to(dtype=torch.float32, memory_format=torch.channels_last)
    b32 = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty((n, cout, h, w), dtype=out_dtype, device=x.device, memory_format=torch.channels_last)
    st = _lib.load().e2f_modulated_deform_conv2d(
        x_cl.data_ptr(), off_cl.data_ptr(), msk_cl.data_ptr(), wp.data_ptr(),
        None if b32 is None else b32.data_ptr(), out.data_ptr(), n, h, w, cin, cout, deform_groups,
        _DT[out_dtype], 0, _stream())
    _lib.check(st, "e2f_modulated_deform_conv2d")
    return out


class GroupedX:
    """DCN input in the group-major fp16 layout [N][G][H][W][16] (see ``dcn_pack_input``)."""

    __slots__ = ("data", "shape")

    def __init__(self, data, shape):
        self.data, self.shape = data, shape       # shape: logical (N, Cin, H, W)


def dcn_pack_input(a, b):
    """``torch.cat([a, b], 1)`` (feat_prop.py:126) converted to fp16 in the group-major layout the deformable sampler
    reads best (adjacent bilinear corners contiguous).  a, b: (N,C,H,W) fp32, C % 16 == 0 -> ``GroupedX``."""
    _need_cuda(a, b)
    n, ca, h, w = a.shape
    cb = b.shape[1]
    a_cl = a.permute(0, 2, 3, 1).contiguous().float()
    b_cl = b.permute(0, 2, 3, 1).contiguous().float()
    xg = torch.empty((n, (ca + cb) // 16, h, w, 16), dtype=torch.float16, device=a.device)
    st = _lib.load().e2f_dcn_pack_input(a_cl.data_ptr(), b_cl.data_ptr(), xg.data_ptr(), n, h, w, ca, cb, _stream())
    _lib.check(st, "e2f_dcn_pack_input")
    return GroupedX(xg, (n, ca + cb, h, w))


def prop_prologue(prop, feat_n2, flow_n1, flow_prev):
    """Everything one propagation step does before its offset-head conv and its DCN (feat_prop.py:106-126), fused:

        cond_n1 = flow_warp(prop, flow_n1);  flow_n2 = flow_n1 + flow_warp(flow_prev, flow_n1)
        cond_n2 = flow_warp(feat_n2, flow_n2);  cat([flow_n1, flow_n2], 1);  cat([prop, feat_n2], 1)

    prop, feat_n2: (n,c,h,w) fp32 channels_last (feat_n2 / flow_prev None on the second frame of a sweep: zeros);
    flow_n1, flow_prev: (n,2,h,w) fp32 with contiguous (h,w) planes (slices ``flows[:, i]`` of the (b,t-1,2,h,w) tensor).
    Returns ``(x, cond_n1, cond_n2, flows, flow_n1, flow_n2)``: x a ``GroupedX`` (DCN input), cond_* / flows ``SplitNHWC``
    conv operands, flow_n1 / flow_n2 (n,2,h,w) views of NHWC buffers (what ``deform_align_fused`` reads without a copy).
    Bit-identical to the unfused sequence of ``flow_warp`` / add / ``split_nhwc`` / ``dcn_pack_input``."""
    _need_cuda(prop, feat_n2, flow_n1, flow_prev)
    n, c, h, w = prop.shape
    if (feat_n2 is None) != (flow_prev is None):
        raise ValueError("prop_prologue: feat_n2 and flow_prev go together")
    if c % 16:
        raise ValueError("prop_prologue: the channel count must be a multiple of 16")

    def nhwc(t):
        return t if (t.dtype == torch.float32 and _is_cl(t)) else t.float().contiguous(memory_format=torch.channels_last)

    def planes(f):
        if f.dtype != torch.float32 or f.stride()[1:] != (h * w, w, 1):
            f = f.float().contiguous()
        return f

    prop, flow_n1 = nhwc(prop), planes(flow_n1)
    if feat_n2 is not None:
        feat_n2, flow_prev = nhwc(feat_n2), planes(flow_prev)
    dev = prop.device
    bf = torch.bfloat16
    c1h, c1l = torch.empty((n, h, w, c), dtype=bf, device=dev), torch.empty((n, h, w, c), dtype=bf, device=dev)
    c2h, c2l = torch.empty((n, h, w, c), dtype=bf, device=dev), torch.empty((n, h, w, c), dtype=bf, device=dev)
    f1 = torch.empty((n, h, w, 2), dtype=torch.float32, device=dev)
    f2 = torch.empty((n, h, w, 2), dtype=torch.float32, device=dev)
    flh, fll = torch.empty((n, h, w, 8), dtype=bf, device=dev), torch.empty((n, h, w, 8), dtype=bf, device=dev)
    xg = torch.empty((n, 2 * c // 16, h, w, 16), dtype=torch.float16, device=dev)
    st = _lib.load().e2f_prop_prologue(
        prop.data_ptr(), None if feat_n2 is None else feat_n2.data_ptr(), flow_n1.data_ptr(), flow_n1.stride(0),
        None if flow_prev is None else flow_prev.data_ptr(), 0 if flow_prev is None else flow_prev.stride(0),
        c1h.data_ptr(), c1l.data_ptr(), c2h.data_ptr(), c2l.data_ptr(), f1.data_ptr(), f2.data_ptr(), flh.data_ptr(),
        fll.data_ptr(), xg.data_ptr(), n, h, w, c, _stream())
    _lib.check(st, "e2f_prop_prologue")
    return (GroupedX(xg, (n, 2 * c, h, w)), SplitNHWC(c1h, c1l, (n, c, h, w)), SplitNHWC(c2h, c2l, (n, c, h, w)),
            SplitNHWC(flh, fll, (n, 4, h, w)), f1.permute(0, 3, 1, 2), f2.permute(0, 3, 1, 2))


def deform_align_fused(x, head, flow_1, flow_2, w_packed, bias, deform_groups, max_residue_magnitude=10.0,
                       out_dtype=torch.float32, out_split=False):
    """feat_prop.py:41-58 in one kernel: 10*tanh + flow.flip(1) add, sigmoid, deformable sampling, GEMM, bias.

    x (n,cin,h,w) fp16 channels_last; head (n,27*dg,h,w) fp32 channels_last (raw conv_offset output);
    flow_k (n,2,h,w) any layout (converted to (n,h,w,2) fp32).  Returns (n,cout,h,w) channels_last; with
    ``out_split=True`` (fp32 output only) ``(tensor, SplitNHWC)`` — the bf16 operand pair of the backbone conv that
    consumes the aligned features, written by the same epilogue.
    """
    grouped = isinstance(x, GroupedX)
    _need_cuda(None if grouped else x, head, flow_1, flow_2, w_packed, bias)
    n, cin, h, w = x.shape
    cout = w_packed.shape[0]
    if grouped:
        x = x.data
    elif x.dtype != torch.float16 or not _is_cl(x):
        x = x.to(dtype=torch.float16, memory_format=torch.channels_last)
    if head.dtype != torch.float32 or not _is_cl(head):
        head = head.to(dtype=torch.float32, memory_format=torch.channels_last)
    f1 = flow_1.permute(0, 2, 3, 1).contiguous().float()
    f2 = flow_2.permute(0, 2, 3, 1).contiguous().float()
    b32 = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty((n, cout, h, w), dtype=out_dtype, device=head.device, memory_format=torch.channels_last)
    if out_split:
        if out_dtype != torch.float32:
            raise ValueError("deform_align_fused: out_split goes with the fp32 output")
        ohi = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=head.device)
        olo = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=head.device)
        with _timed("deform_align_fused", 2.0 * cout * cin * 9 * n * h * w):
            st = _lib.load().e2f_deform_align_fused_split(
                x.data_ptr(), head.data_ptr(), f1.data_ptr(), f2.data_ptr(), w_packed.data_ptr(),
                None if b32 is None else b32.data_ptr(), out.data_ptr(), ohi.data_ptr(), olo.data_ptr(), n, h, w, cin,
                cout, deform_groups, float(max_residue_magnitude), 1 if grouped else 0, _stream())
        _lib.check(st, "e2f_deform_align_fused_split")
        return out, SplitNHWC(ohi, olo, (n, cout, h, w))
    with _timed("deform_align_fused", 2.0 * cout * cin * 9 * n * h * w):
        st = _lib.load().e2f_deform_align_fused(
            x.data_ptr(), head.data_ptr(), f1.data_ptr(), f2.data_ptr(), w_packed.data_ptr(),
            None if b32 is None else b32.data_ptr(), out.data_ptr(), n, h, w, cin, cout, deform_groups,
            float(max_residue_magnitude), _DT[out_dtype], 1 if grouped else 0, _stream())
    _lib.check(st, "e2f_deform_align_fused")
    return out


def focal_window_attention(qkv, qkv_pooled, num_heads, window_size, expand_size, focal_window, scale,
                           out_dtype=torch.float16):
    """softmax(q k_all^T) v_all of tfocal_transformer.py:226-396, un-partitioned output.

    qkv (B,T,H,W,3C) fp16, qkv_pooled (B,T,nWh,nWw,3C) fp16 or None (focal_level 1) -> (B,T,H,W,C) in
    ``out_dtype`` (torch.float32 / torch.float16), or — ``out_dtype="split"`` — a ``SplitMat``, the bf16 (hi, lo)
    operand pair of the following ``linear`` (attn.proj) written by the epilogue.
    """
    _need_cuda(qkv, qkv_pooled)
    B, T, H, W, C3 = qkv.shape
    C = C3 // 3
    wh, ww = window_size
    if H % wh or W % ww:
        raise ValueError(f"token grid {H}x{W} is not a multiple of the window {wh}x{ww}")
    if qkv.dtype != torch.float16 or not qkv.is_contiguous():
        qkv = qkv.contiguous().half()
    use_pooled = qkv_pooled is not None
    if use_pooled:
        if qkv_pooled.shape != (B, T, H // wh, W // ww, C3):
            raise ValueError(f"qkv_pooled shape {tuple(qkv_pooled.shape)} != {(B, T, H // wh, W // ww, C3)}")
        if qkv_pooled.dtype != torch.float16 or not qkv_pooled.is_contiguous():
            qkv_pooled = qkv_pooled.contiguous().half()
    split = isinstance(out_dtype, str) and out_dtype == "split"
    if split:
        out = torch.empty((2, B, T, H, W, C), dtype=torch.bfloat16, device=qkv.device)
    else:
        out = torch.empty((B, T, H, W, C), dtype=out_dtype, device=qkv.device)
    with _timed("focal_window_attention", attention_flops(B, T, H, W, C, window_size, expand_size, focal_window,
                                                            use_pooled)):
        st = _lib.load().e2f_focal_window_attention(
            qkv.data_ptr(), qkv_pooled.data_ptr() if use_pooled else None, out.data_ptr(), B, T, H, W, num_heads,
            C // num_heads, wh, ww, expand_size[0], expand_size[1], focal_window[0], focal_window[1],
            1 if use_pooled else 0, float(scale), 2 if split else _DT[out_dtype], _stream())
    _lib.check(st, "e2f_focal_window_attention")
    return SplitMat(out[0], out[1]) if split else out


class SplitMat:
    """A (..., K) activation as two bf16 terms (hi + lo): the A operand of ``linear`` produced directly by a fused
    producer (LayerNorm, unfold+GELU), so no fp32 round trip and no standalone split launch."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    def view(self, *shape):
        return SplitMat(self.hi.view(*shape), self.lo.view(*shape))


def t2t_unfold(img, kernel_size, stride, padding, gelu=False, out="f32"):
    """``F.unfold(img, k, padding=p, stride=s).permute(0, 2, 1)`` (tfocal_transformer.py:39-43, :94-96) in one
    gather kernel, optionally followed by the exact GELU.  img (BT,C,H,W) fp32 -> tokens (BT, L, C*k*k): fp32
    tensor (out="f32") or a ``SplitMat`` (out="split")."""
    _need_cuda(img)
    (k, k2), (s, s2), (p, p2) = _pair(kernel_size), _pair(stride), _pair(padding)
    if k != k2 or s != s2 or p != p2:
        raise NotImplementedError("square kernel / stride / padding only (E2FGVI uses 7 / 3 / 3)")
    bt, c, h, w = img.shape
    # channels_last storage (conv / linear epilogues write it) is read in place by the staged 7/3/3 kernel
    nhwc = ((k, s, p) == (7, 3, 3) and c % 8 == 0 and img.dtype == torch.float32 and not img.is_contiguous()
            and img.permute(0, 2, 3, 1).is_contiguous() and bt <= 65535
            and 8 * 7 * (w + 6) * 4 <= 200 * 1024)      # the staged kernel's shared-memory row buffer (t2t.cu: U2_CC rows)
    if not nhwc:
        img = img.contiguous().float()
    fh, fw = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    shape = (bt, fh * fw, c * k * k)
    tok = hi = lo = None
    if out == "f32":
        tok = torch.empty(shape, dtype=torch.float32, device=img.device)
    elif out == "split":
        hi = torch.empty(shape, dtype=torch.bfloat16, device=img.device)
        lo = torch.empty(shape, dtype=torch.bfloat16, device=img.device)
    else:
        raise ValueError("out must be 'f32' or 'split'")
    numel = shape[0] * shape[1] * shape[2]
    with _timed("t2t_unfold", float(numel * 4 + img.numel() * 4)):
        fn = _lib.load().e2f_t2t_unfold_nhwc if nhwc else _lib.load().e2f_t2t_unfold
        st = fn(img.data_ptr(), None if tok is None else tok.data_ptr(),
                                        None if hi is None else hi.data_ptr(), None if lo is None else lo.data_ptr(),
                                        bt, c, h, w, k, s, p, 1 if gelu else 0, _stream())
    _lib.check(st, "e2f_t2t_unfold")
    return tok if out == "f32" else SplitMat(hi, lo)


def t2t_fold_unfold(tokens, output_size, kernel_size, stride, padding, gelu=False, out="f32", pitch=None):
    """``unfold(fold(tokens) / fold(ones))`` (+ exact GELU): the middle of FusionFeedForward.forward
    (tfocal_transformer.py:89-96) as ONE kernel for the 7/3/3 geometry — the folded image lives in shared memory only.
    Other geometries compose ``t2t_fold(normalize=True)`` and ``t2t_unfold``.  tokens (BT, L, C*k*k) fp32 -> same
    shape, fp32 (out="f32") or ``SplitMat`` (out="split").  ``pitch`` (multiple of 4 >= C*k*k) pads every output row
    with zero columns — ``linear`` zero-pads its weight to match — so that GEMM rows start on 128-byte lines."""
    _need_cuda(tokens)
    (k, k2), (s, s2), (p, p2) = _pair(kernel_size), _pair(stride), _pair(padding)
    if k != k2 or s != s2 or p != p2:
        raise NotImplementedError("square kernel / stride / padding only (E2FGVI uses 7 / 3 / 3)")
    h, w = output_size
    tokens = tokens.contiguous().float()
    bt, n_tok, ck = tokens.shape
    c = ck // (k * k)
    fh, fw = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    if n_tok != fh * fw or c * k * k != ck:
        raise ValueError(f"t2t_fold_unfold: tokens {tuple(tokens.shape)} do not match output_size {output_size}")
    fused = (k, s, p) == (7, 3, 3) and c % 4 == 0 and bt <= 65535      # any width: the kernel tiles wide images in x
    if not fused:
        img = t2t_fold(tokens, output_size, kernel_size, stride, padding, normalize=True)
        return t2t_unfold(img, kernel_size, stride, padding, gelu=gelu, out=out)
    pitch = ck if pitch is None else int(pitch)
    if pitch < ck or pitch % 4:
        raise ValueError(f"t2t_fold_unfold: pitch {pitch} must be a multiple of 4 >= {ck}")
    oshape = (bt, n_tok, pitch)
    tok = hi = lo = None
    if out == "f32":
        tok = torch.empty(oshape, dtype=torch.float32, device=tokens.device)
    elif out == "split":
        hi = torch.empty(oshape, dtype=torch.bfloat16, device=tokens.device)
        lo = torch.empty(oshape, dtype=torch.bfloat16, device=tokens.device)
    else:
        raise ValueError("out must be 'f32' or 'split'")
    with _timed("t2t_fold_unfold", float(tokens.numel() * 8)):
        st = _lib.load().e2f_t2t_fold_unfold(tokens.data_ptr(), None if tok is None else tok.data_ptr(),
                                             None if hi is None else hi.data_ptr(),
                                             None if lo is None else lo.data_ptr(), bt, c, h, w, k, s, p,
                                             1 if gelu else 0, pitch, _stream())
    _lib.check(st, "e2f_t2t_fold_unfold")
    return tok if out == "f32" else SplitMat(hi, lo)


def window_pool(x, weight, bias, window_size, out="split"):
    """pool_layers[0] (``nn.Linear(wh*ww, 1)`` across the tokens of each window, per channel; tfocal_transformer.py:
    508-516) in one kernel.  x: the split LayerNorm output as a ``SplitMat`` of shape (B,T,H,W,C); weight (1, wh*ww);
    bias (1,).  Returns the pooled tokens ordered (B,T,nWh,nWw,C) — the reference's (B,nWh,nWw,T,C) after its
    ``permute(0,3,1,2,4)`` — as fp32 (out="f32") or ``SplitMat`` (out="split")."""
    if not isinstance(x, SplitMat):
        raise TypeError("window_pool takes the SplitMat written by layer_norm(out='split')")
    _need_cuda(x.hi, weight, bias)
    B, T, H, W, C = x.shape
    wh, ww = window_size
    if H % wh or W % ww:
        raise ValueError(f"token grid {H}x{W} must be a multiple of the window {wh}x{ww}")
    shape = (B, T, H // wh, W // ww, C)
    dev = x.hi.device
    w32 = weight.detach().float().contiguous()
    b32 = None if bias is None else bias.detach().float().contiguous()
    o32 = torch.empty(shape, dtype=torch.float32, device=dev) if out == "f32" else None
    ohi = torch.empty(shape, dtype=torch.bfloat16, device=dev) if out == "split" else None
    olo = torch.empty(shape, dtype=torch.bfloat16, device=dev) if out == "split" else None
    if out not in ("f32", "split"):
        raise ValueError("out must be 'f32' or 'split'")
    with _timed("window_pool", float(x.hi.numel() * 4)):
        st = _lib.load().e2f_window_pool(x.hi.data_ptr(), x.lo.data_ptr(), w32.data_ptr(),
                                         None if b32 is None else b32.data_ptr(),
                                         None if o32 is None else o32.data_ptr(),
                                         None if ohi is None else ohi.data_ptr(),
                                         None if olo is None else olo.data_ptr(), B * T, H, W, C, wh, ww, _stream())
    _lib.check(st, "e2f_window_pool")
    return o32 if out == "f32" else SplitMat(ohi, olo)


def upsample2x_split(x):
    """``F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)`` (deconv.forward, e2fgvi.py:125-129)
    fused with the bf16 split: (N,C,H,W) fp32 -> ``SplitNHWC`` of (N,C,2H,2W); the upsampled fp32 tensor never exists."""
    _need_cuda(x)
    n, c, h, w = x.shape
    if c % 8:
        raise ValueError("upsample2x_split needs C % 8 == 0")
    xcl = x.permute(0, 2, 3, 1).contiguous().float()      # no-op for channels_last inputs
    hi = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
    with _timed("upsample2x_split", float(xcl.numel() * 4 + hi.numel() * 4)):
        st = _lib.load().e2f_upsample2x_split(xcl.data_ptr(), hi.data_ptr(), lo.data_ptr(), n, h, w, c, _stream())
    _lib.check(st, "e2f_upsample2x_split")
    return SplitNHWC(hi, lo, (n, c, 2 * h, 2 * w))


def layer_norm(x, weight, bias, eps=1e-5, out="f32"):
    """``F.layer_norm(x, (C,), weight, bias, eps)`` over the last dim (C = 512).  out = "f32": tensor; "split": a
    ``SplitMat`` (operand of the following ``linear``); "both": (tensor, SplitMat)."""
    _need_cuda(x, weight, bias)
    c = x.shape[-1]
    xc = x.contiguous().float()
    rows = xc.numel() // c
    want_f32, want_split = out in ("f32", "both"), out in ("split", "both")
    if not (want_f32 or want_split):
        raise ValueError("out must be 'f32', 'split' or 'both'")
    o32 = torch.empty_like(xc) if want_f32 else None
    hi = torch.empty(xc.shape, dtype=torch.bfloat16, device=x.device) if want_split else None
    lo = torch.empty(xc.shape, dtype=torch.bfloat16, device=x.device) if want_split else None
    g32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
    with _timed("layernorm_split", float(xc.numel() * 4 * (1 + want_f32 + want_split))):
        st = _lib.load().e2f_layernorm_split(xc.data_ptr(), g32.data_ptr(), b32.data_ptr(),
                                             None if o32 is None else o32.data_ptr(),
                                             None if hi is None else hi.data_ptr(),
                                             None if lo is None else lo.data_ptr(), rows, c, float(eps), _stream())
    _lib.check(st, "e2f_layernorm_split")
    sp = SplitMat(hi, lo) if want_split else None
    return o32 if out == "f32" else sp if out == "split" else (o32, sp)


def layer_norm_pool(x, weight, bias, eps, pool_weight, pool_bias, window_size):
    """norm1 + pool_layers[0] of a TemporalFocalTransformerBlock in one kernel (tfocal_transformer.py:470, :508-516):
    LayerNorm every token and — from the normalised values still in registers — pool every (frame, window) with the
    ``nn.Linear(wh*ww, 1)`` weights.  x (B,T,H,W,C) fp32, C = 512.

    Returns ``(all_rows, n_tok)``: ``all_rows`` a ``SplitMat`` of shape (B*T*H*W + B*T*nWh*nWw, C) — the normalised
    tokens followed by the pooled tokens ordered (B,T,nWh,nWw) — so ONE qkv ``linear`` over it yields qkv and qkv_pooled
    back to back; ``n_tok`` = B*T*H*W."""
    _need_cuda(x, weight, bias, pool_weight, pool_bias)
    B, T, H, W, C = x.shape
    wh, ww = window_size
    if H % wh or W % ww:
        raise ValueError(f"token grid {H}x{W} must be a multiple of the window {wh}x{ww}")
    xc = x.contiguous().float()
    n_tok, n_pool = B * T * H * W, B * T * (H // wh) * (W // ww)
    hi = torch.empty((n_tok + n_pool, C), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((n_tok + n_pool, C), dtype=torch.bfloat16, device=x.device)
    g32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
    pw = pool_weight.detach().float().contiguous()
    pb = None if pool_bias is None else pool_bias.detach().float().contiguous()
    with _timed("layernorm_split", float(xc.numel() * 8 + n_pool * C * 4)):
        st = _lib.load().e2f_layernorm_pool_split(xc.data_ptr(), g32.data_ptr(), b32.data_ptr(), pw.data_ptr(),
                                                  None if pb is None else pb.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                                  B * T, H, W, C, wh, ww, float(eps), _stream())
    _lib.check(st, "e2f_layernorm_pool_split")
    return SplitMat(hi, lo), n_tok


def t2t_fold(tokens, output_size, kernel_size, stride, padding, normalize=False, bias=None, residual=None,
             channels_last=False):
    """``F.fold(tokens.permute(0, 2, 1), output_size, k, padding=p, stride=s)`` (tfocal_transformer.py:65-72,
    :89-96), optionally divided by fold(ones) and/or with a (C,H,W) bias map added.
    tokens (BT, L, C*k*k) fp32 -> img (BT, C, H, W) fp32.  ``channels_last=True`` returns the image in channels_last
    storage (what the decoder's convs read) and lets ``residual`` (BT,C,H,W) be added by the same kernel."""
    _need_cuda(tokens, bias, residual)
    (k, k2), (s, s2), (p, p2) = _pair(kernel_size), _pair(stride), _pair(padding)
    if k != k2 or s != s2 or p != p2:
        raise NotImplementedError("square kernel / stride / padding only (E2FGVI uses 7 / 3 / 3)")
    tokens = tokens.contiguous().float()
    bt, L, ck = tokens.shape
    h, w = output_size
    fh, fw = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    if L != fh * fw or ck % (k * k):
        raise ValueError(f"tokens {tuple(tokens.shape)} do not match output_size {output_size} with k={k}, s={s}, p={p}")
    c = ck // (k * k)
    if bias is not None:
        if tuple(bias.shape) != (c, h, w):
            raise ValueError(f"bias {tuple(bias.shape)} != {(c, h, w)}")
        bias = bias.detach().contiguous().float()
    if residual is not None and tuple(residual.shape) != (bt, c, h, w):
        raise ValueError(f"residual {tuple(residual.shape)} != {(bt, c, h, w)}")
    # the shared-memory fold needs a band of >= 3 image rows x 8 channels (+ two count tables) in 200 KB (t2t.cu)
    band_fits = (8 * 3 * 4 + 8) * (w + 6) <= 200 * 1024
    if channels_last and (k, s, p) == (7, 3, 3) and c % 8 == 0 and bt <= 65535 and band_fits:
        res = None if residual is None else residual.permute(0, 2, 3, 1).contiguous().float()   # no-op if channels_last
        img = torch.empty((bt, h, w, c), dtype=torch.float32, device=tokens.device)
        with _timed("t2t_fold", float(tokens.numel() * 4 + img.numel() * (8 if res is not None else 4))):
            st = _lib.load().e2f_t2t_fold_nhwc(tokens.data_ptr(), None if bias is None else bias.data_ptr(),
                                               None if res is None else res.data_ptr(), img.data_ptr(), bt, c, h, w,
                                               k, s, p, 1 if normalize else 0, _stream())
        _lib.check(st, "e2f_t2t_fold_nhwc")
        return img.permute(0, 3, 1, 2)
    img = torch.empty((bt, c, h, w), dtype=torch.float32, device=tokens.device)
    with _timed("t2t_fold", float(tokens.numel() * 4 + img.numel() * 4)):
        st = _lib.load().e2f_t2t_fold(tokens.data_ptr(), None if bias is None else bias.data_ptr(), img.data_ptr(), bt,
                                      c, h, w, k, s, p, 1 if normalize else 0, _stream())
    _lib.check(st, "e2f_t2t_fold")
    if residual is not None:
        img = img + residual
    return img.contiguous(memory_format=torch.channels_last) if channels_last else img


def split_bf16(x):
    """fp32 tensor -> (hi, lo) bf16 tensors with x ~= hi + lo to 2^-17 relative (numel % 8 == 0)."""
    _need_cuda(x)
    x = x.contiguous().float()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    st = _lib.load().e2f_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    _lib.check(st, "e2f_split_bf16")
    return hi, lo


_WEIGHT_SPLITS = {}  # id(Parameter) -> (weakref to it, (version, data_ptr), hi, lo); dropped when the parameter dies


def _split_weight(weight, k_pad=0):
    """bf16 (hi, lo) split of a (N, K) weight, cached per parameter; k_pad > K appends zero columns (for an A operand
    whose rows are padded to k_pad)."""
    key = (id(weight), k_pad)
    tag = (weight._version, weight.data_ptr())
    hit = _WEIGHT_SPLITS.get(key)
    if hit is None or hit[0]() is not weight or hit[1] != tag:
        w2 = weight.detach().reshape(weight.shape[0], -1)
        if k_pad > w2.shape[1]:
            w2 = torch.nn.functional.pad(w2.float(), (0, k_pad - w2.shape[1]))
        hi, lo = split_bf16(w2)
        if hit is None or hit[0]() is not weight:
            weakref.finalize(weight, _WEIGHT_SPLITS.pop, key, None)
        hit = (weakref.ref(weight), tag, hi, lo)
        _WEIGHT_SPLITS[key] = hit
    return hit[2], hit[3]


def linear(x, weight, bias=None, residual=None, out_dtype=torch.float32, tile_hint=0):
    """``F.linear(x, weight, bias) (+ residual)`` on the bf16 tensor pipe with a 3-term split (fp32-level accuracy).

    x (..., K) fp32, weight (N, K) fp32 nn.Parameter (its bf16 split is cached per parameter object and refreshed
    when the parameter changes), bias (N,), residual (..., N) fp32 -> (..., N) ``out_dtype``."""
    n, k = weight.shape[:2]            # (N, K) Linear weight or (N, K, 1, 1) 1x1-conv weight
    lead = x.shape[:-1]
    k_pad = 0
    if isinstance(x, SplitMat):         # operand pair written by a fused producer
        _need_cuda(weight, bias, residual)
        if x.shape[-1] != k:            # rows padded with zero columns (t2t_fold_unfold pitch): pad the weight alike
            if x.shape[-1] < k:
                raise ValueError(f"linear: operand has {x.shape[-1]} columns, weight expects {k}")
            k_pad = k = x.shape[-1]
        a_hi, a_lo = x.hi.reshape(-1, k), x.lo.reshape(-1, k)
        m = a_hi.shape[0]
    else:
        _need_cuda(x, weight, bias, residual)
        x2 = x.reshape(-1, k)
        m = x2.shape[0]
        a_hi, a_lo = split_bf16(x2)
    w_hi, w_lo = _split_weight(weight, k_pad)
    b32 = None if bias is None else bias.detach().float().contiguous()
    res = None
    if residual is not None:
        res = residual.reshape(m, n).contiguous().float()
    out = torch.empty((m, n), dtype=out_dtype, device=weight.device)
    with _timed("linear_bf16x3", 2.0 * m * n * k):
        st = _lib.load().e2f_linear_bf16x3(a_hi.data_ptr(), a_lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(),
                                           None if b32 is None else b32.data_ptr(),
                                           None if res is None else res.data_ptr(), out.data_ptr(), m, n, k,
                                           _DT[out_dtype], tile_hint, _stream())
    _lib.check(st, "e2f_linear_bf16x3")
    return out.view(*lead, n)


class SplitNHWC:
    """An NHWC activation as two bf16 terms (hi + lo) — the operand format of ``conv3x3``. Reusable across convs."""

    __slots__ = ("hi", "lo", "shape")

    def __init__(self, hi, lo, shape):
        self.hi, self.lo, self.shape = hi, lo, shape   # hi/lo: (N,H,W,Cpad) bf16; shape: logical (N,C,H,W)


def split_nhwc(x):
    """(N,C,H,W) fp32 (any memory format) -> SplitNHWC with channels zero-padded to a multiple of 8."""
    if isinstance(x, SplitNHWC):
        return x
    _need_cuda(x)
    n, c, h, w = x.shape
    nhwc = x.permute(0, 2, 3, 1)
    if c % 8:
        nhwc = torch.nn.functional.pad(nhwc, (0, 8 - c % 8))
    hi, lo = split_bf16(nhwc)            # .contiguous() inside is a no-op for channels_last inputs
    return SplitNHWC(hi, lo, (n, c, h, w))


class RowsNHWC:
    """A small-channel activation in the row-gapped layout of the window-packed conv (include/e2fgvi_b200.h,
    ``e2f_conv2d_rows_bf16x3``): bf16 (hi, lo), flat [N][H][pitch][cin] + tail with pitch = lead + W (+1 rounding
    pixel for odd rows of 4 channels), zeros in gaps / tail / pad channels."""

    __slots__ = ("hi", "lo", "shape", "lead", "cin", "pitch")

    def __init__(self, hi, lo, shape, lead, cin):
        self.hi, self.lo, self.shape, self.lead, self.cin = hi, lo, shape, lead, cin   # shape: logical (N,C,H,W)
        self.pitch = int(_lib.load().e2f_conv_rows_pitch(shape[3], lead, cin))

    def dense(self):
        """fp32 (N,C,H,W) view of the content (tests / debugging)."""
        n, c, h, w = self.shape
        body = (self.hi.float() + self.lo.float())[: n * h * self.pitch * self.cin]
        return body.view(n, h, self.pitch, self.cin)[:, :, self.lead: self.lead + w, :c].permute(0, 3, 1, 2)


def rows_channels(c):
    """Channel count of the row-gapped layout that holds c channels (4, 8, 16 or 32), or None if c > 32."""
    for cin in (4, 8, 16, 32):
        if c <= cin:
            return cin
    return None


def _rows_numel(n, h, w, lead, cin):
    lib = _lib.load()
    return (n * h * int(lib.e2f_conv_rows_pitch(w, lead, cin)) + int(lib.e2f_conv_rows_tail(lead, cin))) * cin


def pack_rows(x, lead, cin=None):
    """(N,C,H,W) fp32 -> ``RowsNHWC`` with ``lead`` zero pixels in front of every row (= the padding of the conv
    that consumes it) and channels zero-padded to ``cin`` (default: the smallest of 4/8/16/32 that holds C)."""
    if isinstance(x, RowsNHWC):
        return x
    _need_cuda(x)
    n, c, h, w = x.shape
    cin = cin or rows_channels(c)
    if cin is None or c > cin:
        raise ValueError(f"pack_rows: {c} channels do not fit a row-gapped layout (<= 32)")
    x = x.contiguous().float()
    numel = _rows_numel(n, h, w, lead, cin)
    hi = torch.empty(numel, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(numel, dtype=torch.bfloat16, device=x.device)
    with _timed("pack_rows", float(x.numel() * 4 + numel * 4)):
        st = _lib.load().e2f_pack_rows_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), n, c, h, w, cin, lead, _stream())
    _lib.check(st, "e2f_pack_rows_bf16")
    return RowsNHWC(hi, lo, (n, c, h, w), lead, cin)


def pack_conv_rows_weight(weight, cin):
    """fp32 (Cout, C, k, k), C <= cin -> (hi, lo) bf16 (Cout, k*G*64) in the window-packed K order of
    ``e2f_conv2d_rows_bf16x3``: chunk (ky, g) holds taps kx = g*PX .. g*PX+PX-1 (PX = 64/cin) x cin channels, zeros
    for kx >= k and for channels >= C."""
    cout, c, kh, kw = weight.shape
    assert kh == kw and c <= cin
    px = 64 // cin
    g = (kw + px - 1) // px
    w = weight.detach().float()
    packed = torch.zeros((cout, kh, g * px, cin), dtype=torch.float32, device=weight.device)
    packed[:, :, :kw, :c] = w.permute(0, 2, 3, 1)          # [co][ky][kx][c]
    return split_bf16(packed.view(cout, kh * g * 64))


def pack_conv3x3_weight(weight, src_channels, groups=1):
    """fp32 (Cout, Cin/G, 3, 3) -> (hi, lo) bf16 (Cout, 9*T*64) in the K order of ``e2f_conv3x3_bf16x3``:
    tap-major, then source, then 64-channel chunk; the group-local input channel axis is the concatenation of the
    sources' per-group slices (exactly the channel order of the torch.cat the reference performs)."""
    cout, cin_g, kh, kw = weight.shape
    assert kh == kw
    taps = kh * kw
    cig = [c // groups for c in src_channels]
    assert sum(cig) == cin_g, (src_channels, groups, cin_g)
    chunks = [(c + 63) // 64 for c in cig]
    T = sum(chunks)
    w = weight.detach().float()
    packed = torch.zeros((cout, taps, T, 64), dtype=torch.float32, device=weight.device)
    off, base = 0, 0
    for c, nch in zip(cig, chunks):
        for j in range(nch):
            cc = min(64, c - 64 * j)
            packed[:, :, base + j, :cc] = w[:, off + 64 * j: off + 64 * j + cc].reshape(cout, cc, taps).transpose(1, 2)
        off += c
        base += nch
    return split_bf16(packed.view(cout, taps * T * 64))


def merge_conv_groups(weight, src_channels, groups):
    """Rewrite a grouped conv whose groups have few output channels as one with fewer, wider groups and
    block-diagonal weights: m = largest power of two with m * Cout/G <= 64 and G % m == 0 groups are merged, so a
    64-wide N tile and the 64-channel K chunks of every source are filled instead of padded (encoder conv 7,
    e2fgvi.py:97: G=8 with 32 outputs and 32 + 48 inputs per group).  Returns (weight', groups')."""
    cout, cin_g = weight.shape[0], weight.shape[1]
    cog = cout // groups
    m = 1
    while groups % (2 * m) == 0 and 2 * m * cog <= 64:
        m *= 2
    if m == 1:
        return weight, groups
    cig = [c // groups for c in src_channels]
    w = weight.detach().float()
    merged = torch.zeros((cout, m * cin_g) + tuple(weight.shape[2:]), dtype=torch.float32, device=weight.device)
    sub = (torch.arange(cout, device=weight.device) % (m * cog)) // cog      # position of o's group in its merge
    off = 0
    for c in cig:
        for j in range(m):
            rows = (sub == j).nonzero().flatten()
            merged[rows, m * off + j * c: m * off + (j + 1) * c] = w[rows, off: off + c]
        off += c
    return merged, groups // m


_CONV_PACKS = {}  # (id(Parameter), src channels, groups) -> (weakref, tag, hi, lo, effective groups)


def _packed_conv_weight(weight, src_channels, groups, rows_cin=0):
    key = (id(weight), tuple(src_channels), groups, rows_cin)
    tag = (weight._version, weight.data_ptr())
    hit = _CONV_PACKS.get(key)
    if hit is None or hit[0]() is not weight or hit[1] != tag:
        if rows_cin:
            (hi, lo), g_eff = pack_conv_rows_weight(weight, rows_cin), 1
        else:
            w_eff, g_eff = merge_conv_groups(weight, src_channels, groups)
            hi, lo = pack_conv3x3_weight(w_eff, src_channels, g_eff)
        if hit is None or hit[0]() is not weight:
            weakref.finalize(weight, _CONV_PACKS.pop, key, None)
        hit = (weakref.ref(weight), tag, hi, lo, g_eff)
        _CONV_PACKS[key] = hit
    return hit[2], hit[3], hit[4]


def conv3x3(sources, weight, bias=None, groups=1, negative_slope=1.0, residual=None, out="f32", stride=1,
            padding=None, out_lead=0):
    """``leaky_relu(F.conv2d(torch.cat(sources, 1) [group-wise for groups > 1], weight, bias, 1, 1, 1, groups),
    negative_slope) (+ residual)`` as one tcgen05 implicit-GEMM launch; the cat is never built.

    Also serves square k x k kernels (k = 3, 7) with stride 1 / 2 (``padding`` defaults to k // 2): the stride-2
    encoder convs and SPyNet's 7x7 convs (negative_slope = 0 is ReLU).

    sources: list of (N,C_i,H,W) fp32 tensors or ``SplitNHWC``, or ONE ``RowsNHWC`` (small-channel input, window-packed
    K; its lead must equal the padding); weight: the nn.Conv2d parameter (Cout, sum C_i / G, k, k).
    out = "f32": (N,Cout,H,W) fp32 channels_last tensor; "split": a ``SplitNHWC`` (the bf16 operand pair of a
    following conv3x3, written by the epilogue, no fp32 round trip); "both": (tensor, SplitNHWC); "rows": a
    ``RowsNHWC`` with ``out_lead`` zero pixels per row (operand of a following small-channel conv with that padding)."""
    cout, ks = weight.shape[0], weight.shape[2]
    pad = ks // 2 if padding is None else padding
    rows_in = sources if isinstance(sources, RowsNHWC) else (
        sources[0] if isinstance(sources, (list, tuple)) and len(sources) == 1 and isinstance(sources[0], RowsNHWC) else None)
    _need_cuda(weight, bias, residual)
    if rows_in is not None:
        if rows_in.lead != pad or groups != 1:
            raise ValueError(f"conv3x3: RowsNHWC source with lead {rows_in.lead} needs padding {rows_in.lead} and groups 1")
        if rows_in.shape[1] != weight.shape[1]:
            raise ValueError("conv3x3: weight does not match the source's channel count")
        splits = [rows_in]
        true_channels = padded_channels = [rows_in.cin]
    else:
        splits = [split_nhwc(s) for s in (sources if isinstance(sources, (list, tuple)) else [sources])]
        true_channels = [s.shape[1] for s in splits]
        padded_channels = [s.hi.shape[-1] for s in splits]
    n, _, h, w = splits[0].shape
    for s in splits:
        if (s.shape[0], s.shape[2], s.shape[3]) != (n, h, w):
            raise ValueError("conv3x3 sources must share N, H, W")
    if groups != 1 and true_channels != padded_channels:
        raise NotImplementedError("grouped conv3x3 needs channel counts that are multiples of 8")
    h_in, w_in = h, w
    h, w = (h_in + 2 * pad - ks) // stride + 1, (w_in + 2 * pad - ks) // stride + 1
    w_hi, w_lo, g_eff = _packed_conv_weight(weight, true_channels, groups, rows_in.cin if rows_in is not None else 0)
    b32 = None if bias is None else bias.detach().float().contiguous()
    res = None
    if residual is not None:
        res = residual.permute(0, 2, 3, 1).contiguous().float()
    want_f32, want_split, want_rows = out in ("f32", "both"), out in ("split", "both"), out == "rows"
    if not (want_f32 or want_split or want_rows):
        raise ValueError("out must be 'f32', 'split', 'both' or 'rows'")
    if want_rows and (out_lead <= 0 or cout not in (8, 16, 32)):
        raise ValueError("conv3x3: out='rows' needs out_lead > 0 and 8, 16 or 32 output channels")
    dev = weight.device
    o32 = torch.empty((n, h, w, cout), dtype=torch.float32, device=dev) if want_f32 else None
    ohi = olo = None
    if want_split:
        ohi = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=dev)
        olo = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=dev)
    elif want_rows:
        numel = _rows_numel(n, h, w, out_lead, cout)
        ohi = torch.empty(numel, dtype=torch.bfloat16, device=dev)
        olo = torch.empty(numel, dtype=torch.bfloat16, device=dev)
    k = len(splits)
    hi_arr = (_lib._vp * k)(*[s.hi.data_ptr() for s in splits])
    lo_arr = (_lib._vp * k)(*[s.lo.data_ptr() for s in splits])
    ch_arr = (_lib._i * k)(*padded_channels)
    with _timed("conv3x3_bf16x3", 2.0 * n * h * w * cout * weight.shape[1] * ks * ks):
        st = _lib.load().e2f_conv2d_rows_bf16x3(k, hi_arr, lo_arr, ch_arr, 1 if rows_in is not None else 0,
                                                w_hi.data_ptr(), w_lo.data_ptr(),
                                                None if b32 is None else b32.data_ptr(),
                                                None if res is None else res.data_ptr(),
                                                None if o32 is None else o32.data_ptr(),
                                                None if ohi is None else ohi.data_ptr(),
                                                None if olo is None else olo.data_ptr(),
                                                out_lead if want_rows else 0, n, h_in, w_in, cout,
                                                g_eff, float(negative_slope), ks, stride, pad, _stream())
    _lib.check(st, "e2f_conv2d_rows_bf16x3")
    if want_rows:
        return RowsNHWC(ohi, olo, (n, cout, h, w), out_lead, cout)
    t32 = o32.permute(0, 3, 1, 2) if want_f32 else None
    sp = SplitNHWC(ohi, olo, (n, cout, h, w)) if want_split else None
    return t32 if out == "f32" else sp if out == "split" else (t32, sp)


# ------------------------------------------------------------------------------------------------- SoftSplit / SoftComp
KXN_CONVS = os.environ.get("E2F_KXN", "1") != "0"     # small-Cout layers on the kx-in-N kernel (E2F_KXN=0: A/B, debugging)

_DERIVED = {}   # (id(param), tag...) -> (weakref, (version, data_ptr) tuple, value): weight-derived operands, per parameter


def _derived(params, tag, build):
    """Cache of operands derived from one or more nn.Parameters (packed weights, folded bias maps); rebuilt when any
    of them changes version or address, dropped when the first one dies."""
    key = (tuple(id(p) for p in params),) + tuple(tag)
    stamp = tuple((p._version, p.data_ptr()) for p in params)
    hit = _DERIVED.get(key)
    if hit is None or any(r() is not p for r, p in zip(hit[0], params)) or hit[1] != stamp:
        val = build()
        if hit is None or hit[0][0]() is not params[0]:
            weakref.finalize(params[0], _DERIVED.pop, key, None)
        hit = (tuple(weakref.ref(p) for p in params), stamp, val)
        _DERIVED[key] = hit
    return hit[2]


def _best_tile(gh, gw, stride):
    """(tile_w, tile_h) with tile_w * tile_h <= 128 that wastes the fewest accumulator rows on a gh x gw GEMM grid
    (12 x 10 tiles the 20x36 and 60x108 token grids exactly; 18 x 7 the 90x162 one)."""
    best, best_cost = (16, 8), None
    for tw in range(1, min(gw, 128, 256 // stride) + 1):
        th = min(128 // tw, gh, 256 // stride)
        cost = -(-gh // th) * -(-gw // tw)                 # number of 128-row tiles
        key = (cost, abs(tw - th), -tw * th)            # fewest tiles, then the squarest (best TMA box / L2 reuse)
        if best_cost is None or key < best_cost:
            best, best_cost = (tw, th), key
    return best


def _as_split_nhwc(x):
    if isinstance(x, SplitNHWC):
        return x
    return split_nhwc(x)


def _nhwc_stride(t, what):
    """Batch stride in PIXELS of a (n, h, w, c) tensor whose inner three dims are dense (a frame slice of a
    (b, t, h, w, c) buffer is such a view)."""
    n, h, w, c = t.shape
    if t.stride()[1:] != (w * c, c, 1) or (n > 1 and t.stride(0) % c):
        raise ValueError(f"{what}: expected a (n, h, w, c) tensor with dense (h, w, c) dims, got strides {t.stride()}")
    return (t.stride(0) // c) if n > 1 else h * w


def _conv_gather(sources, w_hi, w_lo, bias, bias_map, residual, out, cout, stride, grid, taps, phases, ostep,
                 out_size, flops, slope=1.0, into=None):
    """One launch of e2f_conv_gather_bf16x3.  sources: list of ``SplitNHWC`` (hi / lo may be batch-strided views);
    taps: list of (dy, dx); phases: list of (first tap, oy, ox); residual: logical (n, cout, oh, ow) fp32 with NHWC
    storage; ``into`` = (o32, ohi, olo) optional pre-existing (n, oh, ow, cout) NHWC views to write (batch-strided allowed,
    all with the same batch stride, which the residual must share)."""
    import ctypes
    sources = list(sources)
    n, _, h_in, w_in = sources[0].shape
    gh, gw = grid
    oh, ow = out_size
    tw, th = _best_tile(gh, gw, stride)
    dev = sources[0].hi.device
    want_f32, want_split = out in ("f32", "both"), out in ("split", "both")
    if not (want_f32 or want_split):
        raise ValueError("out must be 'f32', 'split' or 'both'")
    o32 = ohi = olo = None
    if into is not None:
        o32, ohi, olo = into
        if (want_f32 and o32 is None) or (want_split and (ohi is None or olo is None)):
            raise ValueError("conv: `into` lacks a buffer for the requested output")
        o32 = o32 if want_f32 else None
        ohi, olo = (ohi, olo) if want_split else (None, None)
    else:
        if want_f32:
            o32 = torch.empty((n, oh, ow, cout), dtype=torch.float32, device=dev)
        if want_split:
            ohi = torch.empty((n, oh, ow, cout), dtype=torch.bfloat16, device=dev)
            olo = torch.empty((n, oh, ow, cout), dtype=torch.bfloat16, device=dev)
    outs = [t for t in (o32, ohi, olo) if t is not None]
    for t in outs:
        if tuple(t.shape) != (n, oh, ow, cout):
            raise ValueError(f"conv: output buffer {tuple(t.shape)} != {(n, oh, ow, cout)}")
    ostrides = {_nhwc_stride(t, "conv output") for t in outs}
    res = None
    if residual is not None:
        res = residual.permute(0, 2, 3, 1)
        if res.dtype != torch.float32 or res.stride()[1:] != (ow * cout, cout, 1):
            res = res.contiguous().float()
        ostrides.add(_nhwc_stride(res, "conv residual"))
    if len(ostrides) != 1:
        raise ValueError(f"conv: outputs and residual must share one batch stride, got {sorted(ostrides)}")
    out_nstride = ostrides.pop()
    k = len(sources)
    for s_ in sources:
        if (s_.shape[0], s_.shape[2], s_.shape[3]) != (n, h_in, w_in):
            raise ValueError("conv sources must share N, H, W")
    nt, nph = len(taps), len(phases)
    dy = (ctypes.c_int8 * nt)(*[t[0] for t in taps])
    dx = (ctypes.c_int8 * nt)(*[t[1] for t in taps])
    tap0 = (ctypes.c_uint8 * (nph + 1))(*([p[0] for p in phases] + [nt]))
    oy = (ctypes.c_uint8 * nph)(*[p[1] for p in phases])
    ox = (ctypes.c_uint8 * nph)(*[p[2] for p in phases])
    hi_arr = (_lib._vp * k)(*[s_.hi.data_ptr() for s_ in sources])
    lo_arr = (_lib._vp * k)(*[s_.lo.data_ptr() for s_ in sources])
    ch_arr = (_lib._i * k)(*[s_.hi.shape[-1] for s_ in sources])
    sn = [_nhwc_stride(s_.hi, "conv source") for s_ in sources]
    for s_, v in zip(sources, sn):
        if _nhwc_stride(s_.lo, "conv source") != v:
            raise ValueError("conv: hi / lo of a source must share the batch stride")
    sn_arr = (ctypes.c_int64 * k)(*sn)
    b32 = None if bias is None else bias.detach().float().contiguous()
    with _timed("conv3x3_bf16x3", flops):
        st = _lib.load().e2f_conv_gather_bf16x3(
            k, hi_arr, lo_arr, ch_arr, w_hi.data_ptr(), w_lo.data_ptr(), None if b32 is None else b32.data_ptr(),
            None if bias_map is None else bias_map.data_ptr(), None if res is None else res.data_ptr(),
            None if o32 is None else o32.data_ptr(), None if ohi is None else ohi.data_ptr(),
            None if olo is None else olo.data_ptr(), n, h_in, w_in, cout, float(slope), stride, gh, gw, tw, th, nt, dy, dx,
            nph, tap0, oy, ox, ostep, oh, ow, sn_arr, out_nstride, _stream())
    _lib.check(st, "e2f_conv_gather_bf16x3")
    t32 = o32.permute(0, 3, 1, 2) if want_f32 else None
    sp = SplitNHWC(ohi, olo, (n, cout, oh, ow)) if want_split else None
    return t32 if out == "f32" else sp if out == "split" else (t32, sp)


def conv_frames(sources, weight, bias=None, negative_slope=1.0, residual=None, out="f32", into=None):
    """``conv3x3`` (k x k, stride 1, pad k//2, groups 1) whose sources, residual and outputs may be FRAME SLICES of
    (b, t, h, w, c) buffers (batch-strided NHWC views): the per-frame tensors of BidirectionalPropagation
    (feat_prop.py:88-149) are read and written in place — no ``x[:, i].contiguous()`` gathers, no stack / cat copies.
    sources: list of ``SplitNHWC`` (or fp32 tensors, split on the fly); ``into`` = (o32, ohi, olo) NHWC views or None."""
    cout, ks = weight.shape[0], weight.shape[2]
    pad = ks // 2
    _need_cuda(weight, bias, residual)
    splits = [_as_split_nhwc(s_) for s_ in (sources if isinstance(sources, (list, tuple)) else [sources])]
    chans = [s_.shape[1] for s_ in splits]         # true channel counts (storage may be zero-padded to a multiple of 8)
    if sum(chans) != weight.shape[1]:
        raise ValueError(f"conv_frames: source channels {chans} do not add up to the weight's {weight.shape[1]} input channels")
    n, _, h, w = splits[0].shape
    w_hi, w_lo, _ = _packed_conv_weight(weight, chans, 1)
    taps = [(ky - pad, kx - pad) for ky in range(ks) for kx in range(ks)]
    return _conv_gather(splits, w_hi, w_lo, bias, None, residual, out, cout, 1, (h, w), taps, [(0, 0, 0)], 1, (h, w),
                        2.0 * n * h * w * cout * weight.shape[1] * ks * ks, slope=negative_slope, into=into)


def soft_split(x, weight, bias, kernel_size, stride, padding):
    """SoftSplit.forward (tfocal_transformer.py:39-46): ``Linear(unfold(x, k, s, p))`` evaluated as the k x k / stride-s
    convolution it is — an implicit GEMM whose A operand is fetched by TMA straight from the NHWC feature map; the
    k*k-times larger unfolded token matrix is never built.

    x (BT,C,H,W) fp32 (any memory format) or ``SplitNHWC``; weight (hidden, C*k*k) = ``ss.embedding.weight``;
    bias (hidden,).  Returns tokens (BT, fh*fw, hidden) fp32 — exactly ``embedding(unfold(x).permute(0, 2, 1))``."""
    (k, k2), (s, s2), (p, p2) = _pair(kernel_size), _pair(stride), _pair(padding)
    if k != k2 or s != s2 or p != p2:
        raise NotImplementedError("square kernel / stride / padding only (E2FGVI uses 7 / 3 / 3)")
    src = _as_split_nhwc(x)
    n, c, h, w = src.shape
    hidden = weight.shape[0]
    if weight.shape[1] != c * k * k or k * k > 64 or c % 8:
        raise ValueError(f"soft_split: weight {tuple(weight.shape)} does not match C={c}, k={k}")
    _need_cuda(weight, bias)
    fh, fw = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    w_hi, w_lo = _derived([weight], ("soft_split", c, k),
                          lambda: pack_conv3x3_weight(weight.detach().view(hidden, c, k, k), [c], 1))
    taps = [(ky - p, kx - p) for ky in range(k) for kx in range(k)]
    tok = _conv_gather([src], w_hi, w_lo, bias, None, None, "f32", hidden, s, (fh, fw), taps, [(0, 0, 0)], 1,
                       (fh, fw), 2.0 * n * fh * fw * hidden * c * k * k)
    return tok.permute(0, 2, 3, 1).reshape(n, fh * fw, hidden)       # NHWC storage: a view


def _soft_comp_tables(k, s, p):
    """Phases and taps of the transposed conv that Linear + fold(k, s, p) is, for p == s (E2FGVI: 7 / 3 / 3): output
    pixel (s*a + ry, s*b + rx) sums token (a + 1 - dy, b + 1 - dx) times the (ky, kx) = (s*dy + ry, s*dx + rx) slice
    of the weight over all dy, dx with ky, kx < k.  Returns (taps [(dy_tok, dx_tok)], phases [(tap0, ry, rx)],
    kernel positions [(ky, kx)] in tap order)."""
    if p != s:
        raise NotImplementedError("soft_comp: the phase decomposition is implemented for padding == stride (7 / 3 / 3)")
    taps, phases, kpos = [], [], []
    for ry in range(s):
        for rx in range(s):
            phases.append((len(taps), ry, rx))
            for dy in range((k - 1 - ry) // s + 1):
                for dx in range((k - 1 - rx) // s + 1):
                    taps.append((1 - dy, 1 - dx))
                    kpos.append((s * dy + ry, s * dx + rx))
    return taps, phases, kpos


def soft_comp(tokens, weight, bias, output_size, kernel_size, stride, padding, bias_map_extra=None, residual=None,
              out="f32"):
    """SoftComp.forward up to (and including) the fold (tfocal_transformer.py:65-72, _hq.py:67-79):
    ``fold(Linear(tokens))`` evaluated as the transposed convolution it is, one implicit-GEMM launch over the nine
    output phases; the (C*k*k)-wide token matrix and the fold pass never exist.

    tokens (BT, fh, fw, hidden) fp32 or ``SplitMat``; weight (C*k*k, hidden) = ``sc.embedding.weight``; bias (C*k*k,);
    ``bias_map_extra``: the base model's ``sc.bias`` (C, H, W) parameter or None; ``residual`` (BT, C, H, W) fp32 added
    in the epilogue (enc_feat + trans_feat, e2fgvi.py:263).  Returns (BT, C, H, W) fp32 channels_last (out="f32"),
    a ``SplitNHWC`` (out="split", operand of the HQ model's bias_conv) or both."""
    (k, k2), (s, s2), (p, p2) = _pair(kernel_size), _pair(stride), _pair(padding)
    if k != k2 or s != s2 or p != p2:
        raise NotImplementedError("square kernel / stride / padding only (E2FGVI uses 7 / 3 / 3)")
    h, w = output_size
    fh, fw = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    hidden = weight.shape[1]
    c = weight.shape[0] // (k * k)
    if isinstance(tokens, SplitMat):
        hi, lo = tokens.hi, tokens.lo
    else:
        _need_cuda(tokens)
        hi, lo = split_bf16(tokens)
    n = hi.numel() // (fh * fw * hidden)
    if hi.numel() != n * fh * fw * hidden or hidden % 64 or c * k * k != weight.shape[0]:
        raise ValueError(f"soft_comp: tokens {tuple(hi.shape)} / weight {tuple(weight.shape)} do not match output_size "
                         f"{output_size}")
    _need_cuda(weight, bias, bias_map_extra, residual)
    if residual is not None and tuple(residual.shape) != (n, c, h, w):
        raise ValueError(f"residual {tuple(residual.shape)} != {(n, c, h, w)}")
    taps, phases, kpos = _soft_comp_tables(k, s, p)
    src = SplitNHWC(hi.view(n, fh, fw, hidden), lo.view(n, fh, fw, hidden), (n, hidden, fh, fw))

    def pack():
        order = torch.tensor([ky * k + kx for ky, kx in kpos], device=weight.device)
        wp = weight.detach().float().view(c, k * k, hidden)[:, order, :].reshape(c, k * k * hidden)
        return split_bf16(wp)

    w_hi, w_lo = _derived([weight], ("soft_comp", k, s), pack)

    def fold_bias():
        # fold of the Linear bias: every pixel receives b[c, ky, kx] from each token patch that covers it (border
        # pixels from fewer patches) + the base model's learned (C, H, W) bias map; layout [H][W][C]
        F = torch.nn.functional
        m = torch.zeros((1, c, h, w), dtype=torch.float32, device=weight.device)
        if bias is not None:
            cols = bias.detach().float().view(1, c * k * k, 1).expand(1, c * k * k, fh * fw)
            m = F.fold(cols, (h, w), (k, k), padding=(p, p), stride=(s, s))
        if bias_map_extra is not None:
            m = m + bias_map_extra.detach().float().view(1, c, h, w)
        return m[0].permute(1, 2, 0).contiguous()

    bparams = [q for q in (bias, bias_map_extra) if q is not None]
    bias_map = _derived(bparams, ("soft_comp_bias", h, w, k, s), fold_bias) if bparams else None
    return _conv_gather([src], w_hi, w_lo, None, bias_map, residual, out, c, 1, (fh, fw), taps, phases, s,
                        (h, w), 2.0 * n * fh * fw * hidden * c * k * k)


def pack_conv_kxn_weight(weight, co_pad, src_channels=None, groups=1):
    """fp32 (Cout, Cin/G, k, k) -> (hi, lo) bf16 (G*k*co_pad, k*chunks*64) in the operand order of ``e2f_conv_kxn_bf16x3``:
    row = g*k*co_pad + kx*co_pad + co, column = (ky*chunks + chunk)*64 + channel, the chunks of source 0 first, then source
    1 (the group-local input channel axis is the concatenation of the sources' per-group slices, like
    ``pack_conv3x3_weight``); zeros for co >= Cout/G and for padded channels."""
    cout, cin_g, kh, kw = weight.shape
    assert kh == kw and cout % groups == 0
    cog = cout // groups
    assert cog <= co_pad
    src_channels = [cin_g * groups] if src_channels is None else list(src_channels)
    cig = [c // groups for c in src_channels]
    assert sum(cig) == cin_g, (src_channels, groups, cin_g)
    nch = [(c + 63) // 64 for c in cig]
    chunks = sum(nch)
    w = weight.detach().float().view(groups, cog, cin_g, kh, kw)
    packed = torch.zeros((groups, kw, co_pad, kh, chunks * 64), dtype=torch.float32, device=weight.device)
    off, base = 0, 0
    for c, n_ in zip(cig, nch):
        # [g][co][c][ky][kx] -> [g][kx][co][ky][c]
        packed[:, :, :cog, :, base * 64: base * 64 + c] = w[:, :, off: off + c].permute(0, 4, 1, 3, 2)
        off += c
        base += n_
    return split_bf16(packed.view(groups * kw * co_pad, kh * chunks * 64))


def _kxn_co_pad(cout, ks):
    """Smallest padded channel count (multiple of 8, <= 32) with ks*co_pad % 16 == 0 that holds cout, or None."""
    for cp in (8, 16, 24, 32):
        if cp >= cout and (ks * cp) % 16 == 0:
            return cp
    return None


def conv_kxn(x, weight, bias=None, negative_slope=1.0, residual=None, out="f32", tanh_nchw=False, groups=1):
    """k x k / stride 1 / pad k//2 conv (k = 3 or 7) with FEW output channels per group (<= 32) on the "kx-in-N" kernel:
    SPyNet's 64 -> 32 / 32 -> 16 / 16 -> 2 convs (flow_comp.py:181-215), the decoder's output conv (e2fgvi.py:149-150) and
    the encoder's groups-of-32 conv (e2fgvi.py:97; two sources, group-wise concatenation never built).
    x: (N,C,H,W) fp32 / ``SplitNHWC`` or a list of up to two of them; residual (N,Cout,H,W) logical with NHWC storage;
    out = "f32" | "split" | "both"; ``tanh_nchw``: tanh + contiguous NCHW fp32 result (the prediction, e2fgvi.py:262)."""
    srcs = [_as_split_nhwc(s_) for s_ in (x if isinstance(x, (list, tuple)) else [x])]
    n, _, h, w = srcs[0].shape
    chans = [s_.shape[1] for s_ in srcs]
    cout, cin_g, ks, ks2 = weight.shape
    co_pad = _kxn_co_pad(cout // groups, ks) if cout % groups == 0 else None
    if (ks != ks2 or ks not in (3, 7) or sum(chans) != cin_g * groups or co_pad is None or len(srcs) > 2
            or any(c % groups or s_.hi.shape[-1] != c for c, s_ in zip(chans, srcs))):
        raise ValueError(f"conv_kxn: unsupported weight {tuple(weight.shape)} / groups {groups} for sources {chans}")
    for s_ in srcs:
        if (s_.shape[0], s_.shape[2], s_.shape[3]) != (n, h, w) or not s_.hi.is_contiguous() or not s_.lo.is_contiguous():
            raise ValueError("conv_kxn: sources must be dense and share N, H, W")
    _need_cuda(weight, bias, residual)
    w_hi, w_lo = _derived([weight], ("kxn", co_pad, tuple(chans), groups),
                          lambda: pack_conv_kxn_weight(weight, co_pad, chans, groups))
    b32 = None if bias is None else bias.detach().float().contiguous()
    dev = weight.device
    want_f32, want_split = out in ("f32", "both"), out in ("split", "both")
    if tanh_nchw and out != "f32":
        raise ValueError("conv_kxn: tanh_nchw returns the fp32 NCHW tensor only")
    if want_split and (cout // groups) % 8:
        raise ValueError("conv_kxn: split output needs Cout / groups % 8 == 0")
    res = None
    if residual is not None:
        res = residual.permute(0, 2, 3, 1).contiguous().float()          # no-op for NHWC storage
    o32 = ohi = olo = None
    if want_f32:
        o32 = torch.empty((n, cout, h, w) if tanh_nchw else (n, h, w, cout), dtype=torch.float32, device=dev)
    if want_split:
        ohi = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=dev)
        olo = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=dev)
    k = len(srcs)
    hi_arr = (_lib._vp * k)(*[s_.hi.data_ptr() for s_ in srcs])
    lo_arr = (_lib._vp * k)(*[s_.lo.data_ptr() for s_ in srcs])
    ch_arr = (_lib._i * k)(*[s_.hi.shape[-1] for s_ in srcs])
    with _timed("conv3x3_bf16x3", 2.0 * n * h * w * cout * cin_g * ks * ks):
        st = _lib.load().e2f_conv_kxn_bf16x3(k, hi_arr, lo_arr, ch_arr, w_hi.data_ptr(), w_lo.data_ptr(),
                                             None if b32 is None else b32.data_ptr(),
                                             None if res is None else res.data_ptr(),
                                             None if o32 is None else o32.data_ptr(), None if ohi is None else ohi.data_ptr(),
                                             None if olo is None else olo.data_ptr(), n, h, w, cout, groups, co_pad, ks,
                                             float(negative_slope), 3 if tanh_nchw else 0, _stream())
    _lib.check(st, "e2f_conv_kxn_bf16x3")
    if tanh_nchw:
        return o32
    t32 = o32.permute(0, 3, 1, 2) if want_f32 else None
    sp = SplitNHWC(ohi, olo, (n, cout, h, w)) if want_split else None
    return t32 if out == "f32" else sp if out == "split" else (t32, sp)


def conv3x3_tanh_nchw(x, weight, bias):
    """``torch.tanh(F.conv2d(x, weight, bias, 1, 1))`` returned as a CONTIGUOUS (N, Cout, H, W) fp32 tensor: the
    decoder's 64 -> 3 output conv (e2fgvi.py:149-150) with the tanh of :262 and the NHWC -> NCHW layout change of the
    prediction fused into the conv epilogue.  x: (N,C,H,W) fp32 or ``SplitNHWC``."""
    if KXN_CONVS and weight.shape[0] <= 32 and tuple(weight.shape[2:]) == (3, 3):
        return conv_kxn(x, weight, bias, tanh_nchw=True)
    src = _as_split_nhwc(x)
    n, c, h, w = src.shape
    cout = weight.shape[0]
    _need_cuda(weight, bias)
    if tuple(weight.shape[1:]) != (c, 3, 3) or cout > 32 or cout % 4 == 0:
        raise ValueError(f"conv3x3_tanh_nchw: weight {tuple(weight.shape)} (needs (Cout, {c}, 3, 3), Cout <= 32, Cout % 4 != 0)")
    w_hi, w_lo, _ = _packed_conv_weight(weight, [c], 1)
    b32 = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty((n, cout, h, w), dtype=torch.float32, device=weight.device)
    with _timed("conv3x3_bf16x3", 2.0 * n * h * w * cout * c * 9):
        st = _lib.load().e2f_conv3x3_tanh_nchw(src.hi.data_ptr(), src.lo.data_ptr(), src.hi.shape[-1], w_hi.data_ptr(),
                                               w_lo.data_ptr(), None if b32 is None else b32.data_ptr(), out.data_ptr(),
                                               n, h, w, cout, _stream())
    _lib.check(st, "e2f_conv3x3_tanh_nchw")
    return out


# ------------------------------------------------------------------------------------------------- SPyNet glue
class FlowPyramid:
    """The six normalised pyramid levels of every local frame (``spynet_pyramid``)."""

    __slots__ = ("levels", "b", "l_t", "size", "up")

    def __init__(self, levels, b, l_t, size, up):
        self.levels, self.b, self.l_t, self.size, self.up = levels, b, l_t, size, up


def spynet_pyramid(masked_frames, num_local_frames, mean, std):
    """Everything in front of SPyNet's first level, once per LOCAL FRAME (the reference does it per (ref, supp) pair):
    ``(x + 1) / 2`` and the 1/4 bilinear downsample of e2fgvi.py:210-218,247, the resize to multiples of 32, the
    normalisation and the five 2x2 average pools of flow_comp.py:95-115,152-158.  masked_frames (b,t,3,H,W) fp32;
    mean / std: the (1,3,1,1) buffers.  Returns a ``FlowPyramid``: levels[k] = (b*l_t, 3, h_up >> k, w_up >> k) fp32."""
    _need_cuda(masked_frames, mean, std)
    b, t, c, H, W = masked_frames.shape
    if c != 3:
        raise ValueError("spynet_pyramid: frames must have 3 channels")
    x = masked_frames.contiguous().float()
    l_t = int(num_local_frames)
    h, w = int(H * 0.25), int(W * 0.25)                 # F.interpolate(scale_factor=1/4, recompute_scale_factor=True)
    hu = h if h % 32 == 0 else 32 * (h // 32 + 1)
    wu = w if w % 32 == 0 else 32 * (w // 32 + 1)
    n = b * l_t
    sizes = [n * 3 * (hu >> k) * (wu >> k) for k in range(6)]
    buf = torch.empty(sum(sizes), dtype=torch.float32, device=x.device)
    m32, s32 = mean.detach().float().contiguous(), std.detach().float().contiguous()
    st = _lib.load().e2f_spynet_pyramid(x.data_ptr(), buf.data_ptr(), b, t, l_t, H, W, h, w, hu, wu, m32.data_ptr(),
                                        s32.data_ptr(), _stream())
    _lib.check(st, "e2f_spynet_pyramid")
    levels, off = [], 0
    for k in range(6):
        levels.append(buf[off: off + sizes[k]].view(n, 3, hu >> k, wu >> k))
        off += sizes[k]
    return FlowPyramid(levels, b, l_t, (h, w), (hu, wu))


def spynet_level_input(pyr, k, prev_flow, lead=3):
    """Input of one SPyNet level for all 2*b*(l_t-1) (ref, supp) pairs (forward pairs, then backward pairs): x2 flow
    upsample * 2, border warp of the support frame, ``cat([ref, warped, flow_up])`` (flow_comp.py:121-133) as the
    row-gapped ``RowsNHWC`` operand of the level's first 7x7 conv, plus ``flow_up`` (P, hk, wk, 2) fp32 — the residual the
    level's last conv adds.  pyr: ``FlowPyramid``; k: pyramid level (5 = coarsest); prev_flow (P, hk/2, wk/2, 2) or None."""
    img = pyr.levels[k]
    _, _, hk, wk = img.shape
    P = 2 * pyr.b * (pyr.l_t - 1)
    numel = _rows_numel(P, hk, wk, lead, 8)
    dev = img.device
    hi = torch.empty(numel, dtype=torch.bfloat16, device=dev)
    lo = torch.empty(numel, dtype=torch.bfloat16, device=dev)
    flow_up = torch.empty((P, hk, wk, 2), dtype=torch.float32, device=dev)
    if prev_flow is not None:
        if tuple(prev_flow.shape) != (P, hk // 2, wk // 2, 2) or not prev_flow.is_contiguous() or prev_flow.dtype != torch.float32:
            raise ValueError(f"spynet_level_input: prev_flow {tuple(prev_flow.shape)} != {(P, hk // 2, wk // 2, 2)} fp32 contiguous")
    st = _lib.load().e2f_spynet_level_input(img.data_ptr(), None if prev_flow is None else prev_flow.data_ptr(),
                                            hi.data_ptr(), lo.data_ptr(), flow_up.data_ptr(), pyr.b, pyr.l_t, hk, wk, lead,
                                            _stream())
    _lib.check(st, "e2f_spynet_level_input")
    return RowsNHWC(hi, lo, (P, 8, hk, wk), lead, 8), flow_up


def spynet_final(flow, pyr):
    """flow_comp.py:160-167 for both directions: level-0 flow (P, h_up, w_up, 2) fp32 -> (flows_forward, flows_backward),
    each (b, l_t-1, 2, h, w) fp32 (resize to (h, w), u * w / w_up, v * h / h_up)."""
    _need_cuda(flow)
    h, w = pyr.size
    hu, wu = pyr.up
    P = 2 * pyr.b * (pyr.l_t - 1)
    if tuple(flow.shape) != (P, hu, wu, 2) or not flow.is_contiguous() or flow.dtype != torch.float32:
        raise ValueError(f"spynet_final: flow {tuple(flow.shape)} != {(P, hu, wu, 2)} fp32 contiguous")
    fwd = torch.empty((pyr.b, pyr.l_t - 1, 2, h, w), dtype=torch.float32, device=flow.device)
    bwd = torch.empty((pyr.b, pyr.l_t - 1, 2, h, w), dtype=torch.float32, device=flow.device)
    st = _lib.load().e2f_spynet_final(flow.data_ptr(), fwd.data_ptr(), bwd.data_ptr(), pyr.b, pyr.l_t, h, w, hu, wu, _stream())
    _lib.check(st, "e2f_spynet_final")
    return fwd, bwd


def attention_flops(B, T, H, W, C, window_size, expand_size, focal_window, use_pooled=True):
    """Algorithmic FLOPs of one attention launch as the REFERENCE counts keys: QK^T + PV over
    T*(own window + listed ring keys incl. duplicates + fh*fw pooled keys incl. masked ones) keys per query."""
    wh, ww = window_size
    eh, ew = expand_size
    ring = 4 * (wh * ww - (wh - eh) * (ww - ew)) if (eh or ew) else 0
    keys = T * (wh * ww + ring + (focal_window[0] * focal_window[1] if use_pooled else 0))
    return 4.0 * B * T * H * W * keys * C


def invalidate_weight_caches():
    """Drop every operand derived from model parameters (bf16 splits, packed conv / gather-conv weights, folded bias
    maps).  The caches key on ``(param._version, data_ptr)``; in-place updates through ``param.data`` (``nn.init`` on
    ``.data``, EMA ``p.data.copy_()``, manual surgery) do NOT bump the version, so call this after such updates.
    ``InpaintGenerator.init_weights`` and ``load_state_dict`` do it for you."""
    _WEIGHT_SPLITS.clear()
    _CONV_PACKS.clear()
    _DERIVED.clear()


_GRAPH_REPLAY_LAUNCHES = 0


def note_graph_replay(kernel_launches):
    """A CUDA-graph replay relaunches the kernels recorded at capture time without passing through the C ABI's launch
    functions; ``e2fgvi_b200.graph`` reports them here so that ``launch_count`` keeps counting kernels, not host calls."""
    global _GRAPH_REPLAY_LAUNCHES
    _GRAPH_REPLAY_LAUNCHES += int(kernel_launches)


def launch_count():
    """Kernels of this library launched so far: host-side launches (``e2f_launch_count``) + kernels replayed by graphs."""
    return int(_lib.load().e2f_launch_count()) + _GRAPH_REPLAY_LAUNCHES
