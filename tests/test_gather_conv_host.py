"""CPU: the tap / phase tables and weight packing that ops.soft_split / ops.soft_comp hand to e2f_conv_gather_bf16x3,
checked by executing the C-ABI's documented semantics (include/e2fgvi_b200.h) in plain torch and comparing with the
reference formulation (unfold + Linear, Linear + fold; tfocal_transformer.py:39-46, 65-72).  The CUDA kernel itself
is compared with the same formulations in tests/test_gpu_ops.py."""
import contextlib

import pytest
import torch
import torch.nn.functional as F

from e2fgvi_b200 import ops


def _split_cpu(x):
    x = x.contiguous().float()
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def _gather_ref(sources, w_hi, w_lo, bias, bias_map, residual, out, cout, stride, grid, taps, phases, ostep,
                out_size, flops, slope=1.0, into=None):
    """e2f_conv_gather_bf16x3 semantics, fp64 on the CPU (sources concatenated along channels, 64-ch chunk padding)."""
    assert len(sources) == 1 and into is None and slope == 1.0
    src = sources[0]
    n, _, h_in, w_in = src.shape
    x = (src.hi.double() + src.lo.double())                      # (n, h_in, w_in, C)
    C = x.shape[-1]
    wt = (w_hi.double() + w_lo.double()).view(cout, len(taps), -1)[:, :, :C]    # K per tap is padded to 64-ch chunks
    gh, gw = grid
    oh, ow = out_size
    y = torch.zeros((n, oh, ow, cout), dtype=torch.float64)
    bounds = [p[0] for p in phases] + [len(taps)]
    gy = torch.arange(gh).view(gh, 1)
    gx = torch.arange(gw).view(1, gw)
    for ph, (t0, oy, ox) in enumerate(phases):
        acc = torch.zeros((n, gh, gw, cout), dtype=torch.float64)
        for t in range(t0, bounds[ph + 1]):
            iy, ix = gy * stride + taps[t][0], gx * stride + taps[t][1]
            ok = ((iy >= 0) & (iy < h_in) & (ix >= 0) & (ix < w_in)).double().view(1, gh, gw, 1)
            g = x[:, iy.clamp(0, h_in - 1).expand(gh, gw), ix.clamp(0, w_in - 1).expand(gh, gw)] * ok
            acc += g @ wt[:, t].T
        Y, X = gy * ostep + oy, gx * ostep + ox
        keep = ((Y < oh) & (X < ow))
        ys, xs = Y.expand(gh, gw)[keep], X.expand(gh, gw)[keep]
        y[:, ys, xs] = acc[:, keep]
    if bias is not None:
        y = y + bias.double()
    if bias_map is not None:
        y = y + bias_map.double()
    if residual is not None:
        y = y + residual.permute(0, 2, 3, 1).double()
    y = y.float()
    t32 = y.permute(0, 3, 1, 2)
    sp = ops.SplitNHWC(*_split_cpu(y), (n, cout, oh, ow))
    return t32 if out == "f32" else sp if out == "split" else (t32, sp)


@contextlib.contextmanager
def cpu_gather():
    saved = {k: getattr(ops, k) for k in ("_conv_gather", "split_bf16", "_need_cuda", "split_nhwc")}
    ops._conv_gather, ops.split_bf16, ops._need_cuda = _gather_ref, _split_cpu, lambda *a: None
    ops.split_nhwc = lambda x: x if isinstance(x, ops.SplitNHWC) else ops.SplitNHWC(
        *_split_cpu(x.permute(0, 2, 3, 1)), tuple(x.shape))
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)


@pytest.mark.parametrize("hw", [(12, 18), (15, 27), (10, 13)])
def test_soft_split_is_a_strided_conv(hw):
    h, w = hw
    g = torch.Generator().manual_seed(3)
    c, hidden = 16, 24
    x = torch.randn(2, c, h, w, generator=g)
    lin = torch.nn.Linear(c * 49, hidden)
    want = lin(F.unfold(x, 7, padding=3, stride=3).permute(0, 2, 1))
    with cpu_gather(), torch.no_grad():
        got = ops.soft_split(x, lin.weight, lin.bias, 7, 3, 3)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 2e-4 * want.abs().max().item()      # bf16x2 operand pairs: ~2^-16


@pytest.mark.parametrize("hw", [(12, 18), (15, 27), (10, 13)])
@pytest.mark.parametrize("out", ["f32", "split"])
def test_soft_comp_is_a_nine_phase_transposed_conv(hw, out):
    h, w = hw
    fh, fw = (h - 1) // 3 + 1, (w - 1) // 3 + 1
    g = torch.Generator().manual_seed(4)
    c, hidden = 8, 64
    tok = torch.randn(2, fh, fw, hidden, generator=g)
    lin = torch.nn.Linear(hidden, c * 49)
    extra = torch.nn.Parameter(torch.randn(c, h, w, generator=g))
    res = torch.randn(2, c, h, w, generator=g)
    want = F.fold(lin(tok.view(2, fh * fw, hidden)).permute(0, 2, 1), (h, w), 7, padding=3, stride=3) + extra + res
    with cpu_gather(), torch.no_grad():
        got = ops.soft_comp(tok, lin.weight, lin.bias, (h, w), 7, 3, 3, bias_map_extra=extra, residual=res, out=out)
    if out == "split":
        got = (got.hi.float() + got.lo.float()).permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 2e-4 * want.abs().max().item()


def test_soft_comp_tables_cover_every_kernel_position_once():
    taps, phases, kpos = ops._soft_comp_tables(7, 3, 3)
    assert len(taps) == 49 and sorted(kpos) == [(a, b) for a in range(7) for b in range(7)]
    assert [p[0] for p in phases] == [0, 9, 15, 21, 27, 31, 35, 41, 45]
    with pytest.raises(NotImplementedError):
        ops._soft_comp_tables(7, 3, 2)


def test_best_tile_fills_the_token_grids():
    for (gh, gw), tiles in (((20, 36), 6), ((60, 108), 54), ((90, 162), 117)):
        tw, th = ops._best_tile(gh, gw, 3)
        assert tw * th <= 128 and -(-gh // th) * -(-gw // tw) == tiles


@pytest.mark.parametrize("case", [(64, 32, 7), (32, 16, 7), (16, 2, 7), (64, 3, 3), (72, 5, 3)])
def test_kxn_weight_packing_and_shift_sum(case):
    """e2f_conv_kxn_bf16x3's documented two-step semantics executed in torch with the packed weight of
    ops.pack_conv_kxn_weight: D[(y, xin), (kx, co)] = sum_{ky, c} X[y+ky-pad, xin, c] W[co, c, ky, kx], then
    out[y, x, co] = sum_kx D[(y, x+kx-pad), (kx, co)]  ==  F.conv2d(x, w, padding=k//2)."""
    cin, cout, ks = case
    pad = ks // 2
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, cin, 9, 13, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5
    co_pad = ops._kxn_co_pad(cout, ks)
    assert co_pad is not None and (ks * co_pad) % 16 == 0
    saved = ops.split_bf16
    ops.split_bf16 = _split_cpu
    try:
        hi, lo = ops.pack_conv_kxn_weight(w, co_pad)
    finally:
        ops.split_bf16 = saved
    chunks = (cin + 63) // 64
    wp = (hi.double() + lo.double()).view(ks, co_pad, ks, chunks * 64)[:, :, :, :cin]         # [kx][co][ky][c]
    xp = F.pad(x.double(), (pad, pad, pad, pad))                                               # zero padding = TMA OOB fill
    n, _, h, wd = x.shape
    # D over the padded column range xin in [-pad, W + pad): D[n, y, xin, kx, co]
    D = torch.zeros(n, h, wd + 2 * pad, ks, co_pad, dtype=torch.float64)
    for ky in range(ks):
        rows = xp[:, :, ky:ky + h, :]                                                           # X[y + ky - pad, xin]
        D += torch.einsum("ncyx,koc->nyxko", rows, wp[:, :, ky, :])
    out = torch.zeros(n, h, wd, co_pad, dtype=torch.float64)
    for kx in range(ks):
        out += D[:, :, kx:kx + wd, kx, :]                                                       # D[(y, x + kx - pad)] in padded coords
    want = F.conv2d(x.double(), w.double(), padding=pad)
    got = out[..., :cout].permute(0, 3, 1, 2)
    assert (got - want).abs().max().item() < 2e-4 * want.abs().max().item()
    assert float(out[..., cout:].abs().max()) == 0.0 if co_pad > cout else True
