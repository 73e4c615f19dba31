"""CPU: the drop-in boundary and the model's host logic (kernels replaced by the oracle, in tests only)."""
import importlib
import json
import os

import pytest
import torch

from e2fgvi_b200.synth import synth_frames, synth_state_dict
from oracle_backend import oracle_ops

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["e2fgvi", "e2fgvi_hq"])
def test_state_dict_layout_matches_reference(name):
    """Keys, order, shapes and dtypes equal the reference generator's (strict checkpoint compatibility)."""
    layout = json.load(open(os.path.join(GOLDEN, "state_dict_layout.json")))[name]
    net = importlib.import_module("model." + name)          # the reference's import path (test.py:117)
    sd = net.InpaintGenerator().state_dict()
    assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == layout


def test_constructor_is_offline_and_default_init():
    net = importlib.import_module("model.e2fgvi")
    g = net.InpaintGenerator()
    for m in g.feat_prop_module.deform_align.values():       # init_offset (feat_prop.py:32-33)
        assert float(m.conv_offset[-1].weight.abs().max()) == 0.0
    assert abs(float(g.encoder.layers[0].weight.std()) - 0.02) < 0.005
    assert g.transformer[0].attn.valid_ind_rolled.numel() == 120


@pytest.mark.parametrize("name", ["e2e_hq_tiny_stress"])
def test_host_logic_against_golden(name):
    g = torch.load(os.path.join(GOLDEN, name + ".pt"))
    c = g["case"]
    net = importlib.import_module("model." + ("e2fgvi_hq" if c["hq"] else "e2fgvi"))
    model = net.InpaintGenerator().eval()
    model.load_state_dict(synth_state_dict(model, c["family"], c["weight_seed"]), strict=True)
    x = synth_frames(1, c["T"], c["H"], c["W"], seed=c["frame_seed"])
    with torch.no_grad(), oracle_ops():
        pred, (ff, fb) = model(x, c["l_t"])
    assert pred.shape == (c["T"], 3, c["H"], c["W"])
    assert (pred - g["pred"]).abs().max() < 5e-5
    assert (ff - g["flows_forward"]).abs().max() < 1e-3 and (fb - g["flows_backward"]).abs().max() < 1e-3


@pytest.mark.parametrize("hq,b,T,l_t,H,W", [(True, 1, 4, 4, 60, 108), (True, 2, 3, 2, 60, 108)])
def test_host_logic_edge_frame_counts(hq, b, T, l_t, H, W):
    """No reference frames (T == l_t: ``get_ref_index`` returned [], test.py:37-52) and the shortest local window: the
    model's host logic (buffer slicing, in-place propagation, flow indexing) against the oracle's restatement."""
    from oracle import restate
    net = importlib.import_module("model." + ("e2fgvi_hq" if hq else "e2fgvi"))
    model = net.InpaintGenerator().eval()
    sd = synth_state_dict(model, "stress", 5)
    model.load_state_dict(sd, strict=True)
    x = synth_frames(b, T, H, W, seed=9)
    with torch.no_grad():
        want, (wf, wb) = restate.inpaint_generator_forward(sd, x, l_t, hq=hq)
        with oracle_ops():
            pred, (ff, fb) = model(x, l_t)
    assert pred.shape == (b * T, 3, H, W) and ff.shape == (b, l_t - 1, 2, H // 4, W // 4)
    assert (pred - want).abs().max() < 5e-5
    assert (ff - wf).abs().max() < 1e-3 and (fb - wb).abs().max() < 1e-3


def test_reference_module_boundaries():
    """The inner operator boundaries keep the reference's call shapes."""
    from model.modules.tfocal_transformer import WindowAttention
    from model.modules.feat_prop import SecondOrderDeformableAlignment
    attn = WindowAttention(512, (2, 4), (5, 9), (5, 9), 2, 4, True, "fc").eval()
    x = torch.randn(1, 2, 10, 18, 512)
    pooled = torch.randn(1, 2, 2, 2, 512)
    with torch.no_grad(), oracle_ops():
        out = attn([x, pooled], [None, None])
    assert out.shape == (4, 90, 512)
    align = SecondOrderDeformableAlignment(256, 128, 3, padding=1, deform_groups=16).eval()
    with torch.no_grad(), oracle_ops():
        y = align(torch.randn(1, 256, 6, 8), torch.randn(1, 384, 6, 8), torch.randn(1, 2, 6, 8), torch.randn(1, 2, 6, 8))
        align.fused = False
        y2 = align(torch.randn(1, 256, 6, 8), torch.randn(1, 384, 6, 8), torch.randn(1, 2, 6, 8), torch.randn(1, 2, 6, 8))
    assert y.shape == y2.shape == (1, 128, 6, 8)


def test_forward_fails_loudly_without_gpu():
    net = importlib.import_module("model.e2fgvi_hq")
    model = net.InpaintGenerator().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        model(synth_frames(1, 3, 60, 108), 2)


def test_merge_conv_groups_is_equivalent_grouped_conv():
    """Block-diagonal group merging (ops.merge_conv_groups) leaves the grouped conv unchanged: checked against
    F.conv2d on the group-wise concatenation the encoder performs (reference e2fgvi.py:103-108)."""
    import torch.nn.functional as F
    from e2fgvi_b200 import ops
    torch.manual_seed(3)
    groups, c1, c2, cout = 8, 32, 48, 32 * 8 // 8 * 2          # 8 groups, 4+6 inputs and 8 outputs per group
    x1, x2 = torch.randn(2, c1, 6, 7), torch.randn(2, c2, 6, 7)
    w = torch.randn(cout, (c1 + c2) // groups, 3, 3)

    def group_cat(g):
        a = x1.view(2, g, -1, 6, 7)
        b = x2.view(2, g, -1, 6, 7)
        return torch.cat([a, b], 2).view(2, -1, 6, 7)

    want = F.conv2d(group_cat(groups), w, None, 1, 1, 1, groups)
    merged, g2 = ops.merge_conv_groups(w, [c1, c2], groups)
    assert g2 < groups and merged.shape[1] == w.shape[1] * (groups // g2)
    got = F.conv2d(group_cat(g2), merged, None, 1, 1, 1, g2)
    assert torch.allclose(got, want, atol=1e-5)
    same, g3 = ops.merge_conv_groups(torch.randn(256, 16, 3, 3), [128], 8)   # 32 outputs per group: merge by 2
    assert g3 == 4 and same.shape == (256, 32, 3, 3)
    keep, g4 = ops.merge_conv_groups(torch.randn(384, 48, 3, 3), [192], 4)   # 96 per group: untouched
    assert g4 == 4 and keep.shape == (384, 48, 3, 3)
