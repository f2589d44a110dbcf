"""CPU tests: the oracle (oracle/*.py) reproduces the golden vectors that tests/golden/make_golden.py produced by
running the reference's own classes (fp32, same name-seeded weights and inputs).  Tolerance: fp32 vs fp32 of the same
arithmetic -> rtol 1e-5 / atol 1e-6 (only kernel-selection noise of the CPU BLAS)."""
import json
import os

import pytest
import torch

from oracle import cases
from oracle.adapter import ControlNetAdapter, ControlNetRouter
from oracle.blocks import ResnetBlock2D
from oracle.controlnet import ControlNetModel, MultiControlNetModel
from oracle.weights import fingerprint, seeded_init_, seeded_tensor

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))
torch.set_grad_enabled(False)


def assert_fp(t, gold, rtol=1e-5, atol=1e-6):
    got = fingerprint(t, len(gold["samples"]))
    assert got["shape"] == gold["shape"]
    torch.testing.assert_close(torch.tensor(got["samples"]), torch.tensor(gold["samples"]), rtol=rtol, atol=atol)
    assert abs(got["mean"] - gold["mean"]) <= atol + rtol * abs(gold["mean"]) + 1e-5 * gold["absmax"]
    assert abs(got["absmax"] - gold["absmax"]) <= atol + rtol * gold["absmax"]


def test_adapter_sdxl_matches_reference():
    m = seeded_init_(ControlNetAdapter(**cases.ADAPTER_SDXL_KW), seed=1).eval()
    down, mid = m(**cases.adapter_sdxl_inputs())
    assert mid is None and GOLD["adapter_sdxl"]["mid"] is None
    assert len(down) == 12
    for t, g in zip(down, GOLD["adapter_sdxl"]["down"]):
        assert_fp(t, g)
    # blocks 9..11 are not selected for SDXL: new zero tensors of the input shape (ctrl_adapter.py:193)
    assert all(float(t.abs().max()) == 0.0 for t in down[9:])


def test_adapter_video_matches_reference():
    m = seeded_init_(ControlNetAdapter(**cases.ADAPTER_VIDEO_KW), seed=2).eval()
    down, mid = m(**cases.adapter_video_inputs())
    for t, g in zip(down, GOLD["adapter_video"]["down"]):
        assert_fp(t, g)
    assert_fp(mid, GOLD["adapter_video"]["mid"])


def test_router_matches_reference():
    r = seeded_init_(ControlNetRouter(**cases.ROUTER_KW), seed=3).eval()
    dw, mw = r(sparse_mask=cases.ROUTER_MASK)
    assert_fp(dw, GOLD["router"]["down"])
    assert_fp(mw, GOLD["router"]["mid"])
    assert dw.shape == (12, 7) and mw.shape == (7,)
    # masked experts get (numerically) zero weight
    assert float(dw[:, [2, 4, 5, 6]].max()) == 0.0
    dw, mw = r(sparse_mask=None)
    assert_fp(dw, GOLD["router_nomask"]["down"])
    assert_fp(mw, GOLD["router_nomask"]["mid"])


def test_controlnet_matches_reference():
    cn = seeded_init_(ControlNetModel(**cases.CONTROLNET_KW), seed=4).eval()
    inp = cases.controlnet_inputs()
    down, mid = cn(**inp)
    assert len(down) == 12
    for t, g in zip(down, GOLD["controlnet"]["down"]):
        assert_fp(t, g, rtol=2e-5, atol=2e-6)
    assert_fp(mid, GOLD["controlnet"]["mid"], rtol=2e-5, atol=2e-6)
    down, mid = cn(**{**inp, "skip_conv_in": True, "conditioning_scale": 0.75})
    for t, g in zip(down, GOLD["controlnet_skip_conv_in"]["down"]):
        assert_fp(t, g, rtol=2e-5, atol=2e-6)
    # MultiControlNet: zip() truncation to the number of provided images, list outputs (multicontrolnet.py:66-99)
    multi = MultiControlNetModel([cn, cn, cn])
    conds = [inp["controlnet_cond"], torch.flip(inp["controlnet_cond"], dims=[3])]
    dl, ml = multi(inp["sample"], inp["timestep"], inp["encoder_hidden_states"], conds, [1.0, 0.5, 0.25], return_dict=False)
    assert len(dl) == GOLD["multicontrolnet"]["n_nets_run"] == 2
    for t, g in zip(dl[1], GOLD["multicontrolnet"]["down1"]):
        assert_fp(t, g, rtol=2e-5, atol=2e-6)
    assert_fp(ml[1], GOLD["multicontrolnet"]["mid1"], rtol=2e-5, atol=2e-6)


def test_resnet_upsample_output_size_matches_reference():
    rb = seeded_init_(ResnetBlock2D(in_channels=320, out_channels=320, temb_channels=320, eps=1e-6,
                                    use_in_shortcut=True, up=True), seed=5).eval()
    x = seeded_tensor("rb_x", (2, 320, 6, 5), 5)
    temb = seeded_tensor("rb_temb", (2, 320), 5)
    assert_fp(rb(x, temb, output_size=(12, 10)), GOLD["resnet_up"])
    assert_fp(rb(x, temb, output_size=(9, 8)), GOLD["resnet_up_odd"])


# ---- self-consistency checks that stand in for the missing upstream KATs ----
def test_timesteps_closed_form():
    from oracle.blocks import Timesteps
    import math
    t = torch.tensor([0.0, 1.0, 500.0, 999.0])
    emb = Timesteps(320, True, 0)(t)
    k = torch.arange(160, dtype=torch.float64)
    freq = torch.exp(-math.log(10000.0) * k / 160)
    arg = t.double()[:, None] * freq[None]
    ref = torch.cat([torch.cos(arg), torch.sin(arg)], -1)
    torch.testing.assert_close(emb.double(), ref, rtol=0, atol=2e-4)


def test_attention_matches_explicit_softmax_fp64():
    from oracle.blocks import Attention
    a = seeded_init_(Attention(query_dim=64, heads=2, dim_head=32), 9).double()
    x = seeded_tensor("att_x", (2, 10, 64), 9).double()
    q, k, v = a.to_q(x), a.to_k(x), a.to_v(x)
    q, k, v = (t.view(2, 10, 2, 32).transpose(1, 2) for t in (q, k, v))
    p = torch.softmax(q @ k.transpose(-1, -2) / 32 ** 0.5, -1)
    ref = a.to_out[0]((p @ v).transpose(1, 2).reshape(2, 10, 64))
    torch.testing.assert_close(a(x), ref, rtol=1e-10, atol=1e-12)


def test_sdxl_unet_param_count_and_adapter_wiring():
    """2 567 463 684 parameters = the published SDXL-base UNet size; 361 279 120 = SD1.5 ControlNet."""
    from oracle.unet_sdxl import UNet2DConditionModel
    with torch.device("meta"):
        u = UNet2DConditionModel()
        c = ControlNetModel(**cases.CONTROLNET_KW)
    assert sum(p.numel() for p in u.parameters()) == 2567463684
    assert sum(p.numel() for p in c.parameters()) == 361279120


def test_unet_svd_matches_reference():
    """Reduced-width SVD UNet (same block types / depths as the released model) incl. the reference's 5-D residual
    injection with zip truncation and the mid residual (svd/.../unet_spatio_temporal_condition.py:457-471, 485-490)."""
    from oracle.unet_svd import UNetSpatioTemporalConditionModel
    u = seeded_init_(UNetSpatioTemporalConditionModel(**cases.UNET_SVD_SMALL_KW), seed=11).eval()
    assert sum(p.numel() for p in u.parameters()) == GOLD["unet_svd_small"]["n_params"]
    assert_fp(u(**cases.unet_svd_inputs(with_residuals=True))[0], GOLD["unet_svd_small"]["with_residuals"], rtol=2e-5, atol=2e-6)
    assert_fp(u(**cases.unet_svd_inputs(with_residuals=False))[0], GOLD["unet_svd_small"]["plain"], rtol=2e-5, atol=2e-6)


def test_unet_i2vgen_matches_reference():
    """Reduced-width I2VGen-XL UNet incl. the residual injection (i2vgen_xl/.../unet_i2vgen_xl.py:681-695, 709-714)."""
    from oracle.unet_i2vgen import I2VGenXLUNet
    u = seeded_init_(I2VGenXLUNet(**cases.UNET_I2VGEN_SMALL_KW), seed=12).eval()
    assert sum(p.numel() for p in u.parameters()) == GOLD["unet_i2vgen_small"]["n_params"]
    assert_fp(u(**cases.unet_i2vgen_small_inputs(with_residuals=True))[0], GOLD["unet_i2vgen_small"]["with_residuals"],
              rtol=2e-5, atol=2e-6)
    assert_fp(u(**cases.unet_i2vgen_small_inputs(with_residuals=False))[0], GOLD["unet_i2vgen_small"]["plain"],
              rtol=2e-5, atol=2e-6)
