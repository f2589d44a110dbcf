"""CPU tests of the host-side logic: weight packing, tile-box heuristics, module structure / state-dict
compatibility with the reference layout (Appendix B of SURVEY.md), timestep forms, router merge indexing."""
import math

import pytest
import torch

from ctrl_adapter_b200 import adapter as A
from ctrl_adapter_b200 import ops
from oracle import cases


def test_choose_box_covers_and_products():
    for (w, h, n) in [(128, 128, 16), (64, 64, 2), (8, 8, 3), (16, 9, 2), (128, 72, 56), (1, 1, 5), (4, 4, 1)]:
        bw, bh, bn = ops.choose_box(w, h, n)
        assert bw * bh * bn == 128
        tiles = math.ceil(w / bw) * math.ceil(h / bh) * math.ceil(n / bn)
        assert tiles * 128 >= w * h * n
    assert ops.choose_box(8, 8, 16) == (8, 8, 2)       # two 8x8 samples per tile, no padding
    assert ops.choose_box(128, 128, 1)[0] == 128


def test_pack_conv_weight_layout():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = ops.pack_conv_weight(w, 8).float()
    assert p.shape == (2, 9 * 8)
    # tap (dy, dx) = (1, 2) -> index 5; channel 2
    assert p[1, 5 * 8 + 2] == w[1, 2, 1, 2]
    assert float(p[:, 3:8].abs().max()) == 0.0  # zero padded channels
    w3 = torch.randn(4, 8, 3, 1, 1)
    p3 = ops.pack_conv_weight(w3).float()
    assert torch.equal(p3[:, 8:16], w3[:, :, 1, 0, 0].to(torch.bfloat16).float())


def test_pack_geglu_interleave():
    d, k = 512, 16
    w = torch.randn(2 * d, k)
    b = torch.randn(2 * d)
    wi, bi = ops.pack_geglu_weight(w, b, 256)
    # tile t holds value rows [128t, 128t+128) then gate rows [d+128t, d+128t+128)
    assert torch.equal(wi[256:384], w[128:256]) and torch.equal(wi[384:512], w[d + 128:d + 256])
    assert torch.equal(bi[0:128], b[0:128]) and torch.equal(bi[128:256], b[d:d + 128])


@pytest.mark.parametrize("kw,okind", [(cases.ADAPTER_SDXL_KW, "sdxl"), (cases.ADAPTER_VIDEO_KW, "video")])
def test_adapter_state_dict_matches_reference_layout(kw, okind):
    from oracle.adapter import ControlNetAdapter as O
    with torch.device("meta"):
        ours, ref = A.ControlNetAdapter(**kw), O(**kw)
    a, b = ours.state_dict(), ref.state_dict()
    assert set(a) == set(b)
    assert all(a[k].shape == b[k].shape for k in a)
    assert ours.get_down_block_ids() == ref.get_down_block_ids()
    assert ours.get_down_block_channels() == ref.get_down_block_channels()


def test_controlnet_and_unet_state_dicts_match_reference_layout():
    from ctrl_adapter_b200.controlnet import ControlNetModel
    from ctrl_adapter_b200.unet_sdxl import UNet2DConditionModel
    from oracle.controlnet import ControlNetModel as OCN
    from oracle.unet_sdxl import UNet2DConditionModel as OU
    with torch.device("meta"):
        for ours, ref in ((ControlNetModel(**cases.CONTROLNET_KW), OCN(**cases.CONTROLNET_KW)),
                          (UNet2DConditionModel(), OU())):
            a, b = ours.state_dict(), ref.state_dict()
            assert set(a) == set(b), (sorted(set(a) - set(b))[:5], sorted(set(b) - set(a))[:5])
            assert all(a[k].shape == b[k].shape for k in a)


def test_video_unet_state_dicts_match_reference_layout():
    """I2VGen-XL and SVD UNets: same keys / shapes as the (oracle restatement of the) reference classes, and the published
    parameter count of the SVD UNet (1 524 623 082 for stable-video-diffusion-img2vid)."""
    from ctrl_adapter_b200.unet_i2vgen import I2VGenXLUNet
    from ctrl_adapter_b200.unet_svd import UNetSpatioTemporalConditionModel
    from oracle.unet_i2vgen import I2VGenXLUNet as OI
    from oracle.unet_svd import UNetSpatioTemporalConditionModel as OS
    with torch.device("meta"):
        for ours, ref in ((I2VGenXLUNet(), OI()),
                          (UNetSpatioTemporalConditionModel(**cases.UNET_SVD_KW), OS(**cases.UNET_SVD_KW))):
            a, b = ours.state_dict(), ref.state_dict()
            assert set(a) == set(b), (sorted(set(a) - set(b))[:5], sorted(set(b) - set(a))[:5])
            assert all(a[k].shape == b[k].shape for k in a), [k for k in a if a[k].shape != b[k].shape][:5]
        n_svd = sum(p.numel() for p in ref.parameters())
    assert n_svd == 1_524_623_082


def test_svd_unet_rejects_unsupported_topologies():
    from ctrl_adapter_b200.unet_svd import UNetSpatioTemporalConditionModel
    with torch.device("meta"):
        with pytest.raises(ValueError):
            UNetSpatioTemporalConditionModel(down_block_types=("DownBlockSpatioTemporal",) * 3)
        with pytest.raises(NotImplementedError):
            UNetSpatioTemporalConditionModel(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4))
        with pytest.raises(NotImplementedError):
            UNetSpatioTemporalConditionModel()  # class default heads (5, 10, 10, 20): 128-wide heads in stage 3


def test_router_state_dict_and_config():
    from oracle.adapter import ControlNetRouter as O
    ours, ref = A.ControlNetRouter(**cases.ROUTER_KW), O(**cases.ROUTER_KW)
    assert set(ours.state_dict()) == set(ref.state_dict())
    assert ours.router_type == "simple_weights" and ours.num_routers == 12 and ours.num_experts == 7
    with pytest.raises(ValueError):
        A.ControlNetRouter(router_type="sparsemax")


def test_timestep_forms():
    dev = torch.device("cpu")
    for t in (981, 981.0, torch.tensor(981.0), torch.tensor([981.0]), torch.tensor([[981.0], [981.0], [981.0]])):
        v = A.timestep_vector(t, 3, dev)
        assert v.shape == (3,) and v.dtype == torch.float32 and float(v[0]) == 981.0
    with pytest.raises(TypeError):
        A.timestep_vector("981", 3, dev)


def test_adapter_rejects_unsupported_configs():
    with pytest.raises(NotImplementedError):
        A.ControlNetAdapter("sdxl", num_repeats=2, add_adapter_location_A=True)
    with pytest.raises(ValueError):
        A.AdapterSpatioTemporal(320, 640)


def test_adapter_save_load_roundtrip(tmp_path):
    kw = dict(cases.ADAPTER_SDXL_KW, add_adapter_location_B=False, add_adapter_location_C=False,
              num_adapters_per_location=1)
    m = A.ControlNetAdapter(**kw)
    m.save_pretrained(str(tmp_path / "adapter"))
    m2 = A.ControlNetAdapter.from_pretrained(str(tmp_path), subfolder="adapter")
    assert m2.config == m.config
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
