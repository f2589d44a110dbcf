"""B200 host mirror of the decoder half of diffusers' ``AutoencoderKL`` -- the VAE ``decode`` the reference pipelines
call after the denoising loop (/root/reference/sdxl/pipelines/sdxl_controlnet_adapter_pipeline.py:1414;
i2vgen_xl/pipelines/i2vgen_xl_controlnet_adapter_pipeline.py:398-418 ``decode_latents``).

Same constructor kwargs, ``decode()`` call form and state-dict keys as the diffusers class (``post_quant_conv.*``,
``decoder.conv_in / mid_block.{resnets,attentions} / up_blocks.N.{resnets,upsamplers} / conv_norm_out / conv_out``); the
encoder half (``encoder.*``, ``quant_conv.*``) is not part of the path and its keys are skipped at load time.  Everything
runs on the kernels of the denoising path: 3x3 / 1x1 convolutions on the implicit-GEMM kernel, GroupNorm(+SiLU) kernels,
the 2x nearest up-sampler; the mid block's single 512-wide attention head (outside ``attention_kernel``'s head dims, and
run once per generation) as QK^T GEMM (fp32) -> ``ca_softmax_rows`` -> PV GEMM, one image at a time.

``AutoencoderKLTemporalDecoder`` (decoder half) is the VAE of the SVD pipeline (svd/pipelines/
svd_controlnet_adapter_pipeline.py:265-292 ``decode_latents``): every decoder ResNet is a SpatioTemporalResBlock (spatial
ResnetBlock2D -> TemporalResnetBlock over the frames of the decode chunk, blended by a learned, switched alpha in the
GEMM epilogue), and a Conv3d (3,1,1) over the decoded RGB frames (``ca_frame_conv_small``) closes it.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from .adapter import AlphaBlender, TemporalResnetBlock, _ConfigDict, as_nchw, to_channels_last_bf16
from .layers import BF16, Conv2d, Linear, Norm, Packable, ResnetBlock2D
from .persistence import PretrainedMixin


class _VaeAttention(nn.Module):
    """diffusers ``Attention(heads=1, dim_head=C, norm_num_groups=32, residual_connection=True, bias=True)`` as built by
    UNetMidBlock2D for the VAE: GroupNorm -> q / k / v Linear (with bias) -> softmax(q k^T / sqrt(C)) v -> to_out + x."""

    def __init__(self, c: int, eps: float):
        super().__init__()
        self.group_norm = Norm(c, eps)
        self.to_q, self.to_k, self.to_v = Linear(c, c), Linear(c, c), Linear(c, c)
        self.to_out = nn.ModuleList([Linear(c, c), nn.Dropout(0.0)])
        self.scale = c ** -0.5

    def forward(self, x):
        n, h, w, c = x.shape
        hw = h * w
        t = self.group_norm.group_norm(x, silu=False).reshape(n * hw, c)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = torch.empty_like(q)
        for i in range(n):  # one image at a time: the [hw, hw] fp32 score matrix is 1 GB at 128 x 128 latents
            sl = slice(i * hw, (i + 1) * hw)
            s = ops.linear(q[sl], k[sl], None, out_fp32=True, out_scale=self.scale)          # [hw, hw] = q k^T * scale
            p = ops.softmax_rows(s)
            vt = ops.nhwc_to_nchw(v[sl].reshape(1, h, w, c)).reshape(c, hw)                   # v^T as the GEMM's B rows
            ops.linear(p, vt, None, out=o[sl])
        w_o, b_o = self.to_out[0].packed()
        return ops.linear(o, w_o, b_o, residual=x.reshape(n * hw, c)).reshape(n, h, w, c)


class _MidBlock(nn.Module):
    def __init__(self, c: int, eps: float):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, eps), ResnetBlock2D(c, c, None, eps)])
        self.attentions = nn.ModuleList([_VaeAttention(c, eps)])

    def forward(self, x):
        x = self.resnets[0](x, None)
        x = self.attentions[0](x)
        return self.resnets[1](x, None)


class _UpDecoderBlock(nn.Module):
    def __init__(self, cin: int, cout: int, layers: int, eps: float, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, eps) for i in range(layers)])
        self.upsamplers = None
        if add_upsample:
            us = nn.Module()
            us.conv = Conv2d(cout, cout, 3)
            self.upsamplers = nn.ModuleList([us])

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(ops.upsample2x(x))
        return x


class Decoder(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, block_out_channels, layers_per_block: int, eps: float):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = Conv2d(in_channels, rev[0], 3)
        self.mid_block = _MidBlock(rev[0], eps)
        blocks, prev = [], rev[0]
        for i, c in enumerate(rev):
            blocks.append(_UpDecoderBlock(prev, c, layers_per_block + 1, eps, i != len(rev) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = Norm(rev[-1], eps)
        self.conv_out = Conv2d(rev[-1], out_channels, 3)
        self.out_channels = out_channels
        self._conv_out_pad = None

    def _conv_out_packed(self):
        """conv_out has 3 output channels; the kernel stores 16-byte vectors, so its weight rows are zero padded to 8."""
        key = self.conv_out._key()
        if self._conv_out_pad is None or self._conv_out_pad[0] != key:
            w, b = self.conv_out.packed()
            wp = torch.zeros((8, w.shape[1]), device=w.device, dtype=w.dtype)
            wp[: w.shape[0]] = w
            bp = torch.zeros(8, device=w.device, dtype=torch.float32)
            bp[: b.shape[0]] = b
            self._conv_out_pad = (key, wp, bp)
        return self._conv_out_pad[1], self._conv_out_pad[2]

    def forward(self, z_nhwc8):
        x = self.conv_in(z_nhwc8)
        x = self.mid_block(x)
        for blk in self.up_blocks:
            x = blk(x)
        x = self.conv_norm_out.group_norm(x, silu=True)
        w, b = self._conv_out_packed()
        return ops.conv2d(x, w, b, ksize=3)  # [n, H, W, 8], the first out_channels are real


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class _DecoderOnlyVae(PretrainedMixin, nn.Module):
    ignore_prefixes = ("encoder.", "quant_conv.")  # the encoder half of a published checkpoint is not on the path

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.startswith(self.ignore_prefixes)}
        return super().load_state_dict(sd, strict=strict, **kw)


class AutoencoderKL(_DecoderOnlyVae):

    def __init__(self, in_channels: int = 3, out_channels: int = 3, down_block_types=("DownEncoderBlock2D",) * 4,
                 up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block: int = 2, act_fn: str = "silu", latent_channels: int = 4, norm_num_groups: int = 32,
                 sample_size: int = 1024, scaling_factor: float = 0.13025, force_upcast: bool = True, **_ignored):
        super().__init__()
        if act_fn != "silu" or norm_num_groups != 32 or any(t != "UpDecoderBlock2D" for t in up_block_types):
            raise NotImplementedError("only the SD / SDXL AutoencoderKL decoder topology is implemented")
        self.config = _ConfigDict(in_channels=in_channels, out_channels=out_channels,
                                  down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                                  block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                  act_fn=act_fn, latent_channels=latent_channels, norm_num_groups=norm_num_groups,
                                  sample_size=sample_size, scaling_factor=scaling_factor, force_upcast=force_upcast)
        self.post_quant_conv = Conv2d(latent_channels, latent_channels, 1)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, 1e-6)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """z [N, 4, h, w] (already divided by scaling_factor by the caller, as the pipelines do) -> images [N, 3, 8h, 8w]
        in [-1, 1] (logical NCHW, bf16)."""
        x = to_channels_last_bf16(z.to(BF16) if z.dtype not in (BF16, torch.float32) else z, 8)  # 4 -> 8 zero-padded channels
        w, b = self.post_quant_conv.packed()
        # 1x1 conv 4 -> 4 on the padded layout: output rows padded to 8 as well
        key = self.post_quant_conv._key()
        if getattr(self, "_pq_pad", None) is None or self._pq_pad[0] != key:
            wp = torch.zeros((8, w.shape[1]), device=w.device, dtype=w.dtype)
            wp[: w.shape[0]] = w
            bp = torch.zeros(8, device=w.device, dtype=torch.float32)
            bp[: b.shape[0]] = b
            self._pq_pad = (key, wp, bp)
        x = ops.conv2d(x, self._pq_pad[1], self._pq_pad[2], ksize=1)
        y = self.decoder(x)
        img = ops.nhwc_to_nchw(y, self.config.out_channels)
        return DecoderOutput(img) if return_dict else (img,)


# ---- SVD: AutoencoderKLTemporalDecoder (diffusers models/autoencoders/autoencoder_kl_temporal_decoder.py) ---------------
class _SpatioTemporalResBlock(nn.Module):
    """diffusers SpatioTemporalResBlock as the temporal decoder builds it: temb_channels=None, eps=1e-6,
    temporal_eps=1e-5, merge_strategy="learned", merge_factor=0.0, switch_spatial_to_temporal_mix=True."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(cin, cout, None, 1e-6)
        self.temporal_res_block = TemporalResnetBlock(cout, None, 1e-5)
        self.time_mixer = AlphaBlender(0.0, switch_spatial_to_temporal_mix=True)

    def forward(self, x, frames: int):
        s = self.spatial_res_block(x, None)
        return self.temporal_res_block(s, frames, None, blend_src=s, blend_alpha=self.time_mixer.alpha())


class _MidBlockTemporalDecoder(nn.Module):
    def __init__(self, c: int, layers: int):
        super().__init__()
        self.resnets = nn.ModuleList([_SpatioTemporalResBlock(c, c) for _ in range(layers)])
        self.attentions = nn.ModuleList([_VaeAttention(c, 1e-6)])

    def forward(self, x, frames: int):
        x = self.resnets[0](x, frames)
        for resnet, attn in zip(self.resnets[1:], self.attentions):
            x = resnet(attn(x), frames)
        return x


class _UpBlockTemporalDecoder(nn.Module):
    def __init__(self, cin: int, cout: int, layers: int, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([_SpatioTemporalResBlock(cin if i == 0 else cout, cout) for i in range(layers)])
        self.upsamplers = None
        if add_upsample:
            us = nn.Module()
            us.conv = Conv2d(cout, cout, 3)
            self.upsamplers = nn.ModuleList([us])

    def forward(self, x, frames: int):
        for r in self.resnets:
            x = r(x, frames)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(ops.upsample2x(x))
        return x


class _FrameConvSmall(Packable):
    """Conv3d(C, C, (3,1,1), padding (1,0,0)) on a handful of channels; its parameters travel to the kernel by value."""

    def __init__(self, c: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c, c, 3, 1, 1))
        self.bias = nn.Parameter(torch.zeros(c))
        nn.init.normal_(self.weight, std=(3 * c) ** -0.5)

    def pack(self):
        co, ci = self.weight.shape[:2]
        return (self.weight.detach().to(BF16).float().reshape(co, ci, 3).cpu().contiguous(),
                self.bias.detach().to(BF16).float().cpu().contiguous())

    def forward(self, x_nhwc, frames: int):
        w, b = self.packed()
        return ops.frame_conv_small(x_nhwc, w, b, frames, w.shape[1])


class TemporalDecoder(nn.Module):
    def __init__(self, in_channels: int = 4, out_channels: int = 3, block_out_channels=(128, 256, 512, 512),
                 layers_per_block: int = 2):
        super().__init__()
        if out_channels > 4:
            raise NotImplementedError("time_conv_out is implemented for at most 4 image channels")
        rev = list(reversed(block_out_channels))
        self.conv_in = Conv2d(in_channels, rev[0], 3)
        self.mid_block = _MidBlockTemporalDecoder(rev[0], layers_per_block)
        blocks, prev = [], rev[0]
        for i, c in enumerate(rev):
            blocks.append(_UpBlockTemporalDecoder(prev, c, layers_per_block + 1, i != len(rev) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = Norm(rev[-1], 1e-6)
        self.conv_out = Conv2d(rev[-1], out_channels, 3)
        self.time_conv_out = _FrameConvSmall(out_channels)
        self._conv_out_pad = None

    _conv_out_packed = Decoder._conv_out_packed

    def forward(self, z_nhwc8, frames: int):
        """z [B*F, h, w, 8] (frames of a clip contiguous) -> images [B*F, 3, 8h, 8w] bf16 (logical NCHW)."""
        x = self.conv_in(z_nhwc8)
        x = self.mid_block(x, frames)
        for blk in self.up_blocks:
            x = blk(x, frames)
        x = self.conv_norm_out.group_norm(x, silu=True)
        w, b = self._conv_out_packed()
        y = ops.conv2d(x, w, b, ksize=3)                       # [n, H, W, 8], the first out_channels are real
        return self.time_conv_out(y, frames)


class AutoencoderKLTemporalDecoder(_DecoderOnlyVae):
    """Decoder half of the SVD VAE (stabilityai/stable-video-diffusion-img2vid[-xt] ``vae``): same constructor kwargs,
    ``decode(z, num_frames)`` call form and ``decoder.*`` state-dict keys as the diffusers class.  There is no
    post_quant_conv in this model; ``encoder.*`` / ``quant_conv.*`` keys of a checkpoint are skipped (the conditioning
    image is encoded once per call, outside the path)."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, down_block_types=("DownEncoderBlock2D",) * 4,
                 block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, latent_channels: int = 4,
                 sample_size: int = 768, scaling_factor: float = 0.18215, force_upcast: bool = True, **_ignored):
        super().__init__()
        self.config = _ConfigDict(in_channels=in_channels, out_channels=out_channels,
                                  down_block_types=tuple(down_block_types), block_out_channels=tuple(block_out_channels),
                                  layers_per_block=layers_per_block, latent_channels=latent_channels,
                                  sample_size=sample_size, scaling_factor=scaling_factor, force_upcast=force_upcast)
        self.decoder = TemporalDecoder(latent_channels, out_channels, block_out_channels, layers_per_block)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, num_frames: int = 1, return_dict: bool = True):
        """z [B*F, 4, h, w] (already divided by scaling_factor, frames of a clip contiguous; B*F a multiple of
        num_frames) -> frames [B*F, 3, 8h, 8w] in [-1, 1] (logical NCHW, bf16)."""
        if z.shape[0] % num_frames:
            raise ValueError(f"batch {z.shape[0]} is not a multiple of num_frames={num_frames}")
        x = to_channels_last_bf16(z.to(BF16) if z.dtype not in (BF16, torch.float32) else z, 8)
        img = self.decoder(x, num_frames)
        return DecoderOutput(img) if return_dict else (img,)


def svd_decode_latents(vae: AutoencoderKLTemporalDecoder, latents: torch.Tensor, num_frames: int,
                       decode_chunk_size: int = 14):
    """svd pipeline :265-292: latents (B, F, 4, h, w) -> video (B, 3, F, H, W) fp32.  Frames are decoded
    ``decode_chunk_size`` at a time and, as in the reference, each chunk is ONE temporal unit for the decoder
    (``num_frames`` = the chunk's length), whatever clip its frames come from."""
    latents = latents.flatten(0, 1)
    latents = 1 / vae.config.scaling_factor * latents
    frames = []
    for i in range(0, latents.shape[0], decode_chunk_size):
        chunk = latents[i:i + decode_chunk_size].contiguous()
        frames.append(vae.decode(chunk, num_frames=chunk.shape[0]).sample)
    frames = torch.cat(frames, dim=0)
    frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
    return frames.float()


# ---- the pipelines' post-processing (diffusers VaeImageProcessor.postprocess / i2vgen tensor2vid) --------------------
def postprocess(image: torch.Tensor, output_type: str = "pil"):
    """[N, 3, H, W] in [-1, 1] -> "pt" tensor in [0, 1] / "np" NHWC float array / list of PIL images."""
    image = (image.float() / 2 + 0.5).clamp(0, 1)
    if output_type == "pt":
        return image
    arr = image.cpu().permute(0, 2, 3, 1).numpy()
    if output_type == "np":
        return arr
    if output_type == "pil":
        from PIL import Image
        return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
    raise ValueError(f"{output_type} does not exist. Please choose one of ['np', 'pt', 'pil']")


def decode_latents(vae: AutoencoderKL, latents: torch.Tensor, decode_chunk_size: Optional[int] = None):
    """i2vgen pipeline :398-418: latents (B, 4, F, h, w) -> video (B, 3, F, H, W) fp32."""
    latents = 1 / vae.config.scaling_factor * latents
    b, c, f, h, w = latents.shape
    lat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    step = decode_chunk_size or lat.shape[0]
    frames = [vae.decode(lat[i:i + step].contiguous()).sample for i in range(0, lat.shape[0], step)]
    image = torch.cat(frames, dim=0)
    return image.reshape(b, f, -1, *image.shape[2:]).permute(0, 2, 1, 3, 4).float()


def tensor2vid(video: torch.Tensor, output_type: str = "np"):
    """i2vgen pipeline :81-99."""
    import numpy as np
    outs = [postprocess(video[i].permute(1, 0, 2, 3), output_type) for i in range(video.shape[0])]
    if output_type == "np":
        return np.stack(outs)
    if output_type == "pt":
        return torch.stack(outs)
    return outs
