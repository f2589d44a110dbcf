"""ORACLE (test infrastructure): the decoder half of diffusers v0.27.2 ``AutoencoderKL`` restated in plain PyTorch --
what ``self.vae.decode`` runs in /root/reference/sdxl/pipelines/sdxl_controlnet_adapter_pipeline.py:1414 and
i2vgen_xl/pipelines/i2vgen_xl_controlnet_adapter_pipeline.py:398-418.  diffusers is a third-party dependency that is not
vendored in the reference and cannot be installed here: this restatement (models/autoencoders/vae.py ``Decoder``,
models/unets/unet_2d_blocks.py ``UNetMidBlock2D`` / ``UpDecoderBlock2D``, models/attention_processor.py ``Attention`` with
``AttnProcessor2_0`` for the deprecated-attention-block form) is **parity unpinned** -- no upstream vectors exist in
/root/reference; the pin is structural (the published SDXL VAE's decoder + post_quant_conv parameter count, 49 490 199, and
state-dict key names).  Not imported by the product."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .blocks import ResnetBlock2D, SpatioTemporalResBlock, Upsample2D


class VaeAttention(nn.Module):
    """Attention(query_dim=C, heads=C // dim_head, dim_head, norm_num_groups=32, eps, residual_connection=True, bias=True,
    rescale_output_factor=1, upcast_softmax=True) on a 4-D input (attention_processor.py AttnProcessor2_0.__call__)."""

    def __init__(self, channels: int, dim_head: int, groups: int, eps: float):
        super().__init__()
        self.heads = channels // dim_head
        self.group_norm = nn.GroupNorm(num_channels=channels, num_groups=groups, eps=eps, affine=True)
        self.to_q = nn.Linear(channels, channels, bias=True)
        self.to_k = nn.Linear(channels, channels, bias=True)
        self.to_v = nn.Linear(channels, channels, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, temb=None):
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(hidden_states), self.to_k(hidden_states), self.to_v(hidden_states)
        hd = c // self.heads
        q, k, v = (t.view(b, -1, self.heads, hd).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, c).to(q.dtype)
        o = self.to_out[1](self.to_out[0](o))
        o = o.transpose(-1, -2).reshape(b, c, h, w)
        return (o + residual) / 1.0


class UNetMidBlock2D(nn.Module):
    def __init__(self, in_channels: int, eps: float, groups: int, attention_head_dim: int):
        super().__init__()
        kw = dict(in_channels=in_channels, out_channels=in_channels, temb_channels=None, eps=eps, groups=groups)
        self.resnets = nn.ModuleList([ResnetBlock2D(**kw), ResnetBlock2D(**kw)])
        self.attentions = nn.ModuleList([VaeAttention(in_channels, attention_head_dim, groups, eps)])

    def forward(self, hidden_states, temb=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, temb=temb)
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, eps: float, groups: int, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels,
                                                    out_channels=out_channels, temb_channels=None, eps=eps, groups=groups)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None

    def forward(self, hidden_states, temb=None):
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb=temb)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                hidden_states = up(hidden_states)
        return hidden_states


class Decoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(in_channels, rev[0], kernel_size=3, stride=1, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0], 1e-6, norm_num_groups, attention_head_dim=rev[0])
        self.up_blocks = nn.ModuleList([])
        out_c = rev[0]
        for i, c in enumerate(rev):
            prev, out_c = out_c, c
            self.up_blocks.append(UpDecoderBlock2D(prev, out_c, layers_per_block + 1, 1e-6, norm_num_groups,
                                                   add_upsample=i != len(rev) - 1))
        self.conv_norm_out = nn.GroupNorm(num_channels=rev[-1], num_groups=norm_num_groups, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, sample, latent_embeds=None):
        sample = self.conv_in(sample)
        upscale_dtype = next(iter(self.up_blocks.parameters())).dtype
        sample = self.mid_block(sample, latent_embeds)
        sample = sample.to(upscale_dtype)
        for up_block in self.up_blocks:
            sample = up_block(sample, latent_embeds)
        sample = self.conv_norm_out(sample)
        sample = self.conv_act(sample)
        return self.conv_out(sample)


class AutoencoderKL(nn.Module):
    """decode() half only: ``z = post_quant_conv(z); dec = decoder(z)`` (autoencoder_kl.py _decode)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.13025, **_unused):
        super().__init__()
        self.scaling_factor = scaling_factor
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def decode(self, z, return_dict=False):
        return (self.decoder(self.post_quant_conv(z)),)


# ---- SVD: diffusers v0.27.2 models/autoencoders/autoencoder_kl_temporal_decoder.py (TemporalDecoder,
# AutoencoderKLTemporalDecoder.decode) and models/unets/unet_3d_blocks.py (MidBlockTemporalDecoder,
# UpBlockTemporalDecoder), restated; what ``self.vae.decode(..., num_frames=)`` runs in
# /root/reference/svd/pipelines/svd_controlnet_adapter_pipeline.py:265-292.  Parity unpinned like the rest of this file.
def _st_block(cin, cout):
    return SpatioTemporalResBlock(in_channels=cin, out_channels=cout, temb_channels=None, eps=1e-6, temporal_eps=1e-5,
                                  merge_factor=0.0, merge_strategy="learned", switch_spatial_to_temporal_mix=True)


class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, attention_head_dim: int = 512, num_layers: int = 1):
        super().__init__()
        self.resnets = nn.ModuleList([_st_block(in_channels if i == 0 else out_channels, out_channels)
                                      for i in range(num_layers)])
        self.attentions = nn.ModuleList([VaeAttention(in_channels, attention_head_dim, 32, 1e-6)])

    def forward(self, hidden_states, image_only_indicator):
        hidden_states = self.resnets[0](hidden_states, image_only_indicator=image_only_indicator)
        for resnet, attn in zip(self.resnets[1:], self.attentions):
            hidden_states = attn(hidden_states)
            hidden_states = resnet(hidden_states, image_only_indicator=image_only_indicator)
        return hidden_states


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int = 1, add_upsample: bool = True):
        super().__init__()
        self.resnets = nn.ModuleList([_st_block(in_channels if i == 0 else out_channels, out_channels)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None

    def forward(self, hidden_states, image_only_indicator):
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, image_only_indicator=image_only_indicator)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                hidden_states = up(hidden_states)
        return hidden_states


class TemporalDecoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[-1], kernel_size=3, stride=1, padding=1)
        self.mid_block = MidBlockTemporalDecoder(num_layers=layers_per_block, in_channels=block_out_channels[-1],
                                                 out_channels=block_out_channels[-1],
                                                 attention_head_dim=block_out_channels[-1])
        self.up_blocks = nn.ModuleList([])
        rev = list(reversed(block_out_channels))
        out_c = rev[0]
        for i, c in enumerate(rev):
            prev, out_c = out_c, c
            self.up_blocks.append(UpBlockTemporalDecoder(num_layers=layers_per_block + 1, in_channels=prev,
                                                         out_channels=out_c, add_upsample=i != len(rev) - 1))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=32, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, kernel_size=(3, 1, 1), padding=(1, 0, 0))

    def forward(self, sample, image_only_indicator, num_frames: int = 1):
        sample = self.conv_in(sample)
        upscale_dtype = next(iter(self.up_blocks.parameters())).dtype
        sample = self.mid_block(sample, image_only_indicator=image_only_indicator)
        sample = sample.to(upscale_dtype)
        for up_block in self.up_blocks:
            sample = up_block(sample, image_only_indicator=image_only_indicator)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        batch_frames, channels, height, width = sample.shape
        batch_size = batch_frames // num_frames
        sample = sample[None, :].reshape(batch_size, num_frames, channels, height, width).permute(0, 2, 1, 3, 4)
        sample = self.time_conv_out(sample)
        return sample.permute(0, 2, 1, 3, 4).reshape(batch_frames, channels, height, width)


class AutoencoderKLTemporalDecoder(nn.Module):
    """decode() half only (no post_quant_conv in this model): image_only_indicator = zeros(batch, num_frames)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, scaling_factor=0.18215, **_unused):
        super().__init__()
        self.scaling_factor = scaling_factor
        self.decoder = TemporalDecoder(latent_channels, out_channels, block_out_channels, layers_per_block)

    def decode(self, z, num_frames: int = 1, return_dict=False):
        batch_size = z.shape[0] // num_frames
        image_only_indicator = torch.zeros(batch_size, num_frames, dtype=z.dtype, device=z.device)
        return (self.decoder(z, num_frames=num_frames, image_only_indicator=image_only_indicator),)
