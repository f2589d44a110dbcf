"""ORACLE (test infrastructure): I2VGen-XL UNet restated on oracle.blocks.

Follows /root/reference/i2vgen_xl/models/unets/unet_i2vgen_xl.py: I2VGenXLTransformerTemporalEncoder (:51-101),
I2VGenXLUNet constructor (:131-316) and forward (:519-761) including the reference's residual-injection additions
(:681-695, :709-714).  Pinned against the reference's own class (run through oracle/diffusers_shim) by
tests/golden/make_golden.py.  Not imported by the product package.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .blocks import (Attention, FeedForward, TimestepEmbedding, Timesteps, TransformerTemporalModel,
                     UNetMidBlock3DCrossAttn, get_down_block_3d, get_up_block_3d)


class I2VGenXLTransformerTemporalEncoder(nn.Module):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, activation_fn: str = "geglu",
                 ff_inner_dim: Optional[int] = None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=True, eps=1e-5)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, bias=False,
                               out_bias=True)
        self.ff = FeedForward(dim, activation_fn=activation_fn, inner_dim=ff_inner_dim, bias=True)

    def forward(self, hidden_states):
        hidden_states = self.attn1(self.norm1(hidden_states), encoder_hidden_states=None) + hidden_states
        return self.ff(hidden_states) + hidden_states  # NB: the FF input is NOT normalised (:95)


class _Config(dict):
    __getattr__ = dict.__getitem__


class I2VGenXLUNet(nn.Module):
    def __init__(self, sample_size=None, in_channels: int = 4, out_channels: int = 4,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 block_out_channels=(320, 640, 1280, 1280), layers_per_block: int = 2, norm_num_groups: int = 32,
                 cross_attention_dim: int = 1024, attention_head_dim=64, num_attention_heads=None):
        super().__init__()
        self.config = _Config(in_channels=in_channels, out_channels=out_channels, cross_attention_dim=cross_attention_dim,
                              block_out_channels=tuple(block_out_channels))
        num_attention_heads = attention_head_dim  # :165 (diffusers naming quirk)
        c0 = block_out_channels[0]
        self.conv_in = nn.Conv2d(in_channels + in_channels, c0, kernel_size=3, padding=1)
        self.transformer_in = TransformerTemporalModel(num_attention_heads=8, attention_head_dim=num_attention_heads,
                                                       in_channels=c0, num_layers=1, norm_num_groups=norm_num_groups)
        self.image_latents_proj_in = nn.Sequential(
            nn.Conv2d(4, in_channels * 4, 3, padding=1), nn.SiLU(),
            nn.Conv2d(in_channels * 4, in_channels * 4, 3, stride=1, padding=1), nn.SiLU(),
            nn.Conv2d(in_channels * 4, in_channels, 3, stride=1, padding=1))
        self.image_latents_temporal_encoder = I2VGenXLTransformerTemporalEncoder(
            dim=in_channels, num_attention_heads=2, ff_inner_dim=in_channels * 4, attention_head_dim=in_channels,
            activation_fn="gelu")
        self.image_latents_context_embedding = nn.Sequential(
            nn.Conv2d(4, in_channels * 8, 3, padding=1), nn.SiLU(), nn.AdaptiveAvgPool2d((32, 32)),
            nn.Conv2d(in_channels * 8, in_channels * 16, 3, stride=2, padding=1), nn.SiLU(),
            nn.Conv2d(in_channels * 16, cross_attention_dim, 3, stride=2, padding=1))
        time_embed_dim = c0 * 4
        self.time_proj = Timesteps(c0, True, 0)
        self.time_embedding = TimestepEmbedding(c0, time_embed_dim, act_fn="silu")
        self.context_embedding = nn.Sequential(nn.Linear(cross_attention_dim, time_embed_dim), nn.SiLU(),
                                               nn.Linear(time_embed_dim, cross_attention_dim * in_channels))
        self.fps_embedding = nn.Sequential(nn.Linear(c0, time_embed_dim), nn.SiLU(),
                                           nn.Linear(time_embed_dim, time_embed_dim))
        n = len(down_block_types)
        if isinstance(num_attention_heads, int):
            num_attention_heads = (num_attention_heads,) * n
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        output_channel = c0
        for i, t in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            self.down_blocks.append(get_down_block_3d(
                t, num_layers=layers_per_block, in_channels=input_channel, out_channels=output_channel,
                temb_channels=time_embed_dim, add_downsample=i != n - 1, resnet_eps=1e-05, resnet_groups=norm_num_groups,
                cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads[i],
                downsample_padding=1, dual_cross_attention=False))
        self.mid_block = UNetMidBlock3DCrossAttn(
            in_channels=block_out_channels[-1], temb_channels=time_embed_dim, resnet_eps=1e-05, output_scale_factor=1,
            cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads[-1],
            resnet_groups=norm_num_groups, dual_cross_attention=False)
        rev_ch = list(reversed(block_out_channels))
        rev_heads = list(reversed(num_attention_heads))
        output_channel = rev_ch[0]
        for i, t in enumerate(up_block_types):
            prev_output_channel = output_channel
            output_channel = rev_ch[i]
            input_channel = rev_ch[min(i + 1, n - 1)]
            self.up_blocks.append(get_up_block_3d(
                t, num_layers=layers_per_block + 1, in_channels=input_channel, out_channels=output_channel,
                prev_output_channel=prev_output_channel, temb_channels=time_embed_dim, add_upsample=i != n - 1,
                resnet_eps=1e-05, resnet_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                num_attention_heads=rev_heads[i], dual_cross_attention=False))
        self.conv_norm_out = nn.GroupNorm(num_channels=c0, num_groups=norm_num_groups, eps=1e-05)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c0, out_channels, kernel_size=3, padding=1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, sample, timestep, fps, image_latents, image_embeddings=None, encoder_hidden_states=None,
                timestep_cond=None, cross_attention_kwargs=None, return_dict: bool = False,
                down_block_additional_residuals=None, mid_block_additional_residual=None):
        batch_size, channels, num_frames, height, width = sample.shape
        # 1./2./3. time + fps embeddings (:576-596)
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timesteps, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_embedding(self.time_proj(timesteps).to(dtype=self.dtype), timestep_cond)
        fps = fps.expand(fps.shape[0])
        fps_emb = self.fps_embedding(self.time_proj(fps).to(dtype=self.dtype))
        emb = (t_emb + fps_emb).repeat_interleave(repeats=num_frames, dim=0)
        # 4. context embeddings: text tokens + 64 image-latent tokens + 4 image-embedding tokens (:598-635)
        context_emb = sample.new_zeros(batch_size, 0, self.config.cross_attention_dim)
        context_emb = torch.cat([context_emb, encoder_hidden_states], dim=1)
        il0 = image_latents[:, :, :1, :]
        il0 = il0.permute(0, 2, 1, 3, 4).reshape(il0.shape[0] * il0.shape[2], il0.shape[1], il0.shape[3], il0.shape[4])
        il0 = self.image_latents_context_embedding(il0)
        b_, c_, h_, w_ = il0.shape
        context_emb = torch.cat([context_emb, il0.permute(0, 2, 3, 1).reshape(b_, h_ * w_, c_)], dim=1)
        image_emb = self.context_embedding(image_embeddings).view(-1, self.config.in_channels,
                                                                  self.config.cross_attention_dim)
        context_emb = torch.cat([context_emb, image_emb], dim=1).repeat_interleave(repeats=num_frames, dim=0)
        # image latents -> per-pixel temporal encoder (:637-651)
        il = image_latents.permute(0, 2, 1, 3, 4).reshape(image_latents.shape[0] * image_latents.shape[2],
                                                         image_latents.shape[1], image_latents.shape[3],
                                                         image_latents.shape[4])
        il = self.image_latents_proj_in(il)
        il = (il[None, :].reshape(batch_size, num_frames, channels, height, width).permute(0, 3, 4, 1, 2)
              .reshape(batch_size * height * width, num_frames, channels))
        il = self.image_latents_temporal_encoder(il)
        il = il.reshape(batch_size, height, width, num_frames, channels).permute(0, 4, 3, 1, 2)
        # 5. pre-process (:654-662)
        sample = torch.cat([sample, il], dim=1)
        sample = sample.permute(0, 2, 1, 3, 4).reshape((sample.shape[0] * num_frames, -1) + sample.shape[3:])
        sample = self.conv_in(sample)
        sample = self.transformer_in(sample, num_frames=num_frames)[0]
        # 6. down (:666-678)
        down_block_res_samples = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, res = blk(hidden_states=sample, temb=emb, encoder_hidden_states=context_emb, num_frames=num_frames)
            else:
                sample, res = blk(hidden_states=sample, temb=emb, num_frames=num_frames)
            down_block_res_samples += res
        # reference addition: ControlNet/adapter residuals (:681-695); 5-D inputs are (b c f h w)
        if down_block_additional_residuals is not None:
            if down_block_additional_residuals[0].dim() == 5:
                down_block_additional_residuals = [r.permute(0, 2, 1, 3, 4).reshape((-1, r.shape[1]) + r.shape[3:])
                                                   for r in down_block_additional_residuals]
            new = ()
            for s, r in zip(down_block_res_samples, down_block_additional_residuals):
                new = new + (s + r,)
            down_block_res_samples = new
        # 7. mid (:698-714)
        sample = self.mid_block(sample, emb, encoder_hidden_states=context_emb, num_frames=num_frames)
        if mid_block_additional_residual is not None:
            m = mid_block_additional_residual
            if m.dim() == 5:
                m = m.permute(0, 2, 1, 3, 4).reshape((-1, m.shape[1]) + m.shape[3:])
            sample = sample + m
        # 8. up (:719-747)
        for blk in self.up_blocks:
            res = down_block_res_samples[-len(blk.resnets):]
            down_block_res_samples = down_block_res_samples[: -len(blk.resnets)]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             encoder_hidden_states=context_emb, num_frames=num_frames)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res, num_frames=num_frames)
        # 9. post-process (:750-756)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        sample = sample[None, :].reshape((-1, num_frames) + sample.shape[1:]).permute(0, 2, 1, 3, 4)
        return (sample,)
