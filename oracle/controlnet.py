"""ORACLE (test infrastructure): SD1.5 ControlNet restated on oracle.blocks.

Follows /root/reference/controlnet/controlnet.py: ControlNetConditioningEmbedding (:62-104), ControlNetModel
constructor (:179-438, default SD1.5 configuration only) and forward (:662-881, including the reference's
``skip_conv_in`` / ``skip_time_emb`` additions at :802-811), and /root/reference/controlnet/multicontrolnet.py:45-99.
Not imported by the product package.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn

from .adapter import zero_module
from .blocks import TimestepEmbedding, Timesteps, UNetMidBlock2DCrossAttn, get_down_block


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, conditioning_embedding_channels: int, conditioning_channels: int = 3,
                 block_out_channels: Tuple[int, ...] = (16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, block_out_channels[0], kernel_size=3, padding=1)
        self.blocks = nn.ModuleList([])
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, kernel_size=3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, kernel_size=3, padding=1, stride=2))
        self.conv_out = zero_module(nn.Conv2d(block_out_channels[-1], conditioning_embedding_channels, kernel_size=3, padding=1))

    def forward(self, conditioning):
        embedding = F.silu(self.conv_in(conditioning))
        for block in self.blocks:
            embedding = F.silu(block(embedding))
        return self.conv_out(embedding)


class _Config(dict):
    __getattr__ = dict.__getitem__


class ControlNetModel(nn.Module):
    def __init__(self, in_channels: int = 4, conditioning_channels: int = 3, flip_sin_to_cos: bool = True,
                 freq_shift: int = 0,
                 down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn", only_cross_attention=False,
                 block_out_channels=(320, 640, 1280, 1280), layers_per_block: int = 2, downsample_padding: int = 1,
                 mid_block_scale_factor: float = 1, act_fn: str = "silu", norm_num_groups: Optional[int] = 32,
                 norm_eps: float = 1e-5, cross_attention_dim: int = 1280, transformer_layers_per_block=1,
                 attention_head_dim=8, num_attention_heads=None, use_linear_projection: bool = False,
                 upcast_attention: bool = False, controlnet_conditioning_channel_order: str = "rgb",
                 conditioning_embedding_out_channels=(16, 32, 96, 256), global_pool_conditions: bool = False):
        super().__init__()
        self.config = _Config(in_channels=in_channels, controlnet_conditioning_channel_order=controlnet_conditioning_channel_order,
                              global_pool_conditions=global_pool_conditions, block_out_channels=tuple(block_out_channels),
                              cross_attention_dim=cross_attention_dim, addition_embed_type=None)
        num_attention_heads = num_attention_heads or attention_head_dim  # :227
        if len(block_out_channels) != len(down_block_types):
            raise ValueError("Must provide the same number of `block_out_channels` as `down_block_types`.")
        n = len(down_block_types)
        if isinstance(transformer_layers_per_block, int):
            transformer_layers_per_block = [transformer_layers_per_block] * n
        if isinstance(only_cross_attention, bool):
            only_cross_attention = [only_cross_attention] * n
        if isinstance(attention_head_dim, int):
            attention_head_dim = (attention_head_dim,) * n
        if isinstance(num_attention_heads, int):
            num_attention_heads = (num_attention_heads,) * n
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        time_embed_dim = block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim, act_fn=act_fn)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(
            conditioning_embedding_channels=block_out_channels[0], block_out_channels=conditioning_embedding_out_channels,
            conditioning_channels=conditioning_channels)
        self.down_blocks = nn.ModuleList([])
        self.controlnet_down_blocks = nn.ModuleList([])
        output_channel = block_out_channels[0]
        self.controlnet_down_blocks.append(zero_module(nn.Conv2d(output_channel, output_channel, kernel_size=1)))
        for i, down_block_type in enumerate(down_block_types):  # :366-401
            input_channel = output_channel
            output_channel = block_out_channels[i]
            is_final_block = i == len(block_out_channels) - 1
            self.down_blocks.append(get_down_block(
                down_block_type, num_layers=layers_per_block, transformer_layers_per_block=transformer_layers_per_block[i],
                in_channels=input_channel, out_channels=output_channel, temb_channels=time_embed_dim,
                add_downsample=not is_final_block, resnet_eps=norm_eps, resnet_act_fn=act_fn, resnet_groups=norm_num_groups,
                cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads[i],
                attention_head_dim=attention_head_dim[i] if attention_head_dim[i] is not None else output_channel,
                downsample_padding=downsample_padding, use_linear_projection=use_linear_projection,
                only_cross_attention=only_cross_attention[i], upcast_attention=upcast_attention))
            for _ in range(layers_per_block):
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(output_channel, output_channel, kernel_size=1)))
            if not is_final_block:
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(output_channel, output_channel, kernel_size=1)))
        mid_block_channel = block_out_channels[-1]
        self.controlnet_mid_block = zero_module(nn.Conv2d(mid_block_channel, mid_block_channel, kernel_size=1))
        assert mid_block_type == "UNetMidBlock2DCrossAttn"
        self.mid_block = UNetMidBlock2DCrossAttn(
            transformer_layers_per_block=transformer_layers_per_block[-1], in_channels=mid_block_channel,
            temb_channels=time_embed_dim, resnet_eps=norm_eps, output_scale_factor=mid_block_scale_factor,
            cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads[-1],
            resnet_groups=norm_num_groups, use_linear_projection=use_linear_projection, upcast_attention=upcast_attention)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale: float = 1.0,
                class_labels=None, timestep_cond=None, attention_mask=None, added_cond_kwargs=None,
                cross_attention_kwargs=None, guess_mode: bool = False, return_dict: bool = True,
                skip_conv_in: bool = False, skip_time_emb: bool = False):
        order = self.config.controlnet_conditioning_channel_order
        if order == "bgr":
            controlnet_cond = torch.flip(controlnet_cond, dims=[1])
        elif order != "rgb":
            raise ValueError(f"unknown `controlnet_conditioning_channel_order`: {order}")
        timesteps = timestep  # :735-749
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)  # :751-756 (exact t, no bf16 rounding of t)
        emb = self.time_embedding(t_emb, timestep_cond)
        sample_conv_in = self.conv_in(sample)  # :802-811
        sample = torch.zeros_like(sample_conv_in) if skip_conv_in else sample_conv_in
        if skip_time_emb:
            emb = torch.zeros_like(emb)
        controlnet_cond = self.controlnet_cond_embedding(controlnet_cond)
        sample = sample + controlnet_cond
        down_block_res_samples = (sample,)
        for block in self.down_blocks:
            if getattr(block, "has_cross_attention", False):
                sample, res_samples = block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample, res_samples = block(hidden_states=sample, temb=emb)
            down_block_res_samples += res_samples
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        outs = ()
        for res, blk in zip(down_block_res_samples, self.controlnet_down_blocks):
            outs = outs + (blk(res),)
        mid = self.controlnet_mid_block(sample)
        if guess_mode and not self.config.global_pool_conditions:  # :861-868
            scales = torch.logspace(-1, 0, len(outs) + 1, device=sample.device) * conditioning_scale
            outs = [s * sc for s, sc in zip(outs, scales)]
            mid = mid * scales[-1]
        else:
            outs = [s * conditioning_scale for s in outs]
            mid = mid * conditioning_scale
        if self.config.global_pool_conditions:
            outs = [torch.mean(s, dim=(2, 3), keepdim=True) for s in outs]
            mid = torch.mean(mid, dim=(2, 3), keepdim=True)
        return (outs, mid)


class MultiControlNetModel(nn.Module):
    """multicontrolnet.py:45-99: runs nets[k] on controlnet_cond[k]; zip() truncates to the shorter list (quirk Q7);
    returns LISTS of per-net outputs (the reference removed the summation so the router can weight them)."""

    def __init__(self, controlnets: Union[List[ControlNetModel], Tuple[ControlNetModel]]):
        super().__init__()
        self.nets = nn.ModuleList(controlnets)

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale, class_labels=None,
                timestep_cond=None, attention_mask=None, added_cond_kwargs=None, cross_attention_kwargs=None,
                guess_mode: bool = False, return_dict: bool = True, skip_conv_in: bool = False,
                skip_time_emb: bool = False):
        all_down, all_mid = [], []
        for image, scale, net in zip(controlnet_cond, conditioning_scale, self.nets):
            down, mid = net(sample=sample, timestep=timestep, encoder_hidden_states=encoder_hidden_states,
                            controlnet_cond=image, conditioning_scale=scale, guess_mode=guess_mode, return_dict=False,
                            skip_conv_in=skip_conv_in, skip_time_emb=skip_time_emb)
            all_down.append(down)
            all_mid.append(mid)
        return all_down, all_mid
