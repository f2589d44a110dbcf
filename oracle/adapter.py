"""ORACLE (test infrastructure): restatement of the reference's adapter modules on top of oracle.blocks.

Follows /root/reference/model/adapter_spatial_temporal.py (AdapterSpatioTemporal, :10-292) and
/root/reference/model/ctrl_adapter.py (ControlNetAdapter, :12-224).  Validated against the reference's own classes
(imported through oracle/diffusers_shim) by tests/golden/make_golden.py; see tests/test_oracle_golden.py.
Not imported by the product package.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .blocks import (AlphaBlender, BasicTransformerBlock, ResnetBlock2D, TemporalBasicTransformerBlock,
                     TemporalResnetBlock, TimestepEmbedding, Timesteps)


def zero_module(module):  # controlnet/controlnet.py:884-887
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


class AdapterSpatioTemporal(nn.Module):
    """adapter_spatial_temporal.py:10-171 (constructor) and :175-292 (forward)."""

    def __init__(self, in_channels: int, out_channels: int, num_layers: int = 1, add_spatial_resnet: bool = True,
                 add_temporal_resnet: bool = True, add_spatial_transformer: bool = True,
                 add_temporal_transformer: bool = True, eps: float = 1e-6, temporal_eps: float = None,
                 merge_factor: float = 0.5, merge_strategy="learned_with_images",
                 switch_spatial_to_temporal_mix: bool = False, up_sampling_scale: float = 1.0,
                 cross_attention_dim: int = 1024, num_attention_heads: int = 8, attention_head_dim: int = 64):
        super().__init__()
        temb_channels = in_channels
        self.num_attention_heads = in_channels // attention_head_dim  # :42 (quirk Q1: heads = C // 64)
        self.num_layers = num_layers
        self.up_sampling_scale = up_sampling_scale
        self.add_spatial_resnet = add_spatial_resnet
        self.add_temporal_resnet = add_temporal_resnet
        self.add_spatial_transformer = add_spatial_transformer
        self.add_temporal_transformer = add_temporal_transformer
        self.add_resnet_time_mixer = add_spatial_resnet and add_temporal_resnet
        self.add_transformer_time_mixer = add_spatial_transformer and add_temporal_transformer

        if add_spatial_resnet or add_temporal_resnet:  # :55-57
            self.resnet_time_proj = Timesteps(out_channels, True, downscale_freq_shift=0)
            self.resnet_time_embedding = TimestepEmbedding(in_channels, in_channels)
        if add_spatial_transformer or add_temporal_transformer:  # :60-69
            self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6)
            self.inner_dim = num_attention_heads * attention_head_dim  # constructor default 8*64 = 512 (quirk Q1)
            if add_temporal_transformer:
                self.transformer_time_embedding = TimestepEmbedding(in_channels, self.inner_dim)
                self.transformer_time_proj = Timesteps(in_channels, True, 0)
            self.proj_in = nn.Linear(in_channels, self.inner_dim)
            self.proj_out = nn.Linear(self.inner_dim, in_channels)

        sr, tr, sa, ta, rm, tm = [], [], [], [], [], []
        for i in range(num_layers):  # :77-152
            if add_spatial_resnet:
                sr.append(ResnetBlock2D(in_channels=in_channels, out_channels=out_channels, temb_channels=temb_channels,
                                        eps=eps, use_in_shortcut=True, up=(i == 0 and up_sampling_scale > 1)))
            if add_temporal_resnet:
                tr.append(TemporalResnetBlock(in_channels=out_channels if add_spatial_resnet else in_channels,
                                              out_channels=out_channels, temb_channels=temb_channels,
                                              eps=temporal_eps if temporal_eps is not None else eps))
            if add_spatial_transformer:
                sa.append(BasicTransformerBlock(self.inner_dim, self.num_attention_heads, attention_head_dim,
                                                cross_attention_dim=cross_attention_dim))
            if add_temporal_transformer:
                ta.append(TemporalBasicTransformerBlock(self.inner_dim, self.inner_dim, self.num_attention_heads,
                                                        attention_head_dim, cross_attention_dim=cross_attention_dim))
            if self.add_resnet_time_mixer:
                rm.append(AlphaBlender(alpha=merge_factor, merge_strategy=merge_strategy,
                                       switch_spatial_to_temporal_mix=switch_spatial_to_temporal_mix))
            if self.add_transformer_time_mixer:
                tm.append(AlphaBlender(alpha=merge_factor, merge_strategy=merge_strategy,
                                       switch_spatial_to_temporal_mix=switch_spatial_to_temporal_mix))
        if add_spatial_resnet:
            self.spatial_resnets = nn.ModuleList(sr)
        if add_temporal_resnet:
            self.temporal_resnets = nn.ModuleList(tr)
        if add_spatial_transformer:
            self.spatial_attentions = nn.ModuleList(sa)
        if add_temporal_transformer:
            self.temporal_attentions = nn.ModuleList(ta)
        if self.add_resnet_time_mixer:
            self.resnets_time_mixer = nn.ModuleList(rm)
        if self.add_transformer_time_mixer:
            self.transformers_time_mixer = nn.ModuleList(tm)

    def forward(self, hidden_states, num_frames: int, timestep=None, encoder_hidden_states=None, sparsity_masking=None):
        batch_frames, channels, height, width = hidden_states.shape
        batch_size = batch_frames // num_frames
        dev = hidden_states.device
        # :190-198 timestep normalisation.  The int/float branch of the reference uses the legacy
        # torch.Tensor(..., device=) constructor (quirk Q18, CPU only); values are identical.
        if isinstance(timestep, (int, float)):
            timestep = torch.tensor([float(timestep)], device=dev).repeat_interleave(batch_frames, dim=0)
        elif isinstance(timestep, torch.Tensor) and timestep.dim() == 0:
            timestep = torch.tensor([float(timestep)], device=dev).repeat_interleave(batch_frames, dim=0)
        elif isinstance(timestep, torch.Tensor) and timestep.dim() == 1 and len(timestep) == 1:
            timestep = timestep.float().to(dev).repeat_interleave(batch_frames, dim=0)
        elif isinstance(timestep, torch.Tensor) and timestep.dim() == 2:
            timestep = timestep.squeeze()
        timestep = timestep.to(hidden_states.dtype)  # :198 (quirk Q2: bf16 rounding of t before the sinusoid)
        image_only_indicator = torch.zeros(batch_size, num_frames, dtype=hidden_states.dtype, device=dev)

        for i in range(self.num_layers):
            if self.add_spatial_resnet or self.add_temporal_resnet:  # :206-209
                resnet_temb = self.resnet_time_proj(timestep)
                resnet_temb = self.resnet_time_embedding(resnet_temb.to(self.resnet_time_embedding.linear_1.weight.dtype)
                                                         if not torch.is_autocast_enabled() else resnet_temb)
                resnet_temb = resnet_temb.to(hidden_states.dtype)
            if self.add_spatial_resnet:  # :213-219
                _, _, height, width = hidden_states.shape
                output_size = (int(height * self.up_sampling_scale), int(width * self.up_sampling_scale)) if i == 0 else None
                hidden_states = self.spatial_resnets[i](hidden_states, resnet_temb, output_size=output_size)
                _, _, height, width = hidden_states.shape
                if self.add_resnet_time_mixer:
                    hidden_states_mix = hidden_states[None, :].reshape(batch_size, num_frames, channels, height, width).permute(0, 2, 1, 3, 4)
            if self.add_temporal_resnet:  # :223-231
                hidden_states = hidden_states[None, :].reshape(batch_size, num_frames, channels, height, width).permute(0, 2, 1, 3, 4)
                resnet_temb = resnet_temb.reshape(batch_size, num_frames, -1)
                hidden_states = self.temporal_resnets[i](hidden_states, resnet_temb)
                if self.add_resnet_time_mixer:
                    hidden_states = self.resnets_time_mixer[i](x_spatial=hidden_states_mix, x_temporal=hidden_states,
                                                               image_only_indicator=image_only_indicator)
                hidden_states = hidden_states.permute(0, 2, 1, 3, 4).reshape(batch_frames, channels, height, width)
            if not self.add_spatial_resnet and not self.add_temporal_resnet and i == 0 and self.up_sampling_scale > 1:
                hidden_states = F.interpolate(hidden_states, scale_factor=self.up_sampling_scale, mode="nearest")  # :235-237
                _, _, height, width = hidden_states.shape
            if self.add_spatial_transformer or self.add_temporal_transformer:  # :239-266
                if encoder_hidden_states.dim() == 2:
                    encoder_hidden_states = encoder_hidden_states.unsqueeze(1)
                if encoder_hidden_states.shape[0] == 1:
                    encoder_hidden_states = encoder_hidden_states.repeat_interleave(batch_frames, dim=0)
                if self.add_temporal_transformer:
                    time_context = encoder_hidden_states
                    tc_first = time_context[None, :].reshape(batch_size, num_frames, -1, time_context.shape[-1])[:, 0]
                    time_context = tc_first[None, :].broadcast_to(height * width, batch_size, 1, time_context.shape[-1])
                    time_context = time_context.reshape(height * width * batch_size, 1, time_context.shape[-1])
                residual = hidden_states
                hidden_states = self.norm(hidden_states)
                inner_dim = hidden_states.shape[1]
                hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch_frames, height * width, inner_dim)
                hidden_states = self.proj_in(hidden_states)
                if self.add_temporal_transformer:
                    num_frames_emb = torch.arange(num_frames, device=dev).repeat(batch_size, 1).reshape(-1)
                    t_emb = self.transformer_time_proj(num_frames_emb).to(dtype=hidden_states.dtype)
                    emb = self.transformer_time_embedding(t_emb)[:, None, :]
            if self.add_spatial_transformer:  # :270-273
                hidden_states = self.spatial_attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states)
                if self.add_transformer_time_mixer:
                    hidden_states_mix = hidden_states
            if self.add_temporal_transformer:  # :278-282
                hidden_states = hidden_states + emb
                hidden_states = self.temporal_attentions[i](hidden_states, num_frames=num_frames,
                                                            encoder_hidden_states=time_context)
                if self.add_transformer_time_mixer:
                    hidden_states = self.transformers_time_mixer[i](x_spatial=hidden_states_mix, x_temporal=hidden_states,
                                                                    image_only_indicator=image_only_indicator)
            if self.add_spatial_transformer or self.add_temporal_transformer:  # :286-289
                hidden_states = self.proj_out(hidden_states)
                hidden_states = hidden_states.reshape(batch_frames, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
                hidden_states = hidden_states + residual
        return hidden_states


class ControlNetAdapter(nn.Module):
    """ctrl_adapter.py:12-224 (num_repeats == 1 path; the experimental num_repeats > 1 branch :208-221 is unused by
    the released configs and is not restated)."""

    def __init__(self, backbone_model_name, num_blocks=2, num_frames=8, num_adapters_per_location=3,
                 cross_attention_dim=None, adapter_type="spatial_temporal_resnet_transformer", add_spatial_resnet=True,
                 add_temporal_resnet=False, add_spatial_transformer=True, add_temporal_transformer=False,
                 add_adapter_location_A=False, add_adapter_location_B=False, add_adapter_location_C=False,
                 add_adapter_location_D=False, add_adapter_location_M=False, num_repeats=1, out_channels=None):
        super().__init__()
        assert num_repeats == 1
        self.config = dict(backbone_model_name=backbone_model_name, num_blocks=num_blocks, num_frames=num_frames,
                           num_adapters_per_location=num_adapters_per_location, cross_attention_dim=cross_attention_dim,
                           adapter_type=adapter_type, add_spatial_resnet=add_spatial_resnet,
                           add_temporal_resnet=add_temporal_resnet, add_spatial_transformer=add_spatial_transformer,
                           add_temporal_transformer=add_temporal_transformer,
                           add_adapter_location_A=add_adapter_location_A, add_adapter_location_B=add_adapter_location_B,
                           add_adapter_location_C=add_adapter_location_C, add_adapter_location_D=add_adapter_location_D,
                           add_adapter_location_M=add_adapter_location_M, num_repeats=num_repeats, out_channels=out_channels)
        self.add_adapter_location_A = add_adapter_location_A
        self.add_adapter_location_B = add_adapter_location_B
        self.add_adapter_location_C = add_adapter_location_C
        self.add_adapter_location_D = add_adapter_location_D
        self.add_adapter_location_M = add_adapter_location_M
        self.num_adapters_per_location = num_adapters_per_location
        self.num_repeats = num_repeats
        chans = self.get_down_block_channels()
        ids = self.get_down_block_ids()
        scale = 2 if backbone_model_name in ["sdxl"] else 1  # :61-66
        self.num_adapters = len(chans)
        self.adapter_type = adapter_type
        common = dict(cross_attention_dim=cross_attention_dim, num_layers=num_blocks, up_sampling_scale=scale,
                      add_spatial_resnet=add_spatial_resnet, add_temporal_resnet=add_temporal_resnet,
                      add_spatial_transformer=add_spatial_transformer, add_temporal_transformer=add_temporal_transformer)
        self.down_blocks_adapter = nn.ModuleList(
            [AdapterSpatioTemporal(in_channels=c, out_channels=c, **common) for c in chans])
        assert len(ids) == len(chans)
        self.mid_block_adapter = AdapterSpatioTemporal(in_channels=1280, out_channels=1280, **common) \
            if add_adapter_location_M else None

    def get_down_block_ids(self):  # :119-139
        ids = []
        n = self.num_adapters_per_location
        for flag, m in ((self.add_adapter_location_A, {3: [0, 1, 2], 2: [0, 2], 1: [2]}),
                        (self.add_adapter_location_B, {3: [3, 4, 5], 2: [3, 5], 1: [5]}),
                        (self.add_adapter_location_C, {3: [6, 7, 8], 2: [6, 8], 1: [8]}),
                        (self.add_adapter_location_D, {3: [9, 10, 11], 2: [9, 11], 1: [11]})):
            if flag:
                ids += m.get(n, [])
        return ids

    def get_down_block_channels(self):  # :142-168
        ch = []
        n = self.num_adapters_per_location
        if self.add_adapter_location_A:
            ch = [320] * n
        if self.add_adapter_location_B:
            ch += {3: [320, 640, 640], 2: [320, 640], 1: [640]}[n]
        if self.add_adapter_location_C:
            ch += {3: [640, 1280, 1280], 2: [640, 1280], 1: [1280]}[n]
        if self.add_adapter_location_D:
            ch += [1280] * n
        return ch

    def forward(self, down_block_res_samples, mid_block_res_sample=None, sparsity_masking=None, num_frames=None,
                timestep=None, encoder_hidden_states=None):
        ids = self.get_down_block_ids()
        out = []
        cur = 0
        for i in range(12):  # :180-193
            if i in ids:
                out.append(self.down_blocks_adapter[cur](down_block_res_samples[i], sparsity_masking=sparsity_masking,
                                                         num_frames=num_frames, timestep=timestep,
                                                         encoder_hidden_states=encoder_hidden_states))
                cur += 1
            else:
                out.append(torch.zeros_like(down_block_res_samples[i]))
        if mid_block_res_sample is not None and self.mid_block_adapter is not None:  # :196-205
            mid = self.mid_block_adapter(mid_block_res_sample, sparsity_masking=sparsity_masking, num_frames=num_frames,
                                         timestep=timestep, encoder_hidden_states=encoder_hidden_states)
        else:
            mid = None
        return out, mid


class ControlNetRouter(nn.Module):
    """model/ctrl_router.py:44-112 with SimpleWeights (:26-40) / EqualWeights (:9-22).  The unconditional ``.cuda()``
    of the reference (:21,38, quirk Q16) is replaced by the parameter's device so the oracle runs on CPU."""

    class _Simple(nn.Module):
        def __init__(self, num_experts):
            super().__init__()
            self.num_experts = num_experts
            self.wg = nn.Linear(1, num_experts, bias=False)

        def forward(self, inputs=None):
            one = torch.ones(1, 1, device=self.wg.weight.device, dtype=self.wg.weight.dtype)
            return self.wg(one)

    class _Equal(nn.Module):
        def __init__(self, num_experts):
            super().__init__()
            self.num_experts = num_experts
            self.register_buffer("holder", torch.tensor([1]), persistent=False)

        def forward(self, inputs=None):
            return torch.zeros([self.num_experts], device=self.holder.device).unsqueeze(0)

    def __init__(self, num_experts=2, backbone_model_name=None, router_type="simple_weights", embedding_dim=None,
                 num_routers=12, add_mid_block_router=True, use_sparsemax=False):
        super().__init__()
        self.num_experts = num_experts
        self.num_routers = num_routers
        self.router_type = router_type
        self.add_mid_block_router = add_mid_block_router
        cls = self._Equal if router_type == "equal_weights" else self._Simple
        self.down_blocks_router = nn.ModuleList([cls(num_experts) for _ in range(num_routers)])
        if add_mid_block_router:
            self.mid_block_router = cls(num_experts)

    def forward(self, router_input=None, sparse_mask=None, fixed_weights=None):
        down = [self.down_blocks_router[i]() for i in range(self.num_routers)]
        mid = self.mid_block_router()
        if sparse_mask is not None:  # :96-103
            for i in range(len(sparse_mask)):
                if sparse_mask[i] == 0:
                    mid[0, i] -= 1e6
                    for j in range(len(down)):
                        down[j][0, i] -= 1e6
        down_w = F.softmax(torch.concat(down), dim=-1)
        mid_w = F.softmax(mid, dim=-1)
        if mid_w.dim() == 2:
            mid_w = mid_w.squeeze(0)
        return down_w, mid_w
