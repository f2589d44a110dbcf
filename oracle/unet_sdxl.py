"""ORACLE (test infrastructure): diffusers v0.27.2 ``UNet2DConditionModel`` restated for the SDXL-base configuration.

The reference calls the stock diffusers class (instantiated at /root/reference/inference.py:369, called at
/root/reference/sdxl/pipelines/sdxl_controlnet_adapter_pipeline.py:1356-1366 with ``down_block_additional_residuals``
and ``mid_block_additional_residual=0``); the source is not vendored in the reference ("parity unpinned" layer, see
oracle/blocks.py).  Configuration = stabilityai/stable-diffusion-xl-base-1.0 unet/config.json.
Not imported by the product package.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .blocks import TimestepEmbedding, Timesteps, UNetMidBlock2DCrossAttn, get_down_block, get_up_block

SDXL_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280),
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    layers_per_block=2, transformer_layers_per_block=(1, 2, 10), attention_head_dim=(5, 10, 20),
    cross_attention_dim=2048, use_linear_projection=True, addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=2816, norm_num_groups=32, norm_eps=1e-5,
)


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280),
                 down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
                 up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"), layers_per_block=2,
                 transformer_layers_per_block=(1, 2, 10), attention_head_dim=(5, 10, 20), cross_attention_dim=2048,
                 use_linear_projection=True, addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816,
                 norm_num_groups=32, norm_eps=1e-5, flip_sin_to_cos=True, freq_shift=0):
        super().__init__()
        n = len(block_out_channels)
        num_attention_heads = tuple(attention_head_dim)  # diffusers naming quirk: attention_head_dim holds head counts
        tl = list(transformer_layers_per_block) if not isinstance(transformer_layers_per_block, int) else [transformer_layers_per_block] * n
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        time_embed_dim = block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim)
        self.add_time_proj = Timesteps(addition_time_embed_dim, flip_sin_to_cos, freq_shift)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, time_embed_dim)
        self.down_blocks = nn.ModuleList([])
        output_channel = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            input_channel = output_channel
            output_channel = block_out_channels[i]
            is_final = i == n - 1
            self.down_blocks.append(get_down_block(
                t, num_layers=layers_per_block, transformer_layers_per_block=tl[i], in_channels=input_channel,
                out_channels=output_channel, temb_channels=time_embed_dim, add_downsample=not is_final,
                resnet_eps=norm_eps, resnet_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                num_attention_heads=num_attention_heads[i], downsample_padding=1,
                use_linear_projection=use_linear_projection))
        self.mid_block = UNetMidBlock2DCrossAttn(
            transformer_layers_per_block=tl[-1], in_channels=block_out_channels[-1], temb_channels=time_embed_dim,
            resnet_eps=norm_eps, cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads[-1],
            resnet_groups=norm_num_groups, use_linear_projection=use_linear_projection)
        self.up_blocks = nn.ModuleList([])
        rev_ch = list(reversed(block_out_channels))
        rev_heads = list(reversed(num_attention_heads))
        rev_tl = list(reversed(tl))
        output_channel = rev_ch[0]
        for i, t in enumerate(up_block_types):
            is_final = i == n - 1
            prev_output_channel = output_channel
            output_channel = rev_ch[i]
            input_channel = rev_ch[min(i + 1, n - 1)]
            self.up_blocks.append(get_up_block(
                t, num_layers=layers_per_block + 1, transformer_layers_per_block=rev_tl[i], in_channels=input_channel,
                out_channels=output_channel, prev_output_channel=prev_output_channel, temb_channels=time_embed_dim,
                add_upsample=not is_final, resnet_eps=norm_eps, resnet_groups=norm_num_groups,
                cross_attention_dim=cross_attention_dim, num_attention_heads=rev_heads[i],
                use_linear_projection=use_linear_projection))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=3, padding=1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs, down_block_additional_residuals=None,
                mid_block_additional_residual=None, timestep_cond=None, cross_attention_kwargs=None, return_dict=False):
        # upsample-size forwarding when the resolution is not a multiple of 2**(#upsamplers)
        default_overall_up_factor = 2 ** (len(self.up_blocks) - 1)
        forward_upsample_size = any(s % default_overall_up_factor != 0 for s in sample.shape[-2:])
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb, timestep_cond)
        text_embeds = added_cond_kwargs["text_embeds"]
        time_ids = added_cond_kwargs["time_ids"]
        time_embeds = self.add_time_proj(time_ids.flatten()).reshape((text_embeds.shape[0], -1))
        add_embeds = torch.concat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
        emb = emb + self.add_embedding(add_embeds)

        sample = self.conv_in(sample)
        is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
        down_block_res_samples = (sample,)
        for block in self.down_blocks:
            if getattr(block, "has_cross_attention", False):
                sample, res = block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample, res = block(hidden_states=sample, temb=emb)
            down_block_res_samples += res
        if is_controlnet:  # zip truncates the 12-long adapter list to the 9 SDXL skips (quirk Q8)
            new = ()
            for r, a in zip(down_block_res_samples, down_block_additional_residuals):
                new = new + (r + a,)
            down_block_res_samples = new
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        if is_controlnet:
            sample = sample + mid_block_additional_residual  # SDXL pipeline passes the python int 0 (quirk Q9)
        for i, block in enumerate(self.up_blocks):
            is_final = i == len(self.up_blocks) - 1
            res = down_block_res_samples[-len(block.resnets):]
            down_block_res_samples = down_block_res_samples[: -len(block.resnets)]
            upsample_size = down_block_res_samples[-1].shape[2:] if (not is_final and forward_upsample_size) else None
            if getattr(block, "has_cross_attention", False):
                sample = block(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                               encoder_hidden_states=encoder_hidden_states, upsample_size=upsample_size)
            else:
                sample = block(hidden_states=sample, temb=emb, res_hidden_states_tuple=res, upsample_size=upsample_size)
        sample = self.conv_norm_out(sample)
        sample = self.conv_act(sample)
        sample = self.conv_out(sample)
        return (sample,)
