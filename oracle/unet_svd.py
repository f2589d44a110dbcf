"""ORACLE (test infrastructure): Stable-Video-Diffusion UNet restated on oracle.blocks.

Follows /root/reference/svd/models/unets/unet_spatio_temporal_condition.py: UNetSpatioTemporalConditionModel
constructor (:71-245) and forward (:357-526) including the reference's residual-injection additions
(down residuals :457-471 -- 5-D "b c f h w" lists are flattened to (b f) c h w and zip() truncates to the shorter
list; mid residual :485-490).  Pinned against the reference's own class (run through oracle/diffusers_shim) by
tests/golden/make_golden.py.  Not imported by the product package.
"""
from __future__ import annotations

import torch
from torch import nn

from .blocks import TimestepEmbedding, Timesteps, UNetMidBlockSpatioTemporal, get_down_block_3d, get_up_block_3d


class _Config(dict):
    __getattr__ = dict.__getitem__


class UNetSpatioTemporalConditionModel(nn.Module):
    def __init__(self, sample_size=None, in_channels: int = 8, out_channels: int = 4,
                 down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                                   "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
                 up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                                 "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
                 block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim: int = 256,
                 projection_class_embeddings_input_dim: int = 768, layers_per_block=2, cross_attention_dim=1024,
                 transformer_layers_per_block=1, num_attention_heads=(5, 10, 10, 20), num_frames: int = 25):
        super().__init__()
        n = len(down_block_types)
        if len(up_block_types) != n or len(block_out_channels) != n:
            raise ValueError("down_block_types, up_block_types and block_out_channels must have the same length")  # :103-112
        self.config = _Config(in_channels=in_channels, out_channels=out_channels, num_frames=num_frames,
                              block_out_channels=tuple(block_out_channels), cross_attention_dim=cross_attention_dim)
        c0 = block_out_channels[0]
        self.conv_in = nn.Conv2d(in_channels, c0, kernel_size=3, padding=1)                     # :128-133
        time_embed_dim = c0 * 4
        self.time_proj = Timesteps(c0, True, downscale_freq_shift=0)                           # :138
        self.time_embedding = TimestepEmbedding(c0, time_embed_dim)
        self.add_time_proj = Timesteps(addition_time_embed_dim, True, downscale_freq_shift=0)  # :143
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, time_embed_dim)

        heads = (num_attention_heads,) * n if isinstance(num_attention_heads, int) else tuple(num_attention_heads)
        xdim = (cross_attention_dim,) * n if isinstance(cross_attention_dim, int) else tuple(cross_attention_dim)
        lpb = [layers_per_block] * n if isinstance(layers_per_block, int) else list(layers_per_block)
        tlpb = ([transformer_layers_per_block] * n if isinstance(transformer_layers_per_block, int)
                else list(transformer_layers_per_block))

        self.down_blocks = nn.ModuleList()
        output_channel = c0
        for i, down_block_type in enumerate(down_block_types):                                 # :165-184
            input_channel, output_channel = output_channel, block_out_channels[i]
            self.down_blocks.append(get_down_block_3d(
                down_block_type, num_layers=lpb[i], transformer_layers_per_block=tlpb[i], in_channels=input_channel,
                out_channels=output_channel, temb_channels=time_embed_dim, add_downsample=i != n - 1, resnet_eps=1e-5,
                cross_attention_dim=xdim[i], num_attention_heads=heads[i], resnet_act_fn="silu"))

        self.mid_block = UNetMidBlockSpatioTemporal(                                            # :187-193
            block_out_channels[-1], temb_channels=time_embed_dim, transformer_layers_per_block=tlpb[-1],
            cross_attention_dim=xdim[-1], num_attention_heads=heads[-1])

        self.num_upsamplers = 0
        self.up_blocks = nn.ModuleList()
        rev_ch, rev_heads, rev_lpb = list(reversed(block_out_channels)), list(reversed(heads)), list(reversed(lpb))
        rev_xdim, rev_tlpb = list(reversed(xdim)), list(reversed(tlpb))
        output_channel = rev_ch[0]
        for i, up_block_type in enumerate(up_block_types):                                     # :206-236
            is_final = i == n - 1
            prev_output_channel, output_channel = output_channel, rev_ch[i]
            input_channel = rev_ch[min(i + 1, n - 1)]
            if not is_final:
                self.num_upsamplers += 1
            self.up_blocks.append(get_up_block_3d(
                up_block_type, num_layers=rev_lpb[i] + 1, transformer_layers_per_block=rev_tlpb[i],
                in_channels=input_channel, out_channels=output_channel, prev_output_channel=prev_output_channel,
                temb_channels=time_embed_dim, add_upsample=not is_final, resnet_eps=1e-5, resolution_idx=i,
                cross_attention_dim=rev_xdim[i], num_attention_heads=rev_heads[i], resnet_act_fn="silu"))

        self.conv_norm_out = nn.GroupNorm(num_channels=c0, num_groups=32, eps=1e-5)           # :239
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c0, out_channels, kernel_size=3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict: bool = False):
        # 1. time (:393-420)
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        batch_size, num_frames = sample.shape[:2]
        timesteps = timesteps.expand(batch_size)
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=sample.dtype))
        time_embeds = self.add_time_proj(added_time_ids.flatten()).reshape((batch_size, -1)).to(emb.dtype)
        emb = emb + self.add_embedding(time_embeds)

        # flatten batch and frames; per-frame copies of the embeddings / context (:422-430)
        sample = sample.flatten(0, 1)
        emb = emb.repeat_interleave(num_frames, dim=0)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(num_frames, dim=0)

        sample = self.conv_in(sample)                                                           # :433
        image_only_indicator = torch.zeros(batch_size, num_frames, dtype=sample.dtype, device=sample.device)

        # 3. down (:438-455)
        down_block_res_samples = (sample,)
        for block in self.down_blocks:
            if getattr(block, "has_cross_attention", False):
                sample, res_samples = block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                            image_only_indicator=image_only_indicator)
            else:
                sample, res_samples = block(hidden_states=sample, temb=emb, image_only_indicator=image_only_indicator)
            down_block_res_samples += res_samples

        # reference addition: ControlNet / adapter residuals on the skip connections (:457-471)
        if down_block_additional_residuals is not None:
            if down_block_additional_residuals[0].dim() == 5:
                down_block_additional_residuals = [r.permute(0, 2, 1, 3, 4).flatten(0, 1)
                                                   for r in down_block_additional_residuals]
            down_block_res_samples = tuple(s + r for s, r in zip(down_block_res_samples, down_block_additional_residuals))

        # 4. mid (:477-490)
        sample = self.mid_block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                image_only_indicator=image_only_indicator)
        if mid_block_additional_residual is not None:
            if torch.is_tensor(mid_block_additional_residual) and mid_block_additional_residual.dim() == 5:
                mid_block_additional_residual = mid_block_additional_residual.permute(0, 2, 1, 3, 4).flatten(0, 1)
            sample = sample + mid_block_additional_residual

        # 5. up (:495-514)
        for block in self.up_blocks:
            res_samples = down_block_res_samples[-len(block.resnets):]
            down_block_res_samples = down_block_res_samples[:-len(block.resnets)]
            if getattr(block, "has_cross_attention", False):
                sample = block(hidden_states=sample, temb=emb, res_hidden_states_tuple=res_samples,
                               encoder_hidden_states=encoder_hidden_states, image_only_indicator=image_only_indicator)
            else:
                sample = block(hidden_states=sample, temb=emb, res_hidden_states_tuple=res_samples,
                               image_only_indicator=image_only_indicator)

        # 6. post-process and un-flatten (:517-526)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        sample = sample.reshape(batch_size, num_frames, *sample.shape[1:])
        return (sample,)
