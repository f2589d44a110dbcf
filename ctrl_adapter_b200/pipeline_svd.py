"""The Stable-Video-Diffusion denoising hot loop (body of
/root/reference/svd/pipelines/svd_controlnet_adapter_pipeline.py:640-787) on the B200 modules:

  [pool to 64x64] -> ControlNet (timestep from the step index) -> [sparse key-frame gather] -> Ctrl-Adapter (spatial +
  temporal) -> [scatter] -> SVD UNet on cat(latents, image latents) with 5-D residual injection -> per-frame CFG ->
  Euler (v-prediction) update.

Latents live in the reference's own (clip, frame, channel, h, w) order.  Per-step scalars sit in device rows so one
captured CUDA graph serves every step.

STATUS: written after the round's GPU budget was spent; verified on CPU through the op-layer emulation against the
restated reference loop (tests/test_host_emulated_cpu.py); its GPU check is in the xfail-guarded pending group.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from .adapter import as_nchw, to_channels_last_bf16
from .schedulers import EulerKarrasVSchedule

BF16 = torch.bfloat16


class SVDControlNetAdapterLoop:
    def __init__(self, controlnet, adapter, unet, *, num_inference_steps: int = 25, min_guidance_scale: float = 1.0,
                 max_guidance_scale: float = 3.0, controlnet_conditioning_scale: float = 1.0, use_size_512: bool = True,
                 skip_conv_in: bool = False, skip_time_emb: bool = False, sparse_frames: Optional[List[int]] = None):
        self.controlnet, self.adapter, self.unet = controlnet, adapter, unet
        self.min_g, self.max_g = float(min_guidance_scale), float(max_guidance_scale)
        self.cond_scale = float(controlnet_conditioning_scale)
        self.use_size_512 = use_size_512
        self.skip_conv_in, self.skip_time_emb = skip_conv_in, skip_time_emb
        self.sparse_frames = None if sparse_frames is None else [int(k) for k in sparse_frames]
        self.schedule = EulerKarrasVSchedule(num_inference_steps)
        self.num_inference_steps = num_inference_steps
        self._graph = None

    def prepare(self, latents, image_latents, image_embeddings, added_time_ids, controlnet_prompt_embeds, control_images):
        """latents (B,F,4,h,w) unit noise; image_latents (2B,F,4,h,w), zeros for the unconditional half (:225-230,
        :582); image_embeddings (2B,1,1024) negative first; added_time_ids (2B,3); controlnet_prompt_embeds
        (2B*F,77,768); control_images (2B*F,3,H,W) in [0,1]."""
        dev = latents.device
        b, f = latents.shape[:2]
        self.batch, self.frames = b, f
        self.latents = (latents.float() * self.schedule.init_noise_sigma).to(BF16).float().contiguous()
        self.table = torch.from_numpy(self.schedule.table()).to(dev)
        self.cn_table = torch.from_numpy(self.schedule.control_timesteps).to(dev)
        self.row = self.table[0].clone()
        self.cn_t = self.cn_table[0:1].clone()
        first_div = torch.tensor(float((self.schedule.sigmas[0] ** 2 + 1) ** 0.5), device=dev).to(BF16)
        self.model_in = (self.latents.to(BF16) / first_div).to(BF16).contiguous()
        # per-frame guidance: linspace in the latent dtype (:616-621)
        self.guidance = torch.linspace(self.min_g, self.max_g, f, device=dev).to(BF16).float().contiguous()
        self.image_latents = image_latents.to(BF16).contiguous()
        self.image_embeddings = image_embeddings.to(BF16).contiguous()
        self.added_time_ids = added_time_ids.float().contiguous()
        self.cn_embeds = controlnet_prompt_embeds.to(BF16).contiguous()
        self.images = control_images.to(BF16).contiguous()
        self.adapter_ctx = self.image_embeddings[-1].unsqueeze(0).contiguous()  # the LAST sample's embedding (:715)
        self._sparse_rows = None
        if self.sparse_frames is not None:
            if not all(0 <= k < f for k in self.sparse_frames):
                raise ValueError("sparse_frames must index frames of the clip")
            self._sparse_rows = torch.tensor([bb * f + k for bb in range(2 * b) for k in self.sparse_frames], device=dev)
        self._graph = None
        self.step_index = 0

    def _body(self):
        b, f = self.batch, self.frames
        t = self.row[0:1]
        lat2 = torch.cat([self.model_in, self.model_in], dim=0)                  # (2B, F, 4, h, w) CFG duplication
        n = 2 * b * f
        h, w = lat2.shape[-2:]
        ctrl_in = lat2.reshape(n, 4, h, w)                                       # "b f c h w -> (b f) c h w"
        images = self.images
        if (h, w) != (64, 64) and self.use_size_512:                             # :664-670
            ctrl_in = as_nchw(ops.avgpool(to_channels_last_bf16(ctrl_in, 8), 64, 64))[:, :4]
            if images.shape[-2:] != (512, 512):
                images = as_nchw(ops.avgpool(to_channels_last_bf16(images, 8), 512, 512))[:, :3]
        down, mid = self.controlnet(ctrl_in, self.cn_t, encoder_hidden_states=self.cn_embeds, controlnet_cond=images,
                                    conditioning_scale=self.cond_scale, guess_mode=False, return_dict=False,
                                    skip_conv_in=self.skip_conv_in, skip_time_emb=self.skip_time_emb)
        if self._sparse_rows is None:
            down_a, mid_a = self.adapter(down, mid_block_res_sample=mid, sparsity_masking=None, num_frames=f,
                                         timestep=self.cn_t, encoder_hidden_states=self.adapter_ctx)
        else:
            rows = self._sparse_rows
            down_k, mid_k = self.adapter([d.index_select(0, rows) for d in down],
                                         mid_block_res_sample=mid.index_select(0, rows),
                                         sparsity_masking=self.sparse_frames, num_frames=len(self.sparse_frames),
                                         timestep=self.cn_t, encoder_hidden_states=self.adapter_ctx)

            def densify(x):
                full = torch.zeros((n, *x.shape[1:]), device=x.device, dtype=x.dtype).contiguous(
                    memory_format=torch.channels_last)
                full.index_copy_(0, rows, x)
                return full
            down_a = [densify(d) for d in down_k]
            mid_a = densify(mid_k) if mid_k is not None else None
        # the UNet takes "(b f) c h w" residuals as well as the reference's 5-D form; no rearrange needed
        residuals = None if self.cond_scale == 0 else down_a
        unet_in = torch.cat([lat2, self.image_latents], dim=2)                   # (2B, F, 8, h, w) channel concat (:755)
        eps = self.unet(unet_in, t, encoder_hidden_states=self.image_embeddings, added_time_ids=self.added_time_ids,
                        down_block_additional_residuals=residuals, mid_block_additional_residual=mid_a,
                        return_dict=False)[0]                                    # (2B, F, 4, h, w)
        ops.cfg_euler_v(eps[:b].contiguous(), eps[b:].contiguous(), self.latents, self.guidance, f, self.row,
                        latents_out=self.latents, model_in_next=self.model_in)

    @torch.no_grad()
    def step(self, i: Optional[int] = None):
        i = self.step_index if i is None else i
        self.row.copy_(self.table[i])
        self.cn_t.copy_(self.cn_table[i:i + 1])
        self._body()
        self.step_index = i + 1
        return self.latents

    @torch.no_grad()
    def capture(self, warmup: int = 2):
        saved = (self.latents.clone(), self.model_in.clone(), self.row.clone(), self.cn_t.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body()
        self._graph = g
        self.latents.copy_(saved[0]); self.model_in.copy_(saved[1]); self.row.copy_(saved[2]); self.cn_t.copy_(saved[3])
        return g

    @torch.no_grad()
    def step_graph(self, i: Optional[int] = None):
        if self._graph is None:
            self.capture()
        i = self.step_index if i is None else i
        self.row.copy_(self.table[i])
        self.cn_t.copy_(self.cn_table[i:i + 1])
        self._graph.replay()
        self.step_index = i + 1
        return self.latents
