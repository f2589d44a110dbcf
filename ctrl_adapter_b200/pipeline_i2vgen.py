"""The I2VGen-XL denoising hot loop (body of /root/reference/i2vgen_xl/pipelines/i2vgen_xl_controlnet_adapter_pipeline.py
:902-1118) on the B200 modules:

  ControlNet (or E ControlNets -> router -> weighted merge) -> Ctrl-Adapter (spatial + temporal) -> I2VGen-XL UNet with
  residual injection -> CFG -> DDIM update.

Latents are kept in (clip, frame, channel, h, w) order internally, i.e. already in the "(b f) c h w" layout the
reference permutes to around the scheduler step (:1107-1115); `latents_bcfhw()` gives the reference's view.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from .adapter import router_merge
from .schedulers import DDIMSchedule

BF16 = torch.bfloat16


class I2VGenXLControlNetAdapterLoop:
    def __init__(self, controlnet, adapter, unet, router=None, *, num_inference_steps: int = 50,
                 guidance_scale: float = 9.0, controlnet_conditioning_scale: float = 1.0,
                 inference_expert_masks: Optional[List[bool]] = None, skip_conv_in: bool = False,
                 skip_time_emb: bool = False, sparse_frames: Optional[List[int]] = None):
        self.controlnet, self.adapter, self.unet, self.router = controlnet, adapter, unet, router
        self.guidance_scale = float(guidance_scale)
        self.cond_scale = float(controlnet_conditioning_scale)
        self.masks = inference_expert_masks
        self.skip_conv_in, self.skip_time_emb = skip_conv_in, skip_time_emb
        # sparse control (:1024-1033, :1053-1073): only these key frames of every clip go through the adapter; the other
        # frames get zero residuals
        self.sparse_frames = None if sparse_frames is None else [int(k) for k in sparse_frames]
        self.schedule = DDIMSchedule(num_inference_steps)
        self.num_inference_steps = num_inference_steps
        self._graph = None

    def prepare(self, latents, prompt_embeds, image_latents, image_embeddings, fps, controlnet_prompt_embeds,
                control_images):
        """latents (B,4,F,h,w) unit noise; prompt_embeds (2B,77,1024) negative first; image_latents (2B,4,F,h,w);
        image_embeddings (2B,1,1024); fps (2B,); controlnet_prompt_embeds (2B*F,77,768); control_images: (2B*F,3,512,512)
        or a list of E such tensors (Multi-ControlNet)."""
        dev = latents.device
        b, c, f, h, w = latents.shape
        self.batch, self.frames = b, f
        lat = latents.float() * self.schedule.init_noise_sigma
        self.latents = lat.permute(0, 2, 1, 3, 4).contiguous().to(BF16).float()          # (B, F, 4, h, w) fp32 master
        self.model_in = self.latents.to(BF16).contiguous()
        self.table = torch.from_numpy(self.schedule.table()).to(dev)
        self.row = self.table[0].clone()
        self.prompt_embeds = prompt_embeds.to(BF16).contiguous()
        self.image_latents = image_latents.to(BF16).contiguous()
        self.image_embeddings = image_embeddings.to(BF16).contiguous()
        self.fps = fps.float().contiguous()
        self.cn_embeds = controlnet_prompt_embeds.to(BF16).contiguous()
        self.images = ([i.to(BF16).contiguous() for i in control_images] if isinstance(control_images, (list, tuple))
                       else control_images.to(BF16).contiguous())
        # adapter context: the LAST sample's image embedding for every sample (reference :1048, quirk Q5)
        self.adapter_ctx = self.image_embeddings[-1].unsqueeze(0).contiguous()
        # timestep-independent UNet conditioning, computed once (exact hoist)
        self.unet_cond = self.unet.prepare_conditioning(self.fps, self.image_latents, self.image_embeddings,
                                                        self.prompt_embeds)
        self.unet._cond_cache = (tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in
                                       (self.fps, self.image_latents, self.image_embeddings, self.prompt_embeds)),
                                 self.unet_cond)
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
        ctrl_in = lat2.reshape(n, *lat2.shape[2:])                               # "(b f) c h w"
        multi = isinstance(self.images, list)
        scale = [self.cond_scale] * len(self.images) if multi else self.cond_scale
        down, mid = self.controlnet(ctrl_in, t, encoder_hidden_states=self.cn_embeds, controlnet_cond=self.images,
                                    conditioning_scale=scale, guess_mode=False, return_dict=False,
                                    skip_conv_in=self.skip_conv_in, skip_time_emb=self.skip_time_emb)
        if self.router is not None:
            dw, mw = self.router(sparse_mask=self.masks)                         # (12, E), (E,)
            active = [e for e in range(self.router.num_experts) if self.masks[e]]
            # down[idx_e][k]: output k of the idx_e-th ControlNet that ran; weights indexed by expert id (Q6 applies)
            lists = {e: i for i, e in enumerate(active)}
            down = [router_merge({e: down[lists[e]][k] for e in active}, dw[k], active, f)
                    for k in range(self.router.num_routers)]
            mid = router_merge({e: mid[lists[e]] for e in active}, mw, active, f) if mw is not None else None
            from .adapter import as_nchw
            down = [as_nchw(d) for d in down]
            mid = as_nchw(mid) if mid is not None else None
        if self._sparse_rows is None:
            down_a, mid_a = self.adapter(down, mid_block_res_sample=mid, sparsity_masking=None, num_frames=f, timestep=t,
                                         encoder_hidden_states=self.adapter_ctx)
        else:
            # key-frame gather -> adapter on len(sparse_frames) frames per clip -> scatter into zero residuals
            # (host-side index glue over ControlNet-sized tensors; the reference's dense tensors are fp32 zeros, which
            # only changes where the skip + residual sum is rounded -- quirk Q21)
            rows = self._sparse_rows
            down_s = [d.index_select(0, rows) for d in down]
            mid_s = mid.index_select(0, rows) if mid is not None else None
            down_k, mid_k = self.adapter(down_s, mid_block_res_sample=mid_s, sparsity_masking=self.sparse_frames,
                                         num_frames=len(self.sparse_frames), timestep=t,
                                         encoder_hidden_states=self.adapter_ctx)

            def densify(x):
                full = torch.zeros((n, *x.shape[1:]), device=x.device, dtype=x.dtype).contiguous(
                    memory_format=torch.channels_last)
                full.index_copy_(0, rows, x)
                return full
            down_a = [densify(d) for d in down_k]
            mid_a = densify(mid_k) if mid_k is not None else None
        residuals = None if self.cond_scale == 0 else down_a                     # mid is still injected (quirk Q9)
        sample = lat2.permute(0, 2, 1, 3, 4)                                     # (2B, 4, F, h, w) view
        eps = self.unet(sample, t, self.fps, self.image_latents, image_embeddings=self.image_embeddings,
                        encoder_hidden_states=self.prompt_embeds, down_block_additional_residuals=residuals,
                        mid_block_additional_residual=mid_a, return_dict=False)[0]
        eps = eps.permute(0, 2, 1, 3, 4)                                          # back to (2B, F, 4, h, w): contiguous
        ops.cfg_ddim(eps[:b].contiguous(), eps[b:].contiguous(), self.latents, self.guidance_scale, self.row,
                     latents_out=self.latents, model_in_next=self.model_in, v_prediction=self.schedule.v_prediction)

    @torch.no_grad()
    def step(self, i: Optional[int] = None):
        i = self.step_index if i is None else i
        self.row.copy_(self.table[i])
        self._body()
        self.step_index = i + 1
        return self.latents

    @torch.no_grad()
    def capture(self, warmup: int = 2):
        saved = (self.latents.clone(), self.model_in.clone(), self.row.clone())
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
        self.latents.copy_(saved[0]); self.model_in.copy_(saved[1]); self.row.copy_(saved[2])
        return g

    @torch.no_grad()
    def step_graph(self, i: Optional[int] = None):
        if self._graph is None:
            self.capture()
        i = self.step_index if i is None else i
        self.row.copy_(self.table[i])
        self._graph.replay()
        self.step_index = i + 1
        return self.latents

    def latents_bcfhw(self):
        return self.latents.permute(0, 2, 1, 3, 4)
