"""ORACLE (test infrastructure): the I2VGen-XL denoising loop body of the reference pipeline restated on the oracle
modules.  Follows /root/reference/i2vgen_xl/pipelines/i2vgen_xl_controlnet_adapter_pipeline.py:902-1118 (CFG on, no
guess mode, dense or sparse key frames) with diffusers v0.27.2 DDIMScheduler restated below.  The I2VGen-XL scheduler_config.json is
not part of the reference repository; its values (squaredcos_cap_v2, rescale_betas_zero_snr, v_prediction,
set_alpha_to_one, leading spacing, steps_offset 1) are restated from the published model card ("parity unpinned").
Not imported by the product package."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, steps_offset=1, set_alpha_to_one=True, prediction_type="v_prediction"):
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        betas = torch.tensor([min(1 - bar((i + 1) / num_train_timesteps) / bar(i / num_train_timesteps), 0.999)
                              for i in range(num_train_timesteps)], dtype=torch.float32)
        # rescale_zero_terminal_snr
        alphas = 1.0 - betas
        sq = alphas.cumprod(0).sqrt()
        s0, sT = sq[0].clone(), sq[-1].clone()
        sq = (sq - sT) * (s0 / (s0 - sT))
        ab = sq ** 2
        alphas = torch.cat([ab[0:1], ab[1:] / ab[:-1]])
        self.alphas_cumprod = torch.cumprod(alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps, self.steps_offset, self.prediction_type = num_train_timesteps, steps_offset, prediction_type
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n, device="cpu"):
        self.num_inference_steps = n
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, t):
        return sample

    def step(self, model_output, timestep, sample):
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        else:  # v_prediction
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        direction = (1 - a_prev) ** 0.5 * eps  # eta = 0
        return a_prev ** 0.5 * x0 + direction


@torch.no_grad()
def i2vgen_step(controlnet, adapter, unet, scheduler, i, latents, prompt_embeds, image_latents, image_embeddings, fps,
                controlnet_prompt_embeds, images, router=None, masks=None, guidance_scale=9.0, cond_scale=1.0,
                sparse_frames=None, use_size_512=True):
    """latents (B,4,F,h,w).  One iteration of :902-1115.  sparse_frames: key-frame indices in [0, F) (:1024-1033,
    :1053-1073); the reference supports one clip with its CFG pair (index s of the conditional half is s + F), restated
    here for 2B clips as b*F + s."""
    t = scheduler.timesteps[i]
    b, c, f, h, w = latents.shape
    latent_model_input = scheduler.scale_model_input(torch.cat([latents] * 2), t)                  # :904-905
    control_in = latent_model_input.permute(0, 2, 1, 3, 4).reshape(2 * b * f, c, h, w)             # :931
    multi = isinstance(images, (list, tuple))
    scale = list(cond_scale) if isinstance(cond_scale, (list, tuple)) else ([cond_scale] * len(images) if multi else cond_scale)
    if (h, w) != (64, 64) and use_size_512:                                                         # :941-947
        control_in = F.adaptive_avg_pool2d(control_in, (64, 64))
        images = ([F.adaptive_avg_pool2d(im, (512, 512)) for im in images] if multi
                  else F.adaptive_avg_pool2d(images, (512, 512)))
    down, mid = controlnet(control_in, t, encoder_hidden_states=controlnet_prompt_embeds, controlnet_cond=images,
                           conditioning_scale=scale, guess_mode=False, return_dict=False)          # :957-968
    if router is not None:                                                                          # :972-1022
        dw, mw = router(sparse_mask=masks)
        E, R = router.num_experts, router.num_routers
        mid_m, idx_e = 0, 0
        for e in range(E):
            if masks[e]:
                mid_m = mid_m + mid[idx_e] * mw.repeat_interleave(f, dim=0)[e]
                idx_e += 1
        down_m = [0 for _ in range(R)]
        for k in range(R):
            idx_e = 0
            for e in range(E):
                if masks[e]:
                    down_m[k] = down_m[k] + down[idx_e][k] * dw[k].repeat_interleave(f, dim=0)[e]
                    idx_e += 1
        down, mid = down_m, mid_m
    n_adapter_frames = f
    if sparse_frames is not None:                                                                   # :1026-1033
        sparse_frames = [int(k) for k in sparse_frames]
        rows = [bb * f + k for bb in range(2 * b) for k in sparse_frames]  # == sparse + [k + F ...] for one clip
        down = [d[rows, :] for d in down]
        mid = mid[rows, :]
        n_adapter_frames = len(sparse_frames)
    a_down, a_mid = adapter(down_block_res_samples=[d.to(latents.dtype) for d in down],
                            mid_block_res_sample=mid.to(latents.dtype), sparsity_masking=sparse_frames,
                            num_frames=n_adapter_frames, timestep=t,
                            encoder_hidden_states=image_embeddings[-1].unsqueeze(0))                # :1042-1049
    if sparse_frames is not None:                                                                   # :1053-1073
        # dense tensors are created by torch.zeros(...) -> float32 whatever the adapter dtype (reference quirk Q21)
        def densify(x):
            full = torch.zeros((2 * b * f, *x.shape[1:]), device=x.device)
            for j, pos in enumerate(rows):
                full[pos] = x[j]
            return full
        a_down = [densify(d) for d in a_down]
        a_mid = densify(a_mid) if a_mid is not None else None
    # "(bs nf) c h w -> bs c nf h w" (the reference hard-codes bs=2; generalised to 2B clips)      # :1080-1083
    re5 = lambda x: x.reshape(2 * b, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)  # noqa: E731
    a_mid5 = re5(a_mid) if a_mid is not None else None
    a_down5 = None if cond_scale == 0 else [re5(d) for d in a_down]
    noise_pred = unet(latent_model_input, t, fps, image_latents, image_embeddings=image_embeddings,
                      encoder_hidden_states=prompt_embeds, down_block_additional_residuals=a_down5,
                      mid_block_additional_residual=a_mid5, return_dict=False)[0]                   # :1088-1099
    u, cnd = noise_pred.chunk(2)
    noise_pred = u + guidance_scale * (cnd - u)                                                     # :1102-1104
    lat2d = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)                                 # :1107-1109
    np2d = noise_pred.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    lat2d = scheduler.step(np2d, t, lat2d)                                                          # :1112
    return lat2d[None, :].reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)                            # :1115
