"""ORACLE (test infrastructure): the Stable-Video-Diffusion denoising loop body of the reference pipeline restated on the
oracle modules.  Follows /root/reference/svd/pipelines/svd_controlnet_adapter_pipeline.py:640-787 (CFG on, no guess
mode, `fixed_controlnet_timestep < 0`, dense or sparse key frames) with the diffusers v0.27.2 EulerDiscreteScheduler
restated below.  The SVD scheduler_config.json is not part of the reference repository; its values (scaled_linear betas
0.00085-0.012, v_prediction, continuous timesteps, Karras sigmas with sigma_min 0.002 / sigma_max 700, leading spacing,
steps_offset 1) are restated from the published model card ("parity unpinned").  Not imported by the product package."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class EulerDiscreteSchedulerSVD:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1, sigma_min=0.002,
                 sigma_max=700.0):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    def set_timesteps(self, n, device="cpu"):
        # the leading-spaced / interpolated sigmas are computed by diffusers and then REPLACED by the Karras ramp between
        # the configured sigma_min / sigma_max (rho = 7)
        rho = 7.0
        ramp = np.linspace(0, 1, n)
        min_inv_rho, max_inv_rho = self.sigma_min ** (1 / rho), self.sigma_max ** (1 / rho)
        sigmas = torch.from_numpy((max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho).to(dtype=torch.float32)
        # timestep_type "continuous" + v_prediction: t = 0.25 * log(sigma)
        self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas]).to(device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1)]).to(device)

    @property
    def init_noise_sigma(self):
        return (self.sigmas.max() ** 2 + 1) ** 0.5  # "leading" spacing

    def scale_model_input(self, sample, i):
        sigma = self.sigmas[i]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, i, sample):
        sample = sample.to(torch.float32)  # diffusers up-casts to avoid precision issues
        sigma = self.sigmas[i]
        # v_prediction: denoised = model_output * c_out + input * c_skip
        pred_original_sample = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        derivative = (sample - pred_original_sample) / sigma  # gamma = 0 -> sigma_hat = sigma
        dt = self.sigmas[i + 1] - sigma
        prev_sample = sample + derivative * dt
        return prev_sample.to(model_output.dtype)


@torch.no_grad()
def svd_step(controlnet, adapter, unet, scheduler, i, latents, image_latents, image_embeddings, added_time_ids,
             controlnet_prompt_embeds, images, min_guidance_scale=1.0, max_guidance_scale=3.0, cond_scale=1.0,
             use_size_512=True, skip_conv_in=False, skip_time_emb=False, sparse_frames=None, num_inference_steps=None):
    """latents (B,F,4,h,w); image_latents (2B,F,4,h,w) (zeros for the unconditional half); image_embeddings (2B,1,1024);
    added_time_ids (2B,3); controlnet_prompt_embeds (2B*F,77,768); images (2B*F,3,H,W).  One iteration of :644-787."""
    t = scheduler.timesteps[i]
    b, f, c, h, w = latents.shape
    n_steps = len(scheduler.timesteps) if num_inference_steps is None else num_inference_steps
    latent_model_input = scheduler.scale_model_input(torch.cat([latents] * 2), i)                    # :646-647
    control_model_input = latent_model_input.reshape(2 * b * f, c, h, w)                              # :660
    if (h, w) != (64, 64) and use_size_512:                                                           # :664-670
        reshaped_in = F.adaptive_avg_pool2d(control_model_input, (64, 64))
        reshaped_images = F.adaptive_avg_pool2d(images, (512, 512))
    else:
        reshaped_in, reshaped_images = control_model_input, images
    # ControlNet / adapter timestep from the step index, not from the continuous SVD timestep (:676-681)
    timestep_interval = 1000 // n_steps
    controlnet_timesteps = torch.Tensor([1000 - (i + 1) * timestep_interval + 1]).round().to(latents.device)
    down, mid = controlnet(reshaped_in, controlnet_timesteps, encoder_hidden_states=controlnet_prompt_embeds,
                           controlnet_cond=reshaped_images, conditioning_scale=cond_scale, guess_mode=False,
                           return_dict=False, skip_conv_in=skip_conv_in, skip_time_emb=skip_time_emb)  # :684-694
    n_adapter_frames = f
    rows = None
    if sparse_frames is not None:                                                                     # :698-704
        sparse_frames = [int(k) for k in sparse_frames]
        rows = [bb * f + k for bb in range(2 * b) for k in sparse_frames]
        down = [d[rows, :] for d in down]
        mid = mid[rows, :]
        n_adapter_frames = len(sparse_frames)
    a_down, a_mid = adapter(down_block_res_samples=[d.to(latents.dtype) for d in down],
                            mid_block_res_sample=mid.to(latents.dtype), sparsity_masking=sparse_frames,
                            num_frames=n_adapter_frames, timestep=controlnet_timesteps,
                            encoder_hidden_states=image_embeddings[-1].unsqueeze(0))                  # :709-715
    if rows is not None:                                                                              # :719-739
        def densify(x):
            full = torch.zeros((2 * b * f, *x.shape[1:]), device=x.device)  # fp32 zeros (quirk Q21)
            for j, pos in enumerate(rows):
                full[pos] = x[j]
            return full
        a_down = [densify(d) for d in a_down]
        a_mid = densify(a_mid) if a_mid is not None else None
    # "(bs nf) c h w -> bs c nf h w" (the reference hard-codes bs=2; generalised to 2B clips)        # :746-747
    re5 = lambda x: x.reshape(2 * b, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)  # noqa: E731
    a_mid5 = re5(a_mid)
    a_down5 = None if cond_scale == 0 else [re5(d) for d in a_down]                                   # :748-749
    unet_in = torch.cat([latent_model_input, image_latents], dim=2)                                   # :755
    noise_pred = unet(unet_in, t, encoder_hidden_states=image_embeddings, added_time_ids=added_time_ids,
                      down_block_additional_residuals=a_down5, mid_block_additional_residual=a_mid5,
                      return_dict=False)[0]                                                           # :758-766
    # per-frame guidance scale, linspace in the latent dtype (:616-621, :770-772)
    g = torch.linspace(min_guidance_scale, max_guidance_scale, f).unsqueeze(0).to(latents.device, latents.dtype)
    g = g.repeat(b, 1)[:, :, None, None, None]
    u, cnd = noise_pred.chunk(2)
    noise_pred = u + g * (cnd - u)
    return scheduler.step(noise_pred, i, latents)                                                     # :775
