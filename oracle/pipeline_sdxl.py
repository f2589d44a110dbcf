"""ORACLE (test infrastructure): the SDXL denoising loop body of the reference pipeline restated on the oracle modules.

Follows /root/reference/sdxl/pipelines/sdxl_controlnet_adapter_pipeline.py:1279-1404 (loop body) with the diffusers
v0.27.2 EulerDiscreteScheduler (SDXL-base scheduler_config: scaled_linear betas 0.00085-0.012, 1000 train steps,
timestep_spacing "leading", steps_offset 1, epsilon prediction) restated below.  Not imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class EulerDiscreteScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset

    def set_timesteps(self, n, device="cpu"):
        step_ratio = self.num_train_timesteps // n
        timesteps = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.float32) + self.steps_offset
        sigmas = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas).to(device)
        self.timesteps = torch.from_numpy(timesteps).to(device)
        self.step_index = 0

    @property
    def init_noise_sigma(self):
        return (self.sigmas.max() ** 2 + 1) ** 0.5  # "leading" spacing

    def scale_model_input(self, sample, i):
        sigma = self.sigmas[i]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, i, sample):
        sample = sample.to(torch.float32)  # diffusers up-casts to avoid precision issues
        sigma = self.sigmas[i]
        pred_original_sample = sample - sigma * model_output  # epsilon prediction, gamma = 0
        derivative = (sample - pred_original_sample) / sigma
        dt = self.sigmas[i + 1] - sigma
        prev_sample = sample + derivative * dt
        return prev_sample.to(model_output.dtype)


@torch.no_grad()
def sdxl_step(controlnet, adapter, unet, scheduler, i, latents, prompt_embeds, add_text_embeds, add_time_ids,
              controlnet_prompt_embeds, images, guidance_scale=5.0, cond_scale=1.0, use_size_512=True):
    """One iteration of the loop at pipeline :1279-1378 (do_classifier_free_guidance=True, guess_mode=False)."""
    t = scheduler.timesteps[i]
    latent_model_input = torch.cat([latents] * 2)  # :1284
    latent_model_input = scheduler.scale_model_input(latent_model_input, i)  # :1285
    control_model_input = latent_model_input  # :1295
    _, _, h, w = control_model_input.shape
    if (h, w) != (64, 64) and use_size_512:  # :1306-1312
        reshaped_in = F.adaptive_avg_pool2d(control_model_input, (64, 64))
        reshaped_images = F.adaptive_avg_pool2d(images, (512, 512))
    else:
        reshaped_in, reshaped_images = control_model_input, images
    down, mid = controlnet(reshaped_in, t, encoder_hidden_states=controlnet_prompt_embeds,
                           controlnet_cond=reshaped_images, conditioning_scale=cond_scale, guess_mode=False,
                           return_dict=False)  # :1323-1334
    adapted_down, adapted_mid = adapter([d.to(latents.dtype) for d in down], sparsity_masking=None, num_frames=1,
                                        timestep=t, encoder_hidden_states=prompt_embeds)  # :1338-1343
    refilled = None if cond_scale == 0 else adapted_down  # :1348-1349
    noise_pred = unet(latent_model_input, t, encoder_hidden_states=prompt_embeds,
                      added_cond_kwargs={"text_embeds": add_text_embeds, "time_ids": add_time_ids},
                      down_block_additional_residuals=refilled, mid_block_additional_residual=0,
                      return_dict=False)[0]  # :1356-1366
    u, c = noise_pred.chunk(2)  # :1369-1371
    noise_pred = u + guidance_scale * (c - u)
    return scheduler.step(noise_pred, i, latents)  # :1378
