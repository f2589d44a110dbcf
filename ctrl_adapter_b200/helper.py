"""``ControlNetHelper`` -- the part of /root/reference/model/ctrl_helper.py the inference pipelines call right before the
denoising loop:

* ``prepare_images`` (:268-296): PIL / ndarray / tensor condition images -> the CFG-duplicated fp tensor the ControlNet
  takes, for ANY batch size (the diffusers ``VaeImageProcessor(do_convert_rgb=True, do_normalize=False).preprocess`` it
  relies on is restated here: RGB conversion, Lanczos resize for PIL, nearest ``F.interpolate`` for tensors, /255);
* ``encode_controlnet_prompt`` (:301-457): the SD1.5 CLIP prompt encoding, on whatever ``tokenizer`` / ``text_encoder``
  objects the helper holds (transformers' ``CLIPTokenizer`` / ``CLIPTextModel`` loaded from a local snapshot, or any
  duck-typed pair).  The CLIP text encoder itself is not re-implemented here: it runs once per generation, outside the
  per-step hot path;
* ``_get_add_time_ids`` (:461-464).

The condition extractors (MiDaS depth, HED, OpenPose ... ``add_*_estimator`` / ``prepare_conditioning_images`` /
``prepare_batch``) are out of scope and fail loudly.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch


class ControlNetHelper:
    def __init__(self, pretrained_model_name_or_path: str = "runwayml/stable-diffusion-v1-5", use_size_512: bool = True,
                 text_encoder=None, tokenizer=None):
        self.pretrained_model_name_or_path = pretrained_model_name_or_path
        self.use_size_512 = use_size_512
        self.weight_dtype = torch.float16
        self.vae_scale_factor = 8
        self.text_encoder, self.tokenizer = text_encoder, tokenizer
        if text_encoder is None and os.path.isdir(str(pretrained_model_name_or_path)):  # local snapshot only: no hub
            from transformers import AutoTokenizer, CLIPTextModel
            self.text_encoder = CLIPTextModel.from_pretrained(pretrained_model_name_or_path, subfolder="text_encoder")
            self.tokenizer = AutoTokenizer.from_pretrained(pretrained_model_name_or_path, subfolder="tokenizer",
                                                           use_fast=False)
            self.text_encoder.requires_grad_(False)

    def to(self, *args, **kwargs):
        if isinstance(self.text_encoder, torch.nn.Module):
            self.text_encoder = self.text_encoder.to(*args, **kwargs)
        return self

    # ---- VaeImageProcessor(do_convert_rgb=True, do_normalize=False).preprocess, restated ---------------------------
    @staticmethod
    def _preprocess(image, height: int, width: int) -> torch.Tensor:
        """one image (PIL / HWC ndarray in [0,1] or uint8 / CHW or NCHW tensor in [0,1]) -> [n, 3, height, width] fp32"""
        try:
            from PIL import Image
        except Exception:  # pragma: no cover
            Image = None
        if Image is not None and isinstance(image, Image.Image):
            image = image.convert("RGB").resize((width, height), resample=Image.LANCZOS)
            arr = np.array(image).astype(np.float32) / 255.0
            return torch.from_numpy(arr[None].transpose(0, 3, 1, 2))
        if isinstance(image, np.ndarray):
            arr = image.astype(np.float32) / (255.0 if image.dtype == np.uint8 else 1.0)
            arr = arr[None] if arr.ndim == 3 else arr
            image = torch.from_numpy(arr.transpose(0, 3, 1, 2))
        if not torch.is_tensor(image):
            raise ValueError(f"unsupported control image type {type(image)}")
        image = image.float()
        image = image.unsqueeze(0) if image.dim() == 3 else image
        if tuple(image.shape[-2:]) != (height, width):
            image = torch.nn.functional.interpolate(image, size=(height, width))
        return image

    @torch.no_grad()
    def prepare_images(self, images, width, height, batch_size, num_images_per_prompt, device, dtype,
                       do_classifier_free_guidance=False, guess_mode=False):
        """ctrl_helper.py:268-296.  `images`: the frames of one clip (a list); returns [1 or 2, F * batch * n, 3, H, W]."""
        pre = torch.cat([self._preprocess(im, height, width) for im in images], dim=0)
        rep = [1] * pre.dim()
        rep[0] = batch_size * num_images_per_prompt
        out = pre.repeat(*rep).unsqueeze(0).to(device=device, dtype=dtype)
        if do_classifier_free_guidance and not guess_mode:
            rep = [1] * out.dim()
            rep[0] = 2
            out = out.repeat(*rep)
        return out

    @torch.no_grad()
    def encode_controlnet_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance,
                                 negative_prompt=None, prompt_embeds: Optional[torch.Tensor] = None,
                                 negative_prompt_embeds: Optional[torch.Tensor] = None, lora_scale=None, clip_skip=None):
        """ctrl_helper.py:301-457 (returns prompt, negative, pooled, negative pooled embeddings)."""
        if self.text_encoder is None or self.tokenizer is None:
            raise NotImplementedError("this ControlNetHelper holds no tokenizer / text encoder: construct it from a local "
                                      "SD1.5 snapshot folder or pass text_encoder= / tokenizer=, or give the pipeline "
                                      "controlnet_prompt_embeds directly")
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        pooled = negative_pooled = None
        if prompt_embeds is None:
            ids = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length,
                                 truncation=True, return_tensors="pt").input_ids
            if clip_skip is None:
                enc = self.text_encoder(ids.to(device), attention_mask=None)
                pooled, prompt_embeds = enc[1], enc[0]
            else:
                enc = self.text_encoder(ids.to(device), attention_mask=None, output_hidden_states=True)
                pooled = enc[1]
                prompt_embeds = self.text_encoder.text_model.final_layer_norm(enc[-1][-(clip_skip + 1)])
        dt = self.text_encoder.dtype if self.text_encoder is not None else prompt_embeds.dtype
        prompt_embeds = prompt_embeds.to(dtype=dt, device=device)
        bs, seq, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            if negative_prompt is None:
                uncond: List[str] = [""] * batch_size
            elif prompt is not None and type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} != "
                                f"{type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError("`negative_prompt` batch size does not match `prompt`")
            else:
                uncond = negative_prompt
            un = self.tokenizer(uncond, padding="max_length", max_length=prompt_embeds.shape[1], truncation=True,
                                return_tensors="pt")
            use_mask = getattr(getattr(self.text_encoder, "config", None), "use_attention_mask", False)
            enc = self.text_encoder(un.input_ids.to(device), attention_mask=un.attention_mask.to(device) if use_mask else None)
            negative_pooled, negative_prompt_embeds = enc[1], enc[0]
        if pooled is not None:
            pooled = pooled.repeat(1, num_images_per_prompt).view(bs * num_images_per_prompt, -1)
        if do_classifier_free_guidance:
            seq = negative_prompt_embeds.shape[1]
            negative_prompt_embeds = negative_prompt_embeds.to(dtype=dt, device=device)
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(
                batch_size * num_images_per_prompt, seq, -1)
            if negative_pooled is not None:
                negative_pooled = negative_pooled.repeat(1, num_images_per_prompt).view(bs * num_images_per_prompt, -1)
        return prompt_embeds, negative_prompt_embeds, pooled, negative_pooled

    def _get_add_time_ids(self, original_size, crops_coords_top_left, target_size, dtype):
        """ctrl_helper.py:461-464"""
        return torch.tensor([list(original_size + crops_coords_top_left + target_size)], dtype=dtype)

    # ---- condition extractors: outside the scope of this package --------------------------------
    def _out_of_scope(self, *_a, **_k):
        raise NotImplementedError("condition extractors (depth / normal / segmentation / softedge / lineart / shuffle / "
                                  "scribble / openpose estimators, prepare_conditioning_images, prepare_batch) are outside "
                                  "ctrl_adapter_b200: produce the control images with the reference's extractors")

    add_depth_estimator = add_normal_estimator = add_segmentation_estimator = add_softedge_estimator = _out_of_scope
    add_lineart_estimator = add_shuffle_estimator = add_scribble_estimator = add_openpose_estimator = _out_of_scope
    post_process_conditioning_pil_and_pixel_values = prepare_conditioning_images = prepare_batch = _out_of_scope
