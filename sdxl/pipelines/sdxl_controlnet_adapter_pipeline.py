"""The denoising loop body of /root/reference/sdxl/pipelines/sdxl_controlnet_adapter_pipeline.py:1279-1404.
The surrounding diffusers pipeline (prompt encoding, VAE) is out of scope."""
from ctrl_adapter_b200.pipeline_sdxl import SDXLControlNetAdapterLoop  # noqa: F401
