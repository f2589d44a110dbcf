"""The denoising loop body of /root/reference/svd/pipelines/svd_controlnet_adapter_pipeline.py:640-787."""
from ctrl_adapter_b200.pipeline_svd import SVDControlNetAdapterLoop  # noqa: F401
