"""The denoising loop body of /root/reference/i2vgen_xl/pipelines/i2vgen_xl_controlnet_adapter_pipeline.py:902-1118."""
from ctrl_adapter_b200.pipeline_i2vgen import I2VGenXLControlNetAdapterLoop  # noqa: F401
