from oracle.unet_sdxl import UNet2DConditionModel  # noqa: F401
