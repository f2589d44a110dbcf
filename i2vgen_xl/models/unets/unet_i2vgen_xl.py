"""Drop-in for /root/reference/i2vgen_xl/models/unets/unet_i2vgen_xl.py."""
from ctrl_adapter_b200.unet_i2vgen import I2VGenXLUNet  # noqa: F401
