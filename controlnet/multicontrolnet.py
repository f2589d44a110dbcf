"""Drop-in for /root/reference/controlnet/multicontrolnet.py."""
from ctrl_adapter_b200.controlnet import MultiControlNetModel  # noqa: F401
