"""Drop-in for /root/reference/controlnet/controlnet.py: `from controlnet.controlnet import ControlNetModel`."""
from ctrl_adapter_b200.controlnet import ControlNetConditioningEmbedding, ControlNetModel  # noqa: F401


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module
