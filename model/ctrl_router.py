"""Drop-in for /root/reference/model/ctrl_router.py: `from model.ctrl_router import ControlNetRouter`."""
from ctrl_adapter_b200.adapter import ControlNetRouter  # noqa: F401
