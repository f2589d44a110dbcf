"""Drop-in for /root/reference/model/ctrl_adapter.py: `from model.ctrl_adapter import ControlNetAdapter`."""
from ctrl_adapter_b200.adapter import AdapterSpatioTemporal, ControlNetAdapter  # noqa: F401
