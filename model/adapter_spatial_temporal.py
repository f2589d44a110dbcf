"""Drop-in for /root/reference/model/adapter_spatial_temporal.py."""
from ctrl_adapter_b200.adapter import AdapterSpatioTemporal  # noqa: F401
