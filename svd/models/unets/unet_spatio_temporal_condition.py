"""Drop-in for /root/reference/svd/models/unets/unet_spatio_temporal_condition.py."""
from ctrl_adapter_b200.unet_svd import UNetSpatioTemporalConditionModel  # noqa: F401
