from oracle.blocks import (CrossAttnDownBlock3D, CrossAttnUpBlock3D, DownBlock3D, UNetMidBlock3DCrossAttn, UpBlock3D)  # noqa: F401
from oracle.blocks import (CrossAttnDownBlockSpatioTemporal, CrossAttnUpBlockSpatioTemporal, DownBlockSpatioTemporal,  # noqa: F401
                           UNetMidBlockSpatioTemporal, UpBlockSpatioTemporal)
from oracle.blocks import get_down_block_3d as get_down_block  # noqa: F401
from oracle.blocks import get_up_block_3d as get_up_block  # noqa: F401
