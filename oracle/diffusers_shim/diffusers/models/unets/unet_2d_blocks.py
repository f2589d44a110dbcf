from oracle.blocks import (CrossAttnDownBlock2D, CrossAttnUpBlock2D, DownBlock2D, UNetMidBlock2DCrossAttn, UpBlock2D,  # noqa: F401
                           get_down_block, get_up_block)


class UNetMidBlock2D:
    def __init__(self, *a, **k):
        raise NotImplementedError("UNetMidBlock2D (attention-free mid block) is not used by the SD1.5 ControlNet")
