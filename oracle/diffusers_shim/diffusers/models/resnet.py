from oracle.blocks import AlphaBlender, ResnetBlock2D, TemporalResnetBlock  # noqa: F401
