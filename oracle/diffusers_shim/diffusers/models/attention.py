from oracle.blocks import Attention, BasicTransformerBlock, FeedForward  # noqa: F401
