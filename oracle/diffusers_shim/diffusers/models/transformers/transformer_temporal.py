from oracle.blocks import BasicTransformerBlock, TemporalBasicTransformerBlock, TransformerTemporalModel  # noqa: F401
