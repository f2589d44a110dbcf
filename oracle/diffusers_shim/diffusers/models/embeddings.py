from oracle.blocks import TimestepEmbedding, Timesteps  # noqa: F401


class _Unsupported:
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the Ctrl-Adapter hot path; not restated in the oracle shim")


TextImageProjection = TextImageTimeEmbedding = TextTimeEmbedding = _Unsupported
