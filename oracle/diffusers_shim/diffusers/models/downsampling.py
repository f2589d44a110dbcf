from oracle.blocks import Downsample2D  # noqa: F401


def _unsupported(*a, **k):
    raise NotImplementedError("FIR / 1-D resamplers are not on the Ctrl-Adapter hot path")


Downsample1D = FirDownsample2D = KDownsample2D = downsample_2d = _unsupported
