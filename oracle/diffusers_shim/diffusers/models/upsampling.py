from oracle.blocks import Upsample2D  # noqa: F401


def _unsupported(*a, **k):
    raise NotImplementedError("FIR / 1-D resamplers are not on the Ctrl-Adapter hot path")


FirUpsample2D = KUpsample2D = Upsample1D = upfirdn2d_native = upsample_2d = _unsupported
