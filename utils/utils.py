"""Drop-in for the names /root/reference/inference.py:15 imports from utils/utils.py.

Only `bool_flag` (argument parsing) is inside this repository's scope; the image / GIF helpers are host-side file I/O of the
reference's command-line tool (out of scope for the denoising hot path) and fail loudly if called.
"""
import argparse


def bool_flag(s):
    """Parse a boolean command-line value ("on"/"true"/"1" or "off"/"false"/"0", any case)."""
    v = str(s).lower()
    if v in ("off", "false", "0"):
        return False
    if v in ("on", "true", "1"):
        return True
    raise argparse.ArgumentTypeError("invalid value for a boolean flag")


def _out_of_scope(name):
    def fn(*_a, **_k):
        raise NotImplementedError(f"utils.utils.{name}: image / GIF file I/O of the reference CLI is outside the scope of "
                                  "ctrl_adapter_b200 (the per-timestep denoising hot path); use the reference's own helper")
    fn.__name__ = name
    return fn


center_crop_and_resize = _out_of_scope("center_crop_and_resize")
save_as_gif = _out_of_scope("save_as_gif")
save_concatenated_gif = _out_of_scope("save_concatenated_gif")
