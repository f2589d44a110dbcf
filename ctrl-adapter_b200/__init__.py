"""ctrl-adapter_b200: B200-native (sm_100a) denoising hot path of Ctrl-Adapter.

The directory name follows the repository naming (``ctrl-adapter_b200``); import it as ``ctrl_adapter_b200``
(the sibling shim package forwards here).
"""
__version__ = "0.1.0"
