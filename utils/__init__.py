"""Import-path shim: exposes the reference module paths (inference.py:9-15,351-367) backed by ctrl_adapter_b200."""
