"""Minimal stand-in for the `diffusers` package (v0.27.2 API surface used by /root/reference) so that the reference's
OWN hot-path modules can be imported and executed on CPU in this container, where diffusers is not installed.
Every class re-exported here is the restatement in oracle/blocks.py.  TEST INFRASTRUCTURE ONLY: used by
tests/golden/make_golden.py to produce golden vectors from the real reference classes; never on the product path."""
__version__ = "0.27.2+oracle-shim"
