"""Minimal stand-in for `gymnasium` so the unmodified reference env modules import
in a container without it. TEST INFRASTRUCTURE ONLY (used by oracle/refload.py)."""
from . import spaces  # noqa: F401


class Env:
    pass
