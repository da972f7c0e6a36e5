"""TEST INFRASTRUCTURE — minimal shim (gymnasium is not installed here)."""
from . import spaces


class Space:
    pass
