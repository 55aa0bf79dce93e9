"""lightning_asr_b200: B200-native (sm_100a) implementation of the lightning-asr training hot path."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
