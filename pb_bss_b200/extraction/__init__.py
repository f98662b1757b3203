"""Beamforming side of the hot path (pb_bss/extraction)."""
from . import linalg  # noqa: F401
