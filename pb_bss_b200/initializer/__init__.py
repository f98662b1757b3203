"""Initialisers of the EM affiliations (pb_bss/initializer/__init__.py:1-3): ``iid`` (random, drawn on the host from
NumPy's global stream exactly like the reference, so seeded runs reproduce it), ``deterministic`` (``flag``) and
``deflation`` (``deflationSeed``: PSD + PCA + beamforming compositions of the device kernels)."""
from . import iid  # noqa: F401
from . import deflation  # noqa: F401
from . import deterministic  # noqa: F401
