"""Complex Watson distribution -- container and the concentration look-up table.

Mirrors pb_bss/distribution/complex_watson.py: ``ComplexWatson(mode,
concentration)`` (:31-71) and ``ComplexWatsonTrainer`` whose only state is the
quadratic spline that inverts the hypergeometric ratio (:237-271).  The spline
is built once on the host with the reference's recipe (SciPy ``hyp1f1`` and
``interp1d(kind='quadratic')``) and evaluated on the device by
``cw_update_kernel``; it is model state, not hot-path arithmetic.
"""
from dataclasses import dataclass
from functools import cached_property

import numpy as np
import torch

from .. import _device
from .utils import _ProbabilisticModel

__all__ = ['ComplexWatson', 'ComplexWatsonTrainer']


@dataclass
class ComplexWatson(_ProbabilisticModel):
    mode: np.array = None  # (..., D)
    concentration: np.array = None  # (...)


class ComplexWatsonTrainer:
    def __init__(self, dimension=None, max_concentration=500,
                 spline_markers=1000):
        self.dimension = dimension
        self.max_concentration = max_concentration
        self.spline_markers = spline_markers

    def hypergeometric_ratio(self, concentration):
        """Largest eigenvalue of the Watson covariance as a function of the
        concentration (complex_watson.py:258-262)."""
        from scipy.special import hyp1f1
        D = self.dimension
        return hyp1f1(2, D + 1, concentration) / (D * hyp1f1(1, D, concentration))

    @cached_property
    def spline(self):
        """scipy interpolant eigenvalue -> concentration (complex_watson.py:237-256)."""
        from scipy.interpolate import interp1d
        assert self.dimension is not None, (
            'You need to specify dimension. This can be done at object '
            'instantiation or it can be inferred when using the fit function.')
        x = np.logspace(-3, np.log10(self.max_concentration), self.spline_markers)
        y = self.hypergeometric_ratio(x)
        return interp1d(y, x, kind='quadratic', assume_sorted=True,
                        bounds_error=False,
                        fill_value=(0, self.max_concentration))

    def hypergeometric_ratio_inverse(self, eigenvalues):
        return self.spline(eigenvalues)

    @cached_property
    def spline_table(self):
        """(knots t[n+3], coefficients c[n]) of the quadratic B-spline as numpy."""
        bs = self.spline._spline
        assert bs.k == 2, bs.k
        return np.ascontiguousarray(bs.t, dtype=np.float64), \
            np.ascontiguousarray(bs.c.ravel(), dtype=np.float64)

    def device_spline_table(self):
        """The table as CUDA tensors (cached per device)."""
        cache = self.__dict__.setdefault('_dev_tables', {})
        dev = _device.device()
        if dev.index not in cache:
            t, c = self.spline_table
            cache[dev.index] = (_device.to_device(t, torch.float64),
                                _device.to_device(c, torch.float64))
        return cache[dev.index]
