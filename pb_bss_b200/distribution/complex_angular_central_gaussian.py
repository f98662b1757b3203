"""Complex angular central Gaussian -- container + observation normalisation.

Mirrors pb_bss/distribution/complex_angular_central_gaussian.py: the model is
stored as eigenvectors / eigenvalues of the covariance (:78-79), ``covariance``
(:140-148) and ``log_determinant`` (:150-152) are derived properties.
"""
from dataclasses import dataclass

import numpy as np
import torch

from .. import _device, _lib
from .utils import _ProbabilisticModel

__all__ = ['ComplexAngularCentralGaussian', 'normalize_observation']


def normalize_observation(observation):
    """(..., N, D) -> unit-norm (..., D, N) on the device.

    complex_angular_central_gaussian.py:34-55 (zero vectors stay zero).
    numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
    """
    like_numpy = not _device.is_tensor(observation)
    y = _device.to_device(observation)
    code = _device.complex_dtype_code(y)
    *independent, N, D = y.shape
    F = int(np.prod(independent)) if independent else 1
    z = torch.empty((*independent, D, N), dtype=y.dtype, device=y.device)
    lib = _lib.load()
    _lib.check(lib.pbb_normalize_observation(
        _device.ptr(y), _device.ptr(z), F, N, D, code, 1,
        _device.stream_ptr()), 'pbb_normalize_observation')
    return _device.to_host(z, like_numpy)


@dataclass
class ComplexAngularCentralGaussian(_ProbabilisticModel):
    covariance_eigenvectors: np.array = None  # (..., D, D)
    covariance_eigenvalues: np.array = None  # (..., D)

    @property
    def covariance(self):
        """V diag(lambda) V^H -- a derived view for inspection, not on the hot
        path (complex_angular_central_gaussian.py:140-148)."""
        V, lam = self.covariance_eigenvectors, self.covariance_eigenvalues
        if _device.is_tensor(V):
            return torch.einsum('...wx,...x,...zx->...wz', V, lam.to(V.dtype), V.conj())
        return np.einsum('...wx,...x,...zx->...wz', V, lam, V.conj())

    @property
    def log_determinant(self):
        lam = self.covariance_eigenvalues
        if _device.is_tensor(lam):
            return torch.sum(torch.log(lam), dim=-1)
        return np.sum(np.log(lam), axis=-1)
