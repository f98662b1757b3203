"""Gaussians over real embedding vectors (pb_bss/distribution/gaussian.py:58-193), as far as the integrated model
needs them: diagonal and spherical covariances, tied over all (bin, frame) observations.

The K x E parameters live on the host / in tiny tensors; the N = F*T observations are only ever touched by the device
kernels (``pbb_gaussian_log_pdf``, ``pbb_gaussian_fit``).  ``covariance_type='full'`` (a Cholesky factor per class)
is not part of the accelerated path."""
from dataclasses import dataclass, field

import numpy as np
import torch

from .. import _device, _lib
from .utils import _ProbabilisticModel


def _np(x):
    return x.detach().cpu().numpy() if _device.is_tensor(x) else np.asarray(x)


@dataclass
class DiagonalGaussian(_ProbabilisticModel):
    mean: np.array = None        # (K, E)
    covariance: np.array = None  # (K, E)
    precision_cholesky: np.array = field(init=False, default=None)          # (K, E)
    log_det_precision_cholesky: np.array = field(init=False, default=None)  # (K,)

    def __post_init__(self):
        cov = _np(self.covariance)
        if np.any(cov <= 0.0):   # sklearn's _compute_precision_cholesky (gaussian.py:67)
            raise ValueError('Fitting the mixture model failed because some components have ill-defined empirical '
                             'covariance (for instance caused by singleton or collapsed samples).')
        self.precision_cholesky = 1.0 / np.sqrt(cov)
        self.log_det_precision_cholesky = np.sum(np.log(self.precision_cholesky), axis=-1)

    def _device_params(self, E):
        return (_device.to_device(_np(self.mean), torch.float64).contiguous(),
                _device.to_device(np.ascontiguousarray(self.precision_cholesky), torch.float64),
                _device.to_device(np.ascontiguousarray(self.log_det_precision_cholesky), torch.float64))

    def log_pdf_fkt(self, embedding):
        """embedding (F, T, E) CUDA tensor -> log pdf (F, K, T)."""
        return _gaussian_log_pdf(self, embedding)


@dataclass
class SphericalGaussian(_ProbabilisticModel):
    mean: np.array = None        # (K, E)
    covariance: np.array = None  # (K,)
    precision_cholesky: np.array = field(init=False, default=None)          # (K,)
    log_det_precision_cholesky: np.array = field(init=False, default=None)  # (K,)

    def __post_init__(self):
        cov = _np(self.covariance)
        if np.any(cov <= 0.0):
            raise ValueError('Fitting the mixture model failed because some components have ill-defined empirical '
                             'covariance (for instance caused by singleton or collapsed samples).')
        E = _np(self.mean).shape[-1]
        self.precision_cholesky = 1.0 / np.sqrt(cov)
        self.log_det_precision_cholesky = E * np.log(self.precision_cholesky)   # gaussian.py:106

    def _device_params(self, E):
        pc = np.repeat(np.asarray(self.precision_cholesky)[:, None], E, axis=1)
        return (_device.to_device(_np(self.mean), torch.float64).contiguous(),
                _device.to_device(np.ascontiguousarray(pc), torch.float64),
                _device.to_device(np.ascontiguousarray(self.log_det_precision_cholesky), torch.float64))

    def log_pdf_fkt(self, embedding):
        return _gaussian_log_pdf(self, embedding)


def _gaussian_log_pdf(model, embedding):
    F, T, E = embedding.shape
    mean, pc, ld = model._device_params(E)
    K = mean.shape[0]
    out = _device.empty((F, K, T), torch.float64)
    lib = _lib.load()
    # DiagonalGaussian: the reference's einsum quirk (gaussian.py:79-87) is reproduced by the kernel, see include/pbb.h
    _lib.check(lib.pbb_gaussian_log_pdf(_device.ptr(embedding), _device.ptr(mean), _device.ptr(pc), _device.ptr(ld),
                                        F, T, E, K, int(isinstance(model, DiagonalGaussian)), _device.ptr(out),
                                        _device.stream_ptr()), 'pbb_gaussian_log_pdf')
    return out


def gaussian_fit_fkt(embedding, weight_fkt, covariance_type):
    """GaussianTrainer._fit (gaussian.py:155-193) over the F*T embeddings with the weights (F, K, T)."""
    if covariance_type not in ('diagonal', 'spherical'):
        if covariance_type == 'full':
            raise NotImplementedError("covariance_type='full' is not on the device path (diagonal / spherical are)")
        raise ValueError(f"Unknown covariance type '{covariance_type}'.")
    F, T, E = embedding.shape
    K = weight_fkt.shape[1]
    spherical = covariance_type == 'spherical'
    mean = _device.empty((K, E), torch.float64)
    cov = _device.empty((K,) if spherical else (K, E), torch.float64)
    lib = _lib.load()
    scratch = _device.empty((int(lib.pbb_gaussian_fit_scratch_doubles(F, E, K)),), torch.float64)
    _lib.check(lib.pbb_gaussian_fit(_device.ptr(embedding), _device.ptr(weight_fkt), F, T, E, K, int(spherical),
                                    _device.ptr(mean), _device.ptr(cov), _device.ptr(scratch),
                                    _device.stream_ptr()), 'pbb_gaussian_fit')
    cls = SphericalGaussian if spherical else DiagonalGaussian
    return cls(mean=mean.cpu().numpy(), covariance=cov.cpu().numpy())
