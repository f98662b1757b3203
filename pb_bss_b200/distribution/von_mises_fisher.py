"""von Mises-Fisher distribution over real embedding vectors (pb_bss/distribution/von_mises_fisher.py:31-144), tied
over all (bin, frame) observations, as the integrated model vmfcacgmm.py uses it.  The K x E parameters and the
normaliser (scipy's exponentially scaled Bessel function on K scalars) stay on the host; the F*T embeddings are only
touched by the device kernels."""
from dataclasses import dataclass

import numpy as np
import torch
from scipy.special import ive

from .. import _device, _lib
from .utils import _ProbabilisticModel


@dataclass
class VonMisesFisher(_ProbabilisticModel):
    mean: np.array = None           # (K, E)
    concentration: np.array = None  # (K,)

    def log_norm(self):
        """von_mises_fisher.py:36-46."""
        D = self.mean.shape[-1]
        kappa = np.asarray(self.concentration)
        return ((D / 2) * np.log(2 * np.pi) + np.log(ive(D / 2 - 1, kappa))
                + (np.abs(kappa) - (D / 2 - 1) * np.log(kappa)))

    def log_pdf_fkt(self, embedding):
        """embedding (F, T, E) CUDA tensor (any norm: normalised inside, von_mises_fisher.py:75-77) -> (F, K, T)."""
        F, T, E = embedding.shape
        K = self.mean.shape[0]
        mean = _device.to_device(np.ascontiguousarray(self.mean), torch.float64)
        kap = _device.to_device(np.ascontiguousarray(np.repeat(np.asarray(self.concentration)[:, None], E, axis=1)),
                                torch.float64)
        ln = _device.to_device(np.ascontiguousarray(self.log_norm()), torch.float64)
        out = _device.empty((F, K, T), torch.float64)
        lib = _lib.load()
        _lib.check(lib.pbb_gaussian_log_pdf(_device.ptr(embedding), _device.ptr(mean), _device.ptr(kap),
                                            _device.ptr(ln), F, T, E, K, 2, _device.ptr(out), _device.stream_ptr()),
                   'pbb_gaussian_log_pdf')
        return out


def vmf_fit_fkt(embedding, weight_fkt, min_concentration, max_concentration):
    """VonMisesFisherTrainer._fit (von_mises_fisher.py:122-144) over the F*T embeddings with weights (F, K, T): the
    weighted resultant comes from the first pass of pbb_gaussian_fit (sum w x, sum w), the K x E rest is host math."""
    F, T, E = embedding.shape
    K = weight_fkt.shape[1]
    mean = _device.empty((K, E), torch.float64)
    cov = _device.empty((K,), torch.float64)
    lib = _lib.load()
    scratch = _device.empty((int(lib.pbb_gaussian_fit_scratch_doubles(F, E, K)),), torch.float64)
    _lib.check(lib.pbb_gaussian_fit(_device.ptr(embedding), _device.ptr(weight_fkt), F, T, E, K, 1, _device.ptr(mean),
                                    _device.ptr(cov), _device.ptr(scratch), _device.stream_ptr()), 'pbb_gaussian_fit')
    total = scratch[F * K * (E + 1):F * K * (E + 1) + K].cpu().numpy()        # sum of the weights per class
    r = mean.cpu().numpy() * total[:, None]                                   # Banerjee2005vMF eq. 2.4
    norm = np.linalg.norm(r, axis=-1)
    direction = r / np.maximum(norm, np.finfo(np.float64).tiny)[..., None]
    r_bar = norm / total                                                      # eq. 2.5
    concentration = (r_bar * E - r_bar ** 3) / (1 - r_bar ** 2)               # eq. 4.4
    concentration = np.clip(concentration, min_concentration, max_concentration)
    return VonMisesFisher(mean=direction, concentration=concentration)
