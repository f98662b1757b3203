"""Batched small-matrix linear algebra on the device (thin ctypes wrappers)."""
import numpy as np
import torch

from .. import _device, _lib


def eigh(a):
    """Batched Hermitian eigendecomposition (..., D, D) -> (w (..., D) ascending,
    v (..., D, D) with eigenvectors as columns), like ``np.linalg.eigh``
    (complex_angular_central_gaussian.py:95, pb_bss/utils.py:154)."""
    like_numpy = not _device.is_tensor(a)
    ad = _device.to_device(a, torch.complex128)
    *lead, D, D2 = ad.shape
    assert D == D2, ad.shape
    n = int(np.prod(lead)) if lead else 1
    w = _device.empty((n, D), torch.float64)
    v = _device.empty((n, D, D), torch.complex128)
    status = torch.zeros(1, dtype=torch.int32, device=ad.device)
    lib = _lib.load()
    _lib.check(lib.pbb_heig_batched(
        _device.ptr(ad), n, D, _device.ptr(w), _device.ptr(v),
        _device.ptr(status), _device.stream_ptr()), 'pbb_heig_batched')
    s = int(status.item())
    if s:
        raise np.linalg.LinAlgError(f'eigh: non-finite input or no convergence in matrix {s - 1}')
    return (_device.to_host(w.reshape(*lead, D), like_numpy),
            _device.to_host(v.reshape(*lead, D, D), like_numpy))
