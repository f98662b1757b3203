"""Independent and identically distributed initialisers (pb_bss/initializer/iid.py:12-217).

The reference draws from NumPy's GLOBAL Mersenne-Twister stream; a seeded caller only gets the reference's numbers if
the same NumPy calls happen in the same order, so the draws stay on the host (they are O(F K T) and run once per
fit).  A CUDA tensor ``Y`` gets its initialisation back as a CUDA tensor (one upload), NumPy in gives NumPy out.
``Y``: (..., N, D) -> affiliations (..., K, N).  ``permutation_free``: one draw shared by all independent dims.
"""
import numpy as np

from .. import _device

__all__ = ['uniform_normalized', 'dirichlet_uniform', 'dirichlet', 'one_hot']


def _shapes(Y, num_classes):
    independent, N = tuple(Y.shape[:-2]), Y.shape[-2]
    return independent, N, (*independent, num_classes, N)


def _like(Y, affiliation):
    if _device.is_tensor(Y):
        import torch
        out = torch.from_numpy(np.ascontiguousarray(affiliation))
        return out.to(Y.device) if Y.is_cuda else out
    return affiliation


def uniform_normalized(Y, num_classes: int, permutation_free: bool = False):
    """uniform(0, 1) per class, normalised over the classes (iid.py:12-69)."""
    independent, N, shape = _shapes(Y, num_classes)
    draw = np.random.uniform(size=shape[-2:] if permutation_free else shape)
    draw /= draw.sum(axis=-2, keepdims=True)
    return _like(Y, np.broadcast_to(draw, shape) if permutation_free else draw)


def dirichlet(Y, num_classes: int, permutation_free: bool = False, alpha=1):
    """Dirichlet(alpha, ..., alpha) per observation (iid.py:87-148)."""
    independent, N, shape = _shapes(Y, num_classes)
    assert np.isscalar(alpha), alpha
    a = np.broadcast_to(alpha, (num_classes,))
    if permutation_free:
        return _like(Y, np.broadcast_to(np.random.dirichlet(a, size=N).T, shape))
    return _like(Y, np.swapaxes(np.random.dirichlet(a, size=(*independent, N)), -1, -2))


def dirichlet_uniform(Y, num_classes, permutation_free=False):
    """Dirichlet(1, ..., 1): uniform on the simplex (iid.py:72-84)."""
    return dirichlet(Y, num_classes, permutation_free, alpha=1)


def one_hot(Y, num_classes: int, permutation_free: bool = False):
    """One class per observation, drawn uniformly (iid.py:151-217)."""
    independent, N, shape = _shapes(Y, num_classes)
    eye = np.eye(num_classes)
    if permutation_free:
        return _like(Y, np.broadcast_to(eye[np.random.randint(num_classes, size=N)].T, shape))
    return _like(Y, np.swapaxes(eye[np.random.randint(num_classes, size=(*independent, N))], -1, -2))
