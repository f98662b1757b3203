"""Deterministic initialisation (pb_bss/initializer/deterministic.py:4-86)."""
import numpy as np

from . import iid


def flag(Y, num_classes: int, permutation_free: bool = False, minimum: float = 0):
    """The time axis is cut into ``num_classes`` consecutive segments, segment k belongs to class k; with
    ``minimum`` > 0 every other class keeps that much probability (deterministic.py:64-86).
    Y (..., N, D) -> (..., K, N)."""
    if not permutation_free:
        raise NotImplementedError(permutation_free)
    *independent, N, _ = Y.shape
    owner = np.linspace(0, num_classes, N, dtype=int, endpoint=False)
    init = np.broadcast_to(np.eye(num_classes)[owner].T, [*independent, num_classes, N])
    if minimum != 0:
        assert 0 < minimum < (1 / num_classes), (minimum, num_classes)
        init = np.maximum(init, minimum / (1 - (num_classes - 1) * minimum))
        init = init / np.sum(init, keepdims=True, axis=-2)
    return iid._like(Y, init)
