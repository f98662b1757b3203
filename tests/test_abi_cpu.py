"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a
GPU and exports every symbol include/pbb.h declares; host-side argument logic."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pbb.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pbb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from pb_bss_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 6, names
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/pbb.h but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature in _lib.py'
    assert set(_lib.SIGNATURES) == set(names)


def test_version_error_string_and_workspace_size():
    from pb_bss_b200 import _lib
    lib = _lib.load()
    assert lib.pbb_version() >= 100
    assert isinstance(lib.pbb_last_error(), bytes)
    n = lib.pbb_cacgmm_workspace_bytes(513, 500, 8, 3)
    assert 513 * 500 * 8 * 16 < n < 1 << 30
    assert lib.pbb_cacgmm_workspace_bytes(0, 1, 1, 1) == 0


def test_bad_arguments_are_rejected_before_any_launch():
    """Negative return = index of the offending argument (LAPACK INFO<0 style);
    no GPU is touched for these."""
    from pb_bss_b200 import _lib
    lib = _lib.load()
    opts = _lib.CacgmmOptions(iterations=1, covariance_norm=1, weight_mode=0, hermitize=1,
                              affiliation_eps=1e-10, eigenvalue_floor=1e-10, frames_per_block=0, reserved=0)
    rc = lib.pbb_cacgmm_fit(None, 1, 1, 1, 4, 2, None, None, None, ctypes.byref(opts),
                            None, None, None, None, 0, None, None)
    assert rc == -1 and b'y is null' in lib.pbb_last_error()
    rc = lib.pbb_cacgmm_fit(1, 7, 1, 1, 4, 2, None, None, None, ctypes.byref(opts),
                            None, None, None, None, 0, None, None)
    assert rc == -2
    rc = lib.pbb_cacgmm_fit(1, 1, 1, 1, 40, 2, None, None, None, ctypes.byref(opts),
                            None, None, None, None, 0, None, None)
    assert rc == -5 and b'D < 35' in lib.pbb_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, 'pbb_cacgmm_fit')


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pb_bss_b200.distribution import CACGMMTrainer
    y = np.ones((2, 10, 4), dtype=np.complex128)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        CACGMMTrainer().fit(y, num_classes=2, iterations=1)


def test_weight_axis_mapping():
    from pb_bss_b200 import _lib
    from pb_bss_b200.distribution.cacgmm import _weight_mode
    assert _weight_mode((-1,), 3) == _lib.WEIGHT_TIME
    assert _weight_mode([-1], 3) == _lib.WEIGHT_TIME
    assert _weight_mode(2, 3) == _lib.WEIGHT_TIME
    assert _weight_mode(-2, 3) == _lib.WEIGHT_CONST
    assert _weight_mode(1, 3) == _lib.WEIGHT_CONST
    assert _weight_mode((-3,), 3) == _lib.WEIGHT_TIED_TIME
    assert _weight_mode((-3, -1), 3) == _lib.WEIGHT_TIED
    assert _weight_mode((-1, -3), 3) == _lib.WEIGHT_TIED
    # more than one independent dim: the bins (axis -3) are tied, the dims in front stay independent fits
    assert _weight_mode((-3,), 4) == _lib.WEIGHT_TIED_TIME
    assert _weight_mode((-3, -1), 5) == _lib.WEIGHT_TIED
    with pytest.raises(NotImplementedError):
        _weight_mode((-4, -1), 4)


@pytest.mark.parametrize('F,I,arrive,cap', [(513, 100, 15, 467), (129, 20, 30, 292), (7, 5, 1, 292), (40, 12, 3, 10),
                                            (513, 1, 8, 292), (1, 9, 4, 4), (2000, 3, 64, 50)])
def test_streamed_task_order_is_a_valid_schedule(F, I, arrive, cap):
    """Host logic of the streamed upload (api_cacgmm.cu, build_streamed_order): every (bin, iteration) exactly
    once, (bin, it) after (bin, it - 1), bins entering in ascending (= arrival) order -- the properties the
    persistent kernel's no-deadlock argument rests on."""
    from pb_bss_b200 import _lib
    lib = _lib.load()
    order = np.zeros(F * I, dtype=np.int32)
    rc = lib.pbb_streamed_task_order(F, I, arrive, cap, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0
    bins, its = order & 0xffff, order >> 16
    assert bins.min() >= 0 and bins.max() == F - 1 and its.min() == 0 and its.max() == I - 1
    assert len(set(zip(bins.tolist(), its.tolist()))) == F * I
    pos = np.empty((F, I), dtype=np.int64)
    pos[bins, its] = np.arange(F * I)
    assert (np.diff(pos, axis=1) > 0).all()
    if cap >= arrive:  # then bins enter in the order in which they arrive
        assert (np.diff(pos[:, 0]) > 0).all()
    assert lib.pbb_streamed_task_order(F, 0, arrive, cap, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == -2
