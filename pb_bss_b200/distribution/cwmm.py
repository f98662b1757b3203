"""Complex Watson mixture model: ``CWMMTrainer.fit / fit_predict`` and
``CWMM.predict`` with the signatures of pb_bss/distribution/cwmm.py, executed by
the kernels behind ``pbb_cwmm_fit`` / ``pbb_cwmm_predict``."""
from dataclasses import dataclass
from functools import cached_property
from operator import xor

import numpy as np
import torch

from .. import _device, _lib
from .cacgmm import _flatten_obs, _status_check, _weight_mode
from .complex_watson import ComplexWatson, ComplexWatsonTrainer
from .utils import _ProbabilisticModel

__all__ = ['CWMM', 'CWMMTrainer']


@dataclass
class CWMM(_ProbabilisticModel):
    weight: np.array = None  # (..., K, 1)
    complex_watson: ComplexWatson = None

    def predict(self, y):
        """Posterior affiliations (..., K, T) for y (..., T, D) (cwmm.py:26-52)."""
        like_numpy = not _device.is_tensor(y)
        yd = _device.to_device(y)
        code = _device.complex_dtype_code(yd)
        independent, F, N, D = _flatten_obs(yd)
        mode = _device.to_device(self.complex_watson.mode, torch.complex128)
        K = mode.shape[-2]
        assert mode.shape[-1] == D, (mode.shape, D)
        mode = mode.expand(*independent, K, D).reshape(F, K, D).contiguous()
        kappa = _device.to_device(self.complex_watson.concentration, torch.float64)
        kappa = kappa.expand(*independent, K).reshape(F, K).contiguous()
        w = _device.to_device(self.weight, torch.float64)
        assert w.shape[-1] == 1, w.shape
        w = w[..., 0].expand(*independent, K).reshape(F, K).contiguous()
        aff = _device.empty((F, K, N), torch.float64)
        status = _device.empty((1,), torch.int32)
        lib = _lib.load()
        nbytes = lib.pbb_cwmm_workspace_bytes(F, N, D, K)
        ws = _device.workspace(nbytes)
        _lib.check(lib.pbb_cwmm_predict(
            _device.ptr(yd), code, F, N, D, K, _device.ptr(mode),
            _device.ptr(kappa), _device.ptr(w), _device.ptr(aff),
            _device.ptr(ws), nbytes, _device.ptr(status),
            _device.stream_ptr()), 'pbb_cwmm_predict')
        _status_check(status, 'CWMM.predict')
        return _device.to_host(aff.reshape(*independent, K, N), like_numpy)


class CWMMTrainer:
    def __init__(self, dimension=None, max_concentration=500,
                 spline_markers=1000):
        self.dimension = dimension
        self.max_concentration = max_concentration
        self.spline_markers = spline_markers

    @cached_property
    def complex_watson_trainer(self):
        return ComplexWatsonTrainer(
            self.dimension, max_concentration=self.max_concentration,
            spline_markers=self.spline_markers)

    def fit(self, y, initialization=None, num_classes=None, iterations=100, *,
            saliency=None, weight_constant_axis=(-1,), affiliation_eps=0,
            inline_permutation_aligner=None):
        """EM for the complex Watson mixture model (cwmm.py:76-149).

        y: (..., T, D); initialization: affiliations (..., K, T) or None with
        ``num_classes`` (then drawn from NumPy's global RNG, cwmm.py:121-127).
        """
        assert xor(initialization is None, num_classes is None), (
            'Incompatible input combination. '
            'Exactly one of the two inputs has to be None: '
            f'{initialization is None} xor {num_classes is None}')
        assert affiliation_eps == 0, affiliation_eps  # cwmm.py:161
        if inline_permutation_aligner is not None:
            raise NotImplementedError(
                'inline_permutation_aligner is not on the device yet '
                '(SURVEY.md section 8f, rank 2)')
        like_numpy = not _device.is_tensor(y)
        yd = _device.to_device(y)
        assert yd.is_complex(), yd.dtype
        assert yd.shape[-1] > 1
        assert iterations > 0, iterations
        code = _device.complex_dtype_code(yd)
        independent, F, N, D = _flatten_obs(yd)
        if initialization is None:
            shape = (*independent, num_classes, N)
            initialization = np.random.uniform(size=shape)
            initialization /= np.einsum('...kn->...n', initialization)[..., None, :]
        K = initialization.shape[-2]
        init = _device.to_device(initialization, torch.float64)
        init = init.expand(*independent, K, N).reshape(F, K, N).contiguous()
        sal = None
        if saliency is not None:
            sal = _device.to_device(saliency, torch.float64)
            sal = sal.expand(*independent, N).reshape(F, N).contiguous()
        if self.dimension is None:
            self.dimension = D
        else:
            assert self.dimension == D, (
                'You initialized the trainer with a different dimension than '
                'you are using to fit a model. Use a new trainer, when you '
                'change the dimension.')
        weight_mode = _weight_mode(weight_constant_axis, len(independent) + 2)
        t_dev, c_dev = self.complex_watson_trainer.device_spline_table()
        mode = _device.empty((F, K, D), torch.complex128)
        kappa = _device.empty((F, K), torch.float64)
        w = _device.empty((F, K), torch.float64)
        status = _device.empty((1,), torch.int32)
        lib = _lib.load()
        nbytes = lib.pbb_cwmm_workspace_bytes(F, N, D, K)
        ws = _device.workspace(nbytes)
        _lib.check(lib.pbb_cwmm_fit(
            _device.ptr(yd), code, F, N, D, K, _device.ptr(init),
            _device.ptr(sal), int(iterations), weight_mode, _device.ptr(t_dev),
            _device.ptr(c_dev), int(c_dev.numel()),
            float(self.max_concentration), _device.ptr(mode),
            _device.ptr(kappa), _device.ptr(w), _device.ptr(ws), nbytes,
            _device.ptr(status), _device.stream_ptr()), 'pbb_cwmm_fit')
        _status_check(status, 'CWMMTrainer.fit')
        if weight_mode == _lib.WEIGHT_CONST:
            weight = np.full([K, 1], 1 / K)
            if not like_numpy:
                weight = _device.to_device(weight)
        else:
            weight = _device.to_host(w.reshape(*independent, K, 1), like_numpy)
        return CWMM(
            weight=weight,
            complex_watson=ComplexWatson(
                mode=_device.to_host(mode.reshape(*independent, K, D), like_numpy),
                concentration=_device.to_host(kappa.reshape(*independent, K), like_numpy)))

    def fit_predict(self, y, initialization=None, num_classes=None,
                    iterations=100, **kwargs):
        model = self.fit(y=y, initialization=initialization,
                         num_classes=num_classes, iterations=iterations,
                         **kwargs)
        return model.predict(y)
