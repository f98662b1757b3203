"""Integrated spatial + spectral model GCACGMM on the device against the unmodified reference
(tests/golden/gcacgmm.npz, oracle/make_golden.py: make_gcacgmm; pb_bss/distribution/gcacgmm.py:38-333)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = {
    'spherical': dict(),
    'diagonal_kt': dict(covariance_type='diagonal', weight_constant_axis=(-3,)),
    'spherical_k_inline': dict(weight_constant_axis=(-3, -1), inline_permutation_alignment=True),
    'spherical_sal_weights': dict(spatial_weight=0.7, spectral_weight=1.3),
}


@pytest.mark.parametrize('name', list(CASES))
def test_gcacgmm_fit_and_predict_match_reference_golden(name):
    from pb_bss_b200.distribution import GCACGMMTrainer
    g = load_golden('gcacgmm')
    kw = dict(CASES[name])
    if name == 'spherical_sal_weights':
        kw['saliency'] = g['saliency']
    model = GCACGMMTrainer().fit(g['y'], g['embedding'], initialization=g['init'], iterations=4, **kw)
    np.testing.assert_allclose(np.asarray(model.weight), g[f'{name}_weight'], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(model.gaussian.mean, g[f'{name}_mean'], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(model.gaussian.covariance, g[f'{name}_gcov'], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(model.cacg.covariance_eigenvalues, g[f'{name}_eigenvalues'], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(model.cacg.covariance, g[f'{name}_covariance'], rtol=1e-6, atol=1e-9)
    aff = model.predict(g['y'], g['embedding'])
    assert aff.shape == g[f'{name}_affiliation'].shape
    np.testing.assert_allclose(aff, g[f'{name}_affiliation'], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(aff.sum(-2), 1, atol=1e-12)


def test_gcacgmm_argument_errors():
    from pb_bss_b200.distribution import GCACGMMTrainer
    g = load_golden('gcacgmm')
    with pytest.raises(AssertionError):
        GCACGMMTrainer().fit(g['y'], g['embedding'])   # neither initialization nor num_classes
    with pytest.raises(NotImplementedError):
        GCACGMMTrainer().fit(g['y'], g['embedding'], initialization=g['init'], iterations=1, covariance_type='full')
    with pytest.raises(ValueError):
        GCACGMMTrainer().fit(g['y'], g['embedding'], initialization=g['init'], iterations=1, covariance_type='nope')
    np.random.seed(1)
    m = GCACGMMTrainer().fit(g['y'], g['embedding'], num_classes=2, iterations=2)
    assert m.predict(g['y'], g['embedding']).shape == (20, 2, 70)


VCASES = {
    'vmf': dict(),
    'vmf_kt_inline': dict(weight_constant_axis=(-3,), inline_permutation_alignment=True, max_concentration=50),
    'vmf_sal': dict(spatial_weight=0.6, spectral_weight=1.2, weight_constant_axis=(-3, -1)),
}


@pytest.mark.parametrize('name', list(VCASES))
def test_vmfcacgmm_fit_and_predict_match_reference_golden(name):
    """von Mises-Fisher + cACG (pb_bss/distribution/vmfcacgmm.py:34-301), fixtures from the unmodified reference."""
    from pb_bss_b200.distribution import VMFCACGMMTrainer
    g = load_golden('vmfcacgmm')
    kw = dict(VCASES[name])
    if name == 'vmf_sal':
        kw['saliency'] = g['saliency']
    model = VMFCACGMMTrainer().fit(g['y'], g['embedding'], initialization=g['init'], iterations=4, **kw)
    np.testing.assert_allclose(np.asarray(model.weight), g[f'{name}_weight'], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(model.vmf.mean, g[f'{name}_mean'], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(model.vmf.concentration, g[f'{name}_concentration'], rtol=1e-7)
    np.testing.assert_allclose(model.cacg.covariance, g[f'{name}_covariance'], rtol=1e-6, atol=1e-9)
    aff = model.predict(g['y'], g['embedding'])
    np.testing.assert_allclose(aff, g[f'{name}_affiliation'], rtol=1e-6, atol=1e-9)
