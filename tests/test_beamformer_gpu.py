"""Parity of the device beamforming chain against the reference's golden
fixtures and its own unit-test known answers.  Eigenvector-type outputs have an
arbitrary phase and are compared by cos-similarity plus their normalisation
(tests/test_extraction/test_beamformer.py:18-22 of the reference)."""
import numpy as np
import pytest

from conftest import load_golden, cos_similarity
from oracle import pb_bss_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def test_psd_matches_reference_golden():
    from pb_bss_b200.extraction import get_power_spectral_density_matrix as psd
    g = load_golden('beamformer')
    Y, mask = g['Y'], g['mask']
    np.testing.assert_allclose(psd(Y, mask), g['psd'], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(psd(Y, mask, normalize=False), g['psd_nonorm'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(psd(Y, mask[:, 0]), g['psd_single'], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(psd(Y), g['psd_nomask'], rtol=1e-11, atol=1e-14)
    # (K, F, T) mask with source_dim=-3 -> (K, F, D, D), beamformer.py:156-158
    out = psd(Y, np.ascontiguousarray(mask.transpose(1, 0, 2)), source_dim=-3)
    np.testing.assert_allclose(out, g['psd'].transpose(1, 0, 2, 3), rtol=1e-11, atol=1e-14)
    # complex64 observation
    np.testing.assert_allclose(psd(Y.astype(np.complex64), mask), g['psd'], rtol=1e-5, atol=1e-6)


def test_psd_properties():
    """Hermitian, PSD, mask-scale invariant (tests/test_extraction/test_covariance_matrix.py:31-107)."""
    from pb_bss_b200.extraction import get_power_spectral_density_matrix as psd
    Y = np.swapaxes(synth.noise_stft(7, 130, 8, seed=5), -1, -2).copy()
    mask = np.random.RandomState(1).uniform(size=(7, 3, 130))
    P = psd(Y, mask)
    np.testing.assert_allclose(P, np.conj(np.swapaxes(P, -1, -2)), atol=1e-15)
    assert np.all(np.linalg.eigvalsh(P) > -1e-12)
    np.testing.assert_allclose(psd(Y, 7.5 * mask), P, rtol=1e-12)
    np.testing.assert_allclose(P, O.power_spectral_density(Y, mask), rtol=1e-11, atol=1e-14)


def test_vectors_match_reference_golden():
    from pb_bss_b200 import extraction as E
    g = load_golden('beamformer')
    target, noise = g['target'], g['noise']
    pca = E.get_pca_vector(target)
    np.testing.assert_allclose(cos_similarity(pca, g['pca']), 1, atol=1e-10)
    np.testing.assert_allclose(np.linalg.norm(pca, axis=-1), 1, atol=1e-12)
    np.testing.assert_allclose(E.get_mvdr_vector(g['pca'], noise), g['mvdr'], rtol=1e-9, atol=1e-12)
    gev = E.get_gev_vector(target, noise)
    np.testing.assert_allclose(cos_similarity(gev, g['gev']), 1, atol=1e-10)
    np.testing.assert_allclose(np.einsum('fa,fab,fb->f', gev.conj(), noise, gev).real, 1, rtol=1e-10)
    np.testing.assert_allclose(np.linalg.norm(gev, axis=-1), np.linalg.norm(g['gev'], axis=-1), rtol=1e-9)
    s, ch = E.get_mvdr_vector_souden(target, noise, return_ref_channel=True)
    assert ch == int(g['ref_channel'])
    np.testing.assert_allclose(s, g['souden'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(E.blind_analytic_normalization(g['gev'], noise), g['ban'], rtol=1e-10)
    np.testing.assert_allclose(E.apply_beamforming_vector(g['gev'], g['Y']), g['applied'], rtol=1e-11, atol=1e-13)


def test_gev_equals_pca_for_identity_noise_and_leading_dims():
    """tests/test_extraction/test_beamformer.py:98-104, shapes :25-118."""
    from pb_bss_b200 import extraction as E
    K, F, D = 2, 51, 6
    target = synth.pos_def_hermitian(K, F, D, D, seed=3)
    noise = np.broadcast_to(np.eye(D, dtype=np.complex128), (K, F, D, D)).copy()
    gev = E.get_gev_vector(target, noise)
    assert gev.shape == (K, F, D)
    np.testing.assert_allclose(cos_similarity(gev, E.get_pca_vector(target)), 1, atol=1e-10)
    ref = O.gev_vector(target, synth.pos_def_hermitian(K, F, D, D, seed=4))
    got = E.get_gev_vector(target, synth.pos_def_hermitian(K, F, D, D, seed=4))
    np.testing.assert_allclose(cos_similarity(got, ref), 1, atol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(got, axis=-1), np.linalg.norm(ref, axis=-1), rtol=1e-8)
    assert E.get_gev_vector(target[0, :1], noise[0, :1]).shape == (1, D)


def test_souden_known_answer():
    """tests/test_extraction/test_beamformer.py:185-209 of the reference."""
    from pb_bss_b200.extraction import get_mvdr_vector_souden
    obs = np.array([[0, 0, 1], [0, 0.1, 1], [0.1, 0, 1]])
    phi_xx = (obs.T.conj() @ obs).astype(np.complex128)
    w, = get_mvdr_vector_souden(phi_xx[None], np.eye(3, dtype=np.complex128)[None])
    np.testing.assert_allclose(w, [0.03311258, 0.03311258, 0.99337748], atol=1e-8)
    w3 = get_mvdr_vector_souden(np.stack([phi_xx] * 3), np.stack([np.eye(3, dtype=np.complex128)] * 3))
    np.testing.assert_allclose(w3, [w] * 3)


def test_deferred_status_raises_at_the_end_of_the_block():
    """pb_bss_b200.deferred_status(): status words are read when the block ends (no per-call synchronisation), the
    first failing call raises its own exception there; outside a block the call raises on the spot."""
    import torch
    import pb_bss_b200
    from pb_bss_b200 import extraction as E
    D = 4
    a = torch.from_numpy(synth.pos_def_hermitian(3, D, D, seed=1)).cuda()
    b = a.clone()
    b[1] = -torch.eye(D, dtype=torch.complex128)  # not positive definite
    with pytest.raises(ValueError, match='frequency 1'):
        E.get_gev_vector(a, b)
    reached = []
    with pytest.raises(ValueError, match='frequency 1'):
        with pb_bss_b200.deferred_status():
            w = E.get_gev_vector(a, b)       # does not raise here
            reached.append(tuple(w.shape))
            E.get_gev_vector(a, a)           # a later, healthy call
            reached.append('second')
    assert reached == [(3, D), 'second']
    with pb_bss_b200.deferred_status():       # nothing wrong: no exception, results as usual
        w = E.get_gev_vector(a, a)
    assert torch.isfinite(torch.view_as_real(w)).all()


def test_error_paths():
    from pb_bss_b200 import extraction as E
    D = 4
    a = synth.pos_def_hermitian(3, D, D, seed=1)
    b = a.copy()
    b[1] = -np.eye(D)  # not positive definite
    with pytest.raises(ValueError):
        E.get_gev_vector(a, b)
    # a singular noise PSD no longer raises: like the reference (beamformer.py:251-256) the bin takes the lstsq
    # fallback; an all-zero matrix gives 0 / 0 there and leaves the other bins untouched
    sing = a.copy()
    sing[2] = 0
    atf = np.ones((3, D), dtype=np.complex128)
    w = E.get_mvdr_vector(atf, sing)
    assert np.all(np.isnan(w[2])) and np.all(np.isfinite(w[:2]))
    np.testing.assert_allclose(w[:2], O.mvdr_vector(atf[:2], sing[:2]), rtol=1e-9)
    with pytest.raises(NotImplementedError):
        E.get_gev_vector(a, a, use_eig=True)


def test_mvdr_distortionless_and_broadcast():
    from pb_bss_b200 import extraction as E
    F, D = 33, 8
    noise = synth.pos_def_hermitian(F, D, D, seed=2)
    atf = np.random.RandomState(0).randn(2, F, D) + 1j * np.random.RandomState(1).randn(2, F, D)
    w = E.get_mvdr_vector(atf, noise)
    assert w.shape == (2, F, D)
    np.testing.assert_allclose(np.einsum('kfd,kfd->kf', w.conj(), atf), 1, atol=1e-10)
    np.testing.assert_allclose(w, O.mvdr_vector(atf, noise), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('name,phase_free', [
    ('pca', True), ('pca+mvdr', True), ('scaled_gev_atf+mvdr', True), ('mvdr_souden', False),
    ('mvdr_souden+ban', False), ('rank1_pca+mvdr_souden', False), ('rank1_gev+mvdr_souden+ban', False),
    ('gev', True), ('gev+ban', True), ('rank1_pca+gev', True), ('ch1', False)])
def test_get_bf_vector_matches_reference_golden(name, phase_free):
    """String dispatcher of pb_bss/extraction/beamformer_wrapper.py:117-236."""
    from pb_bss_b200.extraction import get_bf_vector
    g = load_golden('bf_wrapper')
    w = get_bf_vector(name, g['target'], g['noise'])
    ref = g['bf_' + name]
    assert w.shape == ref.shape
    if phase_free:  # eigenvector based: arbitrary phase per bin
        np.testing.assert_allclose(cos_similarity(w, ref), 1, atol=1e-9)
        np.testing.assert_allclose(np.linalg.norm(w, axis=-1), np.linalg.norm(ref, axis=-1), rtol=1e-8)
    else:
        np.testing.assert_allclose(w, ref, rtol=1e-8, atol=1e-11)


def test_rank_one_estimates_match_reference_golden():
    from pb_bss_b200.extraction.beamformer_wrapper import get_gev_rank_one_estimate, get_pca_rank_one_estimate
    g = load_golden('bf_wrapper')
    np.testing.assert_allclose(get_pca_rank_one_estimate(g['target']), g['rank1_pca'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(get_gev_rank_one_estimate(g['target'], g['noise']), g['rank1_gev'], rtol=1e-8, atol=1e-12)
    with pytest.raises(NotImplementedError):
        from pb_bss_b200.extraction import get_bf_vector
        get_bf_vector('wmwf', g['target'], g['noise'])
    with pytest.raises(ValueError):
        from pb_bss_b200.extraction import get_bf_vector
        get_bf_vector('nonsense', g['target'], g['noise'])


class _Souden:
    """Fixture of tests/test_extraction/test_beamformer.py:185-203 of the reference."""
    obs = np.array([[0, 0, 1], [0, 0.1, 1], [0.1, 0, 1]])
    PhiXX = obs.T.conj() @ obs          # real, like the reference's fixture
    PhiNN = np.eye(3)
    well = np.array([0.03311258, 0.03311258, 0.99337748])


def test_souden_difficulties_like_the_reference():
    """The reference pins the behaviour of get_mvdr_vector_souden on zero / inf PSD matrices
    (tests/test_extraction/test_beamformer.py:211-376): zeros give a zero vector (stable_solve's lstsq fallback,
    math/solve.py:95-114), infinities an AssertionError, and with eps = 0 every broken bin is an AssertionError."""
    from pb_bss_b200.extraction import get_mvdr_vector_souden as souden
    X, N = _Souden.PhiXX, _Souden.PhiNN
    with np.errstate(all='ignore'):
        for args in ((X[None] * 0, N[None]), (X[None], N[None] * 0), (X[None] * 0, N[None] * 0)):
            w = souden(*args)
            assert repr(w) == 'array([[0., 0., 0.]])', repr(w)
        for args in ((X[None] * np.inf, N[None]), (X[None], N[None] * np.inf), (X[None] * np.inf, N[None] * np.inf)):
            with pytest.raises(AssertionError):
                souden(*args)
        # eps = 0, single bin: everything broken is an AssertionError
        for args in ((X[None] * 0, N[None]), (X[None], N[None] * 0), (X[None] * 0, N[None] * 0),
                     (X[None] * np.inf, N[None]), (X[None], N[None] * np.inf), (X[None] * np.inf, N[None] * np.inf)):
            with pytest.raises(AssertionError):
                souden(*args, eps=0)
        # several bins: zero bins only damage themselves
        for args in (([X * 0, X], [N, N]), ([X, X], [N * 0, N]), ([X * 0, X], [N * 0, N])):
            w, ref_channel = souden(np.array(args[0]), np.array(args[1]), return_ref_channel=True)
            assert ref_channel == 2, ref_channel
            np.testing.assert_allclose(w, np.array([[0., 0., 0.], _Souden.well]), atol=1e-8)
        for args in (([X * np.inf, X], [N, N]), ([X, X], [N * np.inf, N]), ([X * np.inf, X], [N * np.inf, N])):
            with pytest.raises(AssertionError):
                souden(np.array(args[0]), np.array(args[1]), return_ref_channel=True)
        for args in (([X * 0, X], [N, N]), ([X, X], [N * 0, N]), ([X * 0, X], [N * 0, N]),
                     ([X * np.inf, X], [N, N]), ([X, X], [N * np.inf, N]), ([X * np.inf, X], [N * np.inf, N])):
            with pytest.raises(AssertionError):
                souden(np.array(args[0]), np.array(args[1]), eps=0, return_ref_channel=True)


def test_singular_systems_take_the_minimum_norm_solution():
    """pbb_solve_batched against np.linalg.lstsq (what stable_solve falls back to, math/solve.py:111-113) on exactly
    singular matrices: a zero matrix, a Hermitian PSD matrix of rank 2, a general matrix with a zero row."""
    from pb_bss_b200.extraction.linalg import solve
    rng = np.random.RandomState(3)
    D = 6
    A = synth.pos_def_hermitian(5, D, D, seed=2)
    B = rng.randn(5, D, D) + 1j * rng.randn(5, D, D)
    A[1] = 0
    v = rng.randn(D, 2) + 1j * rng.randn(D, 2)
    A[2] = v @ v.conj().T                      # rank 2: elimination meets an exactly zero pivot only by luck ...
    A[2][:, 3:] = 0; A[2][3:, :] = 0           # ... so make the deficiency exact
    G = rng.randn(D, D) + 1j * rng.randn(D, D)
    G[4] = 0                                   # non-Hermitian, zero row
    A[3] = G
    X = solve(A, B)
    for i in range(5):
        ref = np.linalg.lstsq(A[i], B[i], rcond=None)[0]
        np.testing.assert_allclose(X[i], ref, rtol=1e-7, atol=1e-9, err_msg=f'system {i}')
