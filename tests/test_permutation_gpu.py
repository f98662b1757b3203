"""DHTV permutation alignment on the device: integer mappings must be EXACTLY
the reference's (golden fixtures produced by the unmodified reference)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pb_bss_oracle as O

pytestmark = pytest.mark.gpu


def test_mapping_matches_reference_golden():
    from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment
    g = load_golden('permutation')
    for tag, stft in (('a', 512), ('b', 1024)):
        al = DHTVPermutationAlignment.from_stft_size(stft)
        assert al.alignment_plan == g[f'{tag}_plan'].tolist()
        mask = g[f'{tag}_mask']
        mapping = al.calculate_mapping(mask)
        assert mapping.dtype == np.int64 and mapping.shape == g[f'{tag}_mapping'].shape
        np.testing.assert_array_equal(mapping, g[f'{tag}_mapping'])
        np.testing.assert_array_equal(al.apply_mapping(mask, mapping), g[f'{tag}_aligned'])
        np.testing.assert_array_equal(al(mask), g[f'{tag}_aligned'])
    al = DHTVPermutationAlignment(stft_size=128, segment_start=20, segment_width=20, segment_shift=5,
                                  main_iterations=5, sub_iterations=2)
    assert al.alignment_plan == g['c_plan'].tolist()
    np.testing.assert_array_equal(al.calculate_mapping(g['c_mask']), g['c_mapping'])


@pytest.mark.parametrize('metric', ['cos', 'multiply', 'euclidean'])
@pytest.mark.parametrize('algorithm', ['greedy', 'optimal'])
def test_dhtv_options_match_reference_golden(metric, algorithm):
    """similarity_metric / algorithm of DHTVPermutationAlignment (permutation_alignment.py:133-163,380-420,556-585)."""
    from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment
    g = load_golden('permutation')
    al = DHTVPermutationAlignment(stft_size=512, segment_start=70, segment_width=100, segment_shift=20,
                                  main_iterations=20, sub_iterations=2, similarity_metric=metric, algorithm=algorithm)
    np.testing.assert_array_equal(al.calculate_mapping(g['a_mask']), g[f'opt_{metric}_{algorithm}'])


def test_dhtv_option_errors():
    from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment
    kw = dict(stft_size=512, segment_start=70, segment_width=100, segment_shift=20, main_iterations=2, sub_iterations=2)
    with pytest.raises(AttributeError):
        DHTVPermutationAlignment(similarity_metric='coss', **kw)
    with pytest.raises(ValueError):
        DHTVPermutationAlignment(algorithm='best', **kw).calculate_mapping(load_golden('permutation')['a_mask'])


def test_full_size_alignment_recovers_a_random_permutation():
    """K=3, F=513, T=500 (BASELINE.json config 3): permute a consistent mask per
    bin, align, and check against the oracle and the sortedness property
    (every bin ends up in the same global order)."""
    from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment, apply_mapping
    rng = np.random.RandomState(0)
    K, F, T = 3, 513, 500
    proto = rng.uniform(size=(K, 1, T)) ** 4
    mask = proto + 0.3 * rng.uniform(size=(K, F, T))
    mask /= mask.sum(0, keepdims=True)
    perm = np.stack([rng.permutation(K) for _ in range(F)], axis=1)
    permuted = mask[perm, np.arange(F)]
    al = DHTVPermutationAlignment.from_stft_size(1024)
    mapping = al.calculate_mapping(permuted)
    np.testing.assert_array_equal(mapping, O.dhtv_calculate_mapping(permuted, O.dhtv_plan_from_stft_size(1024)))
    aligned = apply_mapping(permuted, mapping)
    # up to ONE global permutation the original order is restored
    order = np.argmax(np.einsum('kft,lft->kl', aligned, mask), axis=1)
    assert sorted(order.tolist()) == [0, 1, 2]
    np.testing.assert_array_equal(aligned, mask[order])
    # permutations: every column of the mapping is a permutation of range(K)
    assert np.all(np.sort(mapping, axis=0) == np.arange(K)[:, None])


def test_apply_mapping_trailing_dims_and_device_tensors():
    import torch
    from pb_bss_b200.permutation_alignment import apply_mapping, sample_random_mapping
    rng = np.random.RandomState(1)
    K, F = 4, 9
    mask = rng.uniform(size=(K, F, 5, 3))
    mapping = sample_random_mapping(K, F, rng)
    np.testing.assert_array_equal(apply_mapping(mask, mapping), mask[mapping, range(F)])
    out = apply_mapping(torch.from_numpy(mask).cuda(), torch.from_numpy(mapping).cuda())
    assert out.is_cuda
    np.testing.assert_array_equal(out.cpu().numpy(), mask[mapping, range(F)])


METRICS = ('cos', 'euclidean', 'multiply')


@pytest.mark.parametrize('metric', METRICS)
def test_greedy_and_oracle_alignment_match_reference_golden(metric):
    """GreedyPermutationAlignment / OraclePermutationAlignment (permutation_alignment.py:592-786):
    exact integer equality with the mappings of the unmodified reference."""
    from pb_bss_b200.permutation_alignment import GreedyPermutationAlignment, OraclePermutationAlignment
    g = load_golden('permutation_greedy_oracle')
    for tag, mask, ref in (('', g['mask'], g['reference_mask']), ('noise_', g['noise'], g['noise_reference'])):
        got = GreedyPermutationAlignment(metric).calculate_mapping(mask)
        assert got.dtype == np.int64
        np.testing.assert_array_equal(got, g[f'greedy_{tag}{metric}'])
        for alg in ('greedy', 'optimal'):
            al = OraclePermutationAlignment(metric, alg)
            np.testing.assert_array_equal(al.calculate_mapping(mask, ref), g[f'oracle_{tag}{metric}_{alg}'])
    # the oracle alignment undoes the random permutation of the synthetic mask
    aligned = OraclePermutationAlignment(metric)(g['mask'], g['reference_mask'])
    np.testing.assert_array_equal(aligned, g['reference_mask'])


def test_mapping_from_score_matrix_known_answers_and_errors():
    from pb_bss_b200.permutation_alignment import (GreedyPermutationAlignment, OraclePermutationAlignment,
                                                   _mapping_from_score_matrix)
    sm = np.array([[11, 10, 0], [4, 5, 10], [6, 0, 5]])  # doctest, permutation_alignment.py:475-508
    np.testing.assert_array_equal(_mapping_from_score_matrix(sm, 'optimal'), [1, 2, 0])
    np.testing.assert_array_equal(_mapping_from_score_matrix(sm, 'greedy'), [0, 2, 1])
    np.testing.assert_array_equal(_mapping_from_score_matrix([sm, sm], 'optimal'), [[1, 1], [2, 2], [0, 0]])
    with pytest.raises(ValueError, match='infeasible'):
        _mapping_from_score_matrix([[np.inf, 0], [1, 2]])
    with pytest.raises(ValueError):
        _mapping_from_score_matrix(sm, 'hungarian')
    with pytest.raises(ValueError):
        GreedyPermutationAlignment('coss')
    with pytest.raises(AttributeError, match='Suggestions'):
        OraclePermutationAlignment('coss')
    with pytest.raises(AssertionError):
        GreedyPermutationAlignment('cos').calculate_mapping(np.ones((3, 4, 5)))  # even F


def test_full_size_greedy_alignment_against_the_oracle():
    """F=513, K=3, T=500 (C3 shapes): the device mapping equals the NumPy restatement."""
    from pb_bss_b200.permutation_alignment import GreedyPermutationAlignment
    rng = np.random.RandomState(5)
    K, F, T = 3, 513, 500
    proto = rng.uniform(size=(K, 1, T)) ** 4
    mask = proto + 0.5 * rng.uniform(size=(K, F, T))
    mask /= mask.sum(0, keepdims=True)
    perm = np.stack([rng.permutation(K) for _ in range(F)], axis=1)
    mask = mask[perm, np.arange(F)]
    got = GreedyPermutationAlignment('cos').calculate_mapping(mask)
    np.testing.assert_array_equal(got, O.greedy_permutation_alignment(mask, 'cos'))
    aligned = mask[got, np.arange(F)]
    # every bin now carries the same source order as bin 0
    assert (np.argmax(np.einsum('kft,jt->fkj', aligned, aligned[:, 0]), axis=-1) == np.arange(K)).all()
