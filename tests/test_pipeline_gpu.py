"""End-to-end separation pipeline (BASELINE.json config 3: cACGMM + permutation
alignment + PSD + GEV beamforming) on the device against the oracle, single
rank and -- when the box has two GPUs -- bin-sharded over two NCCL ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import cos_similarity, ROOT
from oracle import pb_bss_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def _oracle_pipeline(y, init, iterations, stft_size):
    model = O.cacgmm_fit(y, init, iterations)
    aff = O.cacgmm_predict(y, model)                       # (F, K, T)
    plan = O.dhtv_plan_from_stft_size(stft_size)
    mask = np.ascontiguousarray(aff.transpose(1, 0, 2))
    mapping = O.dhtv_calculate_mapping(mask, plan)
    aligned = O.apply_mapping(mask, mapping).transpose(1, 0, 2)
    Y = np.ascontiguousarray(np.swapaxes(y, -1, -2))
    psd = O.power_spectral_density(Y, aligned)
    noise = psd.sum(1, keepdims=True) - psd
    vec = O.gev_vector(psd, noise)
    enh = O.apply_beamforming_vector(vec.transpose(1, 0, 2), Y[None])
    return aligned, mapping, vec, enh.transpose(1, 0, 2)


def test_single_rank_pipeline_matches_oracle():
    import torch
    from pb_bss_b200.parallel import sharded_separation
    F, T, D, K, I = 257, 120, 6, 3, 12
    y, _ = synth.structured_stft(F, T, D, K, seed=11)
    init = synth.init_affiliation(F, K, T, seed=3)
    aligned, mapping, vec, enh = _oracle_pipeline(y, init, I, 512)
    out = sharded_separation(torch.from_numpy(y).cuda(), torch.from_numpy(init).cuda(), F,
                             iterations=I, stft_size=512)
    np.testing.assert_array_equal(out['mapping'].cpu().numpy(), mapping)
    np.testing.assert_allclose(out['affiliation'].cpu().numpy(), aligned, rtol=1e-5, atol=1e-8)
    v = out['vectors'].cpu().numpy()
    np.testing.assert_allclose(cos_similarity(v, vec), 1, atol=1e-7)
    np.testing.assert_allclose(np.abs(out['enhanced'].cpu().numpy()), np.abs(enh), rtol=1e-5, atol=1e-8)


def test_two_rank_bin_sharded_pipeline_matches_single_rank(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = tmp_path / 'c3.npz'
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', '29731',
           os.path.join(ROOT, 'scripts', 'run_c3.py'), '--check', str(out)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    assert out.exists()
