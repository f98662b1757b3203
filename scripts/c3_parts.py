"""Per-stage device timing of the config-3 pipeline on one GPU (structured data, as scripts/run_c3.py)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
from pb_bss_b200.extraction import (apply_beamforming_vector, get_gev_vector, get_power_spectral_density_matrix)
from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment, apply_mapping

def timed(name, fn, reps=5):
    out = fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print('%-28s %8.3f ms (median of %d, max %.3f)' % (name, ts[len(ts) // 2] * 1e3, reps, ts[-1] * 1e3), flush=True)
    return out

F, T, D, K = 513, 500, 8, 3
y, _ = synth.structured_stft(F, T, D, K, seed=5)
init = synth.init_affiliation(F, K, T, seed=7)
yd, idv = torch.from_numpy(y).cuda(), torch.from_numpy(init).cuda()
model = timed('fit 100 it', lambda: CACGMMTrainer().fit(yd, initialization=idv, iterations=100))
aff = timed('predict', lambda: model.predict(yd))
mask = aff.permute(1, 0, 2).contiguous()
al = DHTVPermutationAlignment.from_stft_size(1024)
mapping = timed('DHTV calculate_mapping', lambda: al.calculate_mapping(mask))
aligned = timed('apply_mapping', lambda: apply_mapping(mask, mapping)).permute(1, 0, 2).contiguous()
Y = yd.transpose(-1, -2).contiguous()
psd = timed('PSD', lambda: get_power_spectral_density_matrix(Y, aligned))
noise = (psd.sum(1, keepdim=True) - psd).contiguous()
vec = timed('GEV', lambda: get_gev_vector(psd, noise))
timed('apply', lambda: apply_beamforming_vector(vec.permute(1, 0, 2).contiguous(), Y.unsqueeze(0).expand(K, *Y.shape)))
