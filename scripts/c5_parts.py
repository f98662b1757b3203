"""Per-stage device timing of the config-5 pipeline for the 8 utterances of one rank in ONE batched call."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
from pb_bss_b200 import extraction as E

def timed(name, fn, reps=3):
    out = fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print('%-28s %8.3f ms' % (name, e0.elapsed_time(e1) / reps), flush=True)
    return out

F, T, D, K, I, U = 513, 500, 8, 2, 100, 8
yb = torch.stack([torch.from_numpy(synth.noise_stft(F, T, D, seed=50 + u)) for u in range(U)]).cuda()
ib = torch.stack([torch.from_numpy(synth.init_affiliation(F, K, T, seed=7 + u)) for u in range(U)]).cuda()
model = timed('fit 100 it (8 utterances)', lambda: CACGMMTrainer().fit(yb, initialization=ib, iterations=I))
aff = timed('predict', lambda: model.predict(yb))
Yb = timed('transpose', lambda: yb.transpose(-1, -2).contiguous())
psd = timed('PSD', lambda: E.get_power_spectral_density_matrix(Yb, aff))
atf = timed('PCA vector', lambda: E.get_pca_vector(psd[..., 0, :, :]))
noise = psd[..., 1, :, :].contiguous()
w = timed('MVDR vector', lambda: E.get_mvdr_vector(atf, noise))
timed('apply', lambda: E.apply_beamforming_vector(w, Yb))
