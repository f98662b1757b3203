"""Device-resident timings of the BASELINE.json configs other than C2 (C1, C4, C5-like per-utterance pipeline)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer, CWMMTrainer
from pb_bss_b200 import extraction as E

def timed(fn, reps=5):
    """median over reps of the device time of one call (a single slow rep -- host jitter on a shared box -- does not
    move it)"""
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], out

# C1
F, T, D, K, I = 129, 200, 4, 2, 20
y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda(); init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
ms, _ = timed(lambda: CACGMMTrainer().fit(y, initialization=init, iterations=I))
print(f'C1 cACGMM F=129 T=200 D=4 K=2 I=20: {ms:.3f} ms -> {I/ms*1e3:.0f} it/s')
# C4
F, T, D, K, I = 257, 1000, 6, 4, 50
y = torch.from_numpy(synth.noise_stft(F, T, D, seed=4)).cuda(); init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
tr = CWMMTrainer()
ms, m = timed(lambda: tr.fit(y, initialization=init, iterations=I))
print(f'C4 CWMM  F=257 T=1000 D=6 K=4 I=50: {ms:.3f} ms -> {I/ms*1e3:.0f} it/s, {I*F*T/ms*1e3:.3e} frames*bins/s')
# C5-like: one utterance K=2 cACGMM + PSD + Souden-MVDR / PCA-MVDR
F, T, D, K, I = 513, 500, 8, 2, 100
y = torch.from_numpy(synth.noise_stft(F, T, D, seed=5)).cuda(); init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
def c5():
    model = CACGMMTrainer().fit(y, initialization=init, iterations=I)
    aff = model.predict(y)
    Y = y.transpose(-1, -2).contiguous()
    psd = E.get_power_spectral_density_matrix(Y, aff)
    w = E.get_mvdr_vector(E.get_pca_vector(psd[:, 0]), psd[:, 1].contiguous())
    return E.apply_beamforming_vector(w, Y)
ms, _ = timed(c5)
print(f'C5 per utterance (K=2 fit 100 it + predict + PSD + PCA + MVDR + apply): {ms:.3f} ms')
ms, _ = timed(lambda: CACGMMTrainer().fit(y, initialization=init, iterations=I))
print(f'   of which fit: {ms:.3f} ms -> {I/ms*1e3:.0f} it/s')
# C5 as one rank sees it: 8 utterances (64 over 8 GPUs) in ONE batched call, (U, F, T, D) -> U * F independent bins
U = 8
yb = torch.stack([torch.from_numpy(synth.noise_stft(F, T, D, seed=50 + u)) for u in range(U)]).cuda()
ib = torch.stack([torch.from_numpy(synth.init_affiliation(F, K, T, seed=7 + u)) for u in range(U)]).cuda()
def c5_batch():
    model = CACGMMTrainer().fit(yb, initialization=ib, iterations=I)
    aff = model.predict(yb)
    Yb = yb.transpose(-1, -2).contiguous()
    psd = E.get_power_spectral_density_matrix(Yb, aff)
    w = E.get_mvdr_vector(E.get_pca_vector(psd[..., 0, :, :]), psd[..., 1, :, :].contiguous())
    return E.apply_beamforming_vector(w, Yb)
try:
    ms, _ = timed(c5_batch, reps=3)
    print(f'C5 one rank, {U} utterances batched (K=2 fit 100 it + predict + PSD + PCA + MVDR + apply): {ms:.3f} ms = {ms/U:.3f} ms per utterance')
except Exception as e:  # noqa: BLE001
    print('C5 batched pipeline failed:', repr(e))
ms, _ = timed(lambda: CACGMMTrainer().fit(yb, initialization=ib, iterations=I), reps=3)
print(f'   of which fit: {ms:.3f} ms = {ms/U:.3f} ms per utterance -> {U*I/ms*1e3:.0f} it/s')
a = torch.from_numpy(synth.pos_def_hermitian(1539, 8, 8)).cuda(); b = torch.from_numpy(synth.pos_def_hermitian(1539, 8, 8, seed=2)).cuda()
from pb_bss_b200.extraction.linalg import eigh
ms, _ = timed(lambda: eigh(a)); print(f'eigh 1539 x 8x8: {ms*1e3:.1f} us')
ms, _ = timed(lambda: E.get_gev_vector(a, b)); print(f'gev  1539 x 8x8: {ms*1e3:.1f} us')
# coupled EM (section 8 f2): frequency-tied weights and the inline permutation alignment at the C2 shape, 20 iterations
from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment
F, T, D, K, I = 513, 500, 8, 3, 20
y = torch.from_numpy(synth.structured_stft(F, T, D, K, seed=5)[0]).cuda(); init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
ms, _ = timed(lambda: CACGMMTrainer().fit(y, initialization=init, iterations=I, weight_constant_axis=(-3,)), reps=3)
print(f'coupled cACGMM, tied weights (-3,), F=513 T=500 D=8 K=3 I={I}: {ms:.3f} ms = {ms/I:.3f} ms per iteration')
al = DHTVPermutationAlignment.from_stft_size(1024)
ms, _ = timed(lambda: CACGMMTrainer().fit(y, initialization=init, iterations=I, weight_constant_axis=(-3,),
                                          inline_permutation_aligner=al), reps=3)
print(f'coupled cACGMM, tied weights + inline DHTV alignment, I={I}: {ms:.3f} ms = {ms/I:.3f} ms per iteration')
