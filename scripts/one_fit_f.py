"""One device-resident cACGMM fit with F bins (phase-timing builds): python scripts/one_fit_f.py F [iterations]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
F = int(sys.argv[1]); I = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T, D, K = 500, 8, 3
y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
tr = CACGMMTrainer()
tr.fit(y, initialization=init, iterations=2)
torch.cuda.synchronize(); t0 = time.perf_counter()
m = tr.fit(y, initialization=init, iterations=I)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('F=%d: %.3f ms, %.3f us per bin-iteration' % (F, dt * 1e3, dt * 1e6 / (F * I)), flush=True)
