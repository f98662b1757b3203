"""One device-resident config-4 complex-Watson fit (for ncu / phase builds): python scripts/one_fit_cw.py [iterations] [F]"""
import sys, time
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CWMMTrainer
I = int(sys.argv[1]) if len(sys.argv) > 1 else 50
F = int(sys.argv[2]) if len(sys.argv) > 2 else 257
T, D, K = 1000, 6, 4
y = torch.from_numpy(synth.noise_stft(F, T, D, seed=4)).cuda()
init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
tr = CWMMTrainer()
tr.fit(y, initialization=init, iterations=I)  # (same launch shape as the timed fits: ncu -c 1 captures this one)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = tr.fit(y, initialization=init, iterations=I)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('CWMM F=%d T=%d D=%d K=%d: %.3f ms per %d-iteration fit (%.0f it/s)' % (F, T, D, K, dt * 1e3, I, I / dt), flush=True)
