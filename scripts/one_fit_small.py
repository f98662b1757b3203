"""One device-resident fit of a rank's shard of C3 (65 bins, T=500, D=8, K=3, 100 iterations): em_sticky_kernel."""
import sys, time
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
F, T, D, K, I = 65, 500, 8, 3, 100
y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
tr = CACGMMTrainer()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.fit(y, initialization=init, iterations=I)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('F=%d: %.3f ms per %d-iteration fit' % (F, dt * 1e3, I), flush=True)
