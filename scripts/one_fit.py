"""One device-resident C2 cACGMM fit (for ncu): python scripts/one_fit.py [iterations] [c64]"""
import sys
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
I = int(sys.argv[1]) if len(sys.argv) > 1 else 10
F, T, D, K = 513, 500, 8, 3
y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
if len(sys.argv) > 2: y = y.to(torch.complex64)
init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
m = CACGMMTrainer().fit(y, initialization=init, iterations=I)
torch.cuda.synchronize()
print('done', float(m.weight.sum()))
