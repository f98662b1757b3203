"""C1 (cACGMM F=129 T=200 D=4 K=2, 20 iterations) with and without the frame split; also 100 iterations."""
import os, sys
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
F, T, D, K = 129, 200, 4, 2
y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda(); init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
tr = CACGMMTrainer()
for I in (20, 100):
    for S in ('1', '2', None):
        if S is None: os.environ.pop('PBB_TSPLIT', None)
        else: os.environ['PBB_TSPLIT'] = S
        for _ in range(3): tr.fit(y, initialization=init, iterations=I)
        ts = []
        for _ in range(9):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); tr.fit(y, initialization=init, iterations=I); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f'C1 I={I} PBB_TSPLIT={S}: min {ts[0]:.3f} median {ts[4]:.3f} ms', flush=True)
