"""Frame split of em_ws_kernel (PBB_TSPLIT = S): 100-iteration C2-shaped fit for F bins, S = 1, 2, 4 and the
library's own choice.  python scripts/tsplit_sweep.py"""
import os, sys
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
tr = CACGMMTrainer()
T, D, K, I = 500, 8, 3, 100
for F in (17, 33, 65, 129, 172, 257, 513):
    y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
    init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
    row = []
    for S in ('1', '2', '4', None):
        if S is None:
            os.environ.pop('PBB_TSPLIT', None)
        else:
            os.environ['PBB_TSPLIT'] = S
        for _ in range(2): tr.fit(y, initialization=init, iterations=I)
        ts = []
        for _ in range(6):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); tr.fit(y, initialization=init, iterations=I); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(min(ts))
    print(f'F={F:4d}: S=1 {row[0]:.3f}  S=2 {row[1]:.3f}  S=4 {row[2]:.3f}  auto {row[3]:.3f} ms per 100-iteration fit', flush=True)
