"""Frame split of the single-role persistent kernel (PBB_TSPLIT = S) on config 4 (complex Watson, F=257 T=1000 D=6 K=4,
50 iterations) and on smaller bin counts.  python scripts/tsplit_sweep_cw.py"""
import os, sys
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CWMMTrainer
tr = CWMMTrainer()
T, D, K, I = 1000, 6, 4, 50
for F in (33, 65, 129, 257, 513):
    y = torch.from_numpy(synth.noise_stft(F, T, D, seed=4)).cuda()
    init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
    row = []
    for S in ('1', '2', '4', '8', None):
        if S is None:
            os.environ.pop('PBB_TSPLIT', None)
        else:
            os.environ['PBB_TSPLIT'] = S
        for _ in range(2): tr.fit(y, initialization=init, iterations=I)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); tr.fit(y, initialization=init, iterations=I); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(min(ts))
    print(f'CW F={F:4d}: S=1 {row[0]:.3f}  S=2 {row[1]:.3f}  S=4 {row[2]:.3f}  S=8 {row[3]:.3f}  auto {row[4]:.3f} ms per 50-iteration fit', flush=True)
