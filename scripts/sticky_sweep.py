"""em_sticky_kernel vs em_ws_kernel (+ frame split) for few bins: 100-iteration C2-shaped fit.  python scripts/sticky_sweep.py"""
import os, sys
import torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
tr = CACGMMTrainer()
D, K, I = 8, 3, 100
def timed(y, init):
    for _ in range(2): tr.fit(y, initialization=init, iterations=I)
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); tr.fit(y, initialization=init, iterations=I); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
for (F, T) in ((17, 500), (33, 500), (65, 500), (129, 500), (148, 500), (257, 500), (65, 250), (129, 1000), (64, 1100)):
    y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
    init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
    row = {}
    os.environ['PBB_STICKY'] = '0'; row['ws'] = timed(y, init)
    for S in ('1', '2', '4'):
        os.environ['PBB_STICKY'] = S
        try:
            row['S' + S] = timed(y, init)
        except Exception as e:  # noqa: BLE001
            row['S' + S] = float('nan')
    os.environ.pop('PBB_STICKY'); row['auto'] = timed(y, init)
    print(f'F={F:4d} T={T:5d}: ' + '  '.join(f'{k} {v:.3f}' for k, v in row.items()) + '  ms per 100-iteration fit', flush=True)
