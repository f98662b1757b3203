"""Ad-hoc timing of the C2 cACGMM fit (device-resident), for tuning."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer

F, T, D, K, I = 513, 500, 8, 3, 100
y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
tr = CACGMMTrainer()
for fpb in (0, 64, 96, 128, 160, 256, 512):
    for dt in (torch.complex128, torch.complex64):
        yy = y.to(dt)
        tr.fit(yy, initialization=init, iterations=3, frames_per_block=fpb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m = tr.fit(yy, initialization=init, iterations=I, frames_per_block=fpb); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f'fpb={fpb:4d} {str(dt):18s} fit {ms:8.3f} ms  {I/ms*1e3:9.1f} it/s  {I*F*T/ms*1e3:.3e} frames*bins/s', flush=True)
