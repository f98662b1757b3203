"""End to end through the drop-in API: NumPy in -> NumPy out (page-locked for the call or staged through a pageable
copy), pinned tensors in -> pinned model out, device resident.  python scripts/e2e_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
F, T, D, K, I = 513, 500, 8, 3, 100
y = synth.noise_stft(F, T, D); init = synth.init_affiliation(F, K, T)
tr = CACGMMTrainer()
def timed(name, fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort()
    print('%-46s min %.3f  median %.3f ms' % (name, ts[0] * 1e3, ts[len(ts) // 2] * 1e3), flush=True)
    return out
a = timed('numpy -> numpy (host-registered for the call)', lambda: tr.fit(y, initialization=init, iterations=I))
os.environ['PBB_NO_HOST_REGISTER'] = '1'
b = timed('numpy -> numpy (pageable copies)', lambda: tr.fit(y, initialization=init, iterations=I))
del os.environ['PBB_NO_HOST_REGISTER']
yp, ip = torch.from_numpy(y).pin_memory(), torch.from_numpy(init).pin_memory()
c = timed('pinned tensors -> pinned model', lambda: tr.fit(yp, initialization=ip, iterations=I))
yd, idv = yp.cuda(), ip.cuda()
d = timed('device resident', lambda: tr.fit(yd, initialization=idv, iterations=I))
print('identical models:', np.array_equal(a.weight, b.weight), np.array_equal(a.weight, c.weight.numpy()),
      np.array_equal(a.cacg.covariance_eigenvalues, d.cacg.covariance_eigenvalues.cpu().numpy()))
