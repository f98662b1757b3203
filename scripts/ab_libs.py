"""A/B timing of the C2 fit for several builds of the library: python scripts/ab_libs.py lib1.so lib2.so ..."""
import os, sys, subprocess
if sys.argv[1] == 'child':
    import time, torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200.distribution import CACGMMTrainer
    F, T, D, K, I = 513, 500, 8, 3, 100
    y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
    init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
    tr = CACGMMTrainer()
    for _ in range(3): tr.fit(y, initialization=init, iterations=I)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); tr.fit(y, initialization=init, iterations=I); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    m = tr.fit(y, initialization=init, iterations=I)
    print('%-50s min %.3f  median %.3f ms  checksum %.15e' % (os.environ.get('PBB_LIB', 'default'), ts[0], ts[len(ts) // 2], float(m.cacg.covariance_eigenvalues.sum())), flush=True)
else:
    for lib in sys.argv[1:]:
        e = dict(os.environ)
        if lib != 'default': e['PBB_LIB'] = lib
        subprocess.run(['timeout', '120', sys.executable, __file__, 'child'], env=e)
