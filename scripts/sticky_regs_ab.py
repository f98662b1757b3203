"""A/B of the sticky kernel's register split (PBB_LIB = variant build): F = 65 / 17, T = 500, 100 iterations."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200.distribution import CACGMMTrainer
    tr = CACGMMTrainer()
    for F in (65, 17):
        y = torch.from_numpy(synth.noise_stft(F, 500, 8)).cuda(); init = torch.from_numpy(synth.init_affiliation(F, 3, 500)).cuda()
        for _ in range(3): m = tr.fit(y, initialization=init, iterations=100)
        ts = []
        for _ in range(8):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); m = tr.fit(y, initialization=init, iterations=100); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('[%s] F=%d: min %.3f ms  checksum %.12e' % (sys.argv[2], F, min(ts), float(m.cacg.covariance_eigenvalues.sum())), flush=True)
else:
    for tag, lib in (('208/48', None), ('200/56', 'pb_bss_b200/libpbb_s200.so')):
        e = dict(os.environ)
        if lib: e['PBB_LIB'] = lib
        subprocess.run(['timeout', '60', sys.executable, __file__, 'child', tag], env=e)
