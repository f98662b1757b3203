"""A/B of the D = 8 persistent EM kernels (PBB_EM_KERNEL = ls | ws | single): small parity cases against the
oracle, then the C2 fit timed with CUDA events.  python scripts/ab_kernels.py [ls ws single] [--quick]"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np
    import torch
    sys.path.insert(0, '.')
    from oracle import synth, pb_bss_oracle as O
    from pb_bss_b200.distribution import CACGMMTrainer
    tag = os.environ.get('PBB_EM_KERNEL', 'ls')
    quick = '--quick' in sys.argv
    tr = CACGMMTrainer()
    for (F, T, K, I) in ((5, 64, 3, 3), (21, 333, 3, 9), (10, 128, 2, 5), (6, 500, 4, 4), (3, 1100, 3, 4), (40, 500, 3, 30)):
        y, _ = synth.structured_stft(F, T, 8, K, seed=3)
        init = synth.init_affiliation(F, K, T, seed=7)
        m = tr.fit(y, initialization=init, iterations=I)
        ref = O.cacgmm_fit(y, init, I)
        err_w = np.abs(m.weight - ref['weight']).max()
        cov_ref = np.einsum('...de,...e,...fe->...df', ref['eigenvectors'], ref['eigenvalues'], ref['eigenvectors'].conj())
        err_c = np.abs(m.cacg.covariance - cov_ref).max()
        print(f'[{tag}] F={F} T={T} K={K} I={I}: max |dw| {err_w:.2e}  max |dcov| {err_c:.2e}', flush=True)
        assert err_w < 1e-7 and err_c < 1e-6, 'parity'
    # pinned host input (streamed upload)
    y, _ = synth.structured_stft(24, 500, 8, 3, seed=5)
    init = synth.init_affiliation(24, 3, 500, seed=7)
    a = tr.fit(y, initialization=init, iterations=6)
    b = tr.fit(torch.from_numpy(y).pin_memory(), initialization=torch.from_numpy(init).pin_memory(), iterations=6)
    print(f'[{tag}] streamed upload: max |dw| {np.abs(a.weight - b.weight.numpy()).max():.2e}', flush=True)
    if not quick:
        F, T, D, K, I = 513, 500, 8, 3, 100
        y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
        init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
        for _ in range(3):
            tr.fit(y, initialization=init, iterations=I)
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            m = tr.fit(y, initialization=init, iterations=I)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print('[%s] C2 fit: min %.3f  median %.3f ms  checksum %.15e' % (
            tag, ts[0], ts[len(ts) // 2], float(m.cacg.covariance_eigenvalues.sum())), flush=True)
else:
    kernels = [a for a in sys.argv[1:] if not a.startswith('--')] or ['ls', 'ws']
    flags = [a for a in sys.argv[1:] if a.startswith('--')]
    for k in kernels:
        e = dict(os.environ)
        e['PBB_EM_KERNEL'] = k
        r = subprocess.run(['timeout', '150', sys.executable, __file__, 'child'] + flags, env=e)
        print(f'[{k}] exit code {r.returncode}', flush=True)
