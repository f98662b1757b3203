"""Small correctness + timing check of the warp-specialised kernel against the classic persistent kernel."""
import os, sys, subprocess, time
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np, torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200.distribution import CACGMMTrainer
    out = {}
    for (F, T, D, K, I) in [(3, 130, 8, 3, 3), (40, 333, 8, 3, 12), (513, 500, 8, 3, 100), (64, 500, 8, 2, 20), (33, 257, 8, 4, 7)]:
        y, _ = synth.structured_stft(F, T, D, K, seed=3)
        init = synth.init_affiliation(F, K, T, seed=7)
        yd, idv = torch.from_numpy(y).cuda(), torch.from_numpy(init).cuda()
        m = CACGMMTrainer().fit(yd, initialization=idv, iterations=I)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            m = CACGMMTrainer().fit(yd, initialization=idv, iterations=I); torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print((F, T, D, K, I), 'min %.3f ms' % (min(ts) * 1e3), flush=True)
        out[str((F, T, D, K, I))] = m.cacg.covariance_eigenvalues.cpu().numpy()
    np.savez(sys.argv[2], **out)
else:
    import numpy as np
    for tag, env in (('ws', {}), ('classic', {'PBB_NO_WS': '1'})):
        e = dict(os.environ); e.update(env)
        print('==', tag, flush=True)
        r = subprocess.run(['timeout', '120', sys.executable, __file__, 'child', f'/tmp/ws_{tag}.npz'], env=e)
        print('rc', r.returncode, flush=True)
    a, b = np.load('/tmp/ws_ws.npz'), np.load('/tmp/ws_classic.npz')
    for k in a.files:
        print(k, 'max |d eigenvalues|', float(np.abs(a[k] - b[k]).max()), 'equal', bool((a[k] == b[k]).all()))
