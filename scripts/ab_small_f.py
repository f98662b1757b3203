"""Fit time of the D = 8 persistent kernels for few bins (bin-sharded C3: 65 bins per GPU), PBB_EM_KERNEL = ls | ws."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200.distribution import CACGMMTrainer
    tag = os.environ.get('PBB_EM_KERNEL', 'ls')
    tr = CACGMMTrainer()
    for F in (33, 65, 129, 257, 513, 1026):
        T, D, K, I = 500, 8, 3, 100
        y = torch.from_numpy(synth.noise_stft(F, T, D)).cuda()
        init = torch.from_numpy(synth.init_affiliation(F, K, T)).cuda()
        for _ in range(2): tr.fit(y, initialization=init, iterations=I)
        ts = []
        for _ in range(6):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); tr.fit(y, initialization=init, iterations=I); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f'[{tag}] F={F:5d}: {ts[0]:.3f} ms per 100-iteration fit  ({ts[0] * 10:.1f} us per iteration)', flush=True)
else:
    for k in (sys.argv[1:] or ['ls', 'ws']):
        e = dict(os.environ); e['PBB_EM_KERNEL'] = k
        subprocess.run(['timeout', '120', sys.executable, __file__, 'child'], env=e)
