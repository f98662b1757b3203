"""Kernel-level view of the streamed upload: loader vs EM kernel durations for a few settings."""
import os, sys, time, subprocess
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200 import _lib
    from pb_bss_b200.distribution import CACGMMTrainer
    F, T, D, K, I = 513, 500, 8, 3, 100
    y_pin = torch.from_numpy(synth.noise_stft(F, T, D)).pin_memory()
    init_pin = torch.from_numpy(synth.init_affiliation(F, K, T)).pin_memory()
    lib = _lib.load(); tr = CACGMMTrainer()
    for _ in range(3):
        tr.fit(y_pin, initialization=init_pin, iterations=I)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tr.fit(y_pin, initialization=init_pin, iterations=I); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print('e2e min %.3f ms' % (min(ts) * 1e3), flush=True)
    lib.pbb_profile_reset(); lib.pbb_profile_enable(1)
    tr.fit(y_pin, initialization=init_pin, iterations=I); torch.cuda.synchronize()
    lib.pbb_profile_dump(); lib.pbb_profile_enable(0)
else:
    A = {'PBB_WAVE_C': '8', 'PBB_ORDER_CAP': '292'}
    B = {'PBB_WAVE_C': '15', 'PBB_ORDER_CAP': '467'}
    C = {'PBB_WAVE_C': '12', 'PBB_ORDER_CAP': '400'}
    for env in (A, B, C, A, B, C, A, B, C):
        print('==', env, flush=True)
        e = dict(os.environ); e.update(env)
        subprocess.run(['timeout', '100', sys.executable, __file__, 'child'], env=e)
