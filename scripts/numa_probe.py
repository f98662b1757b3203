"""Does the NUMA node of the pinned buffer matter for kernel reads over PCIe?"""
import os, sys, subprocess, time
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    cpus = sys.argv[2]
    if cpus != 'any':
        lo, hi = cpus.split('-'); os.sched_setaffinity(0, range(int(lo), int(hi) + 1))
    import torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200 import _lib, _device
    F, T, D = 513, 500, 8
    y_pin = torch.from_numpy(synth.noise_stft(F, T, D)).pin_memory()
    z = torch.empty(F, D, T, dtype=torch.complex128, device='cuda')
    lib = _lib.load(); st = _device.stream_ptr()
    def f():
        _lib.check(lib.pbb_normalize_observation(y_pin.data_ptr(), z.data_ptr(), F, T, D, _lib.PBB_C128, 1, st), 'n')
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ms = min(ts) * 1e3
    print('cpus %-10s zero-copy read %.3f ms = %.1f GB/s' % (cpus, ms, y_pin.numel() * 16 / ms / 1e6), flush=True)
else:
    print(subprocess.run('nvidia-smi topo -m | head -8; lscpu | grep -i "numa\\|^CPU(s)"', shell=True, capture_output=True, text=True).stdout)
    n = os.cpu_count()
    for cpus in ('any', f'0-{n // 4 - 1}', f'{n // 4}-{n // 2 - 1}', f'{n // 2}-{3 * n // 4 - 1}', f'{3 * n // 4}-{n - 1}', 'any'):
        subprocess.run([sys.executable, __file__, 'child', cpus])
