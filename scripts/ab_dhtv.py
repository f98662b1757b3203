"""A/B of the DHTV alignment kernels: cluster / DSMEM kernel (default) vs grid-barrier kernel (PBB_DHTV_COOP=1) vs the
launch pair per iteration (PBB_DHTV_MULTI=1): mappings must be identical; time per calculate_mapping at C3 size."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np, torch
    sys.path.insert(0, '.')
    from oracle import synth
    from pb_bss_b200.distribution import CACGMMTrainer
    from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment
    out = {}
    for (F, T, K, seed) in ((513, 500, 3, 5), (257, 300, 2, 6), (513, 200, 4, 7)):
        y, _ = synth.structured_stft(F, T, 8, K, seed=seed)
        init = synth.init_affiliation(F, K, T, seed=7)
        m = CACGMMTrainer().fit(torch.from_numpy(y).cuda(), initialization=torch.from_numpy(init).cuda(), iterations=15)
        mask = m.predict(torch.from_numpy(y).cuda()).permute(1, 0, 2).contiguous()
        al = DHTVPermutationAlignment.from_stft_size(2 * (F - 1))
        for metric in ('cos', 'euclidean'):
            al.similarity_metric = metric
            out[f'{F}_{K}_{metric}'] = al.calculate_mapping(mask).cpu().numpy()
        al.similarity_metric = 'cos'
        if F == 513 and K == 3:
            for _ in range(3): al.calculate_mapping(mask)
            ts = []
            for _ in range(9):
                torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); al.calculate_mapping(mask); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            print(f'[{sys.argv[2]}] calculate_mapping F=513 T=500 K=3: min {ts[0]:.3f} median {ts[4]:.3f} ms', flush=True)
    np.savez(sys.argv[3], **out)
else:
    import numpy as np
    res = {}
    for tag, env in (('cluster', {}), ('coop', {'PBB_DHTV_COOP': '1'}), ('multi', {'PBB_DHTV_MULTI': '1'})):
        e = dict(os.environ); e.update(env)
        path = f'/tmp/ab_dhtv_{tag}.npz'
        subprocess.run(['timeout', '200', sys.executable, __file__, 'child', tag, path], env=e, check=True)
        res[tag] = np.load(path)
    for k in res['cluster'].files:
        same_coop = np.array_equal(res['cluster'][k], res['coop'][k])
        same_multi = np.array_equal(res['cluster'][k], res['multi'][k]) if k.endswith('cos') else None
        nonid = int((res['cluster'][k] != np.arange(res['cluster'][k].shape[0])[:, None]).any(0).sum())
        print(f'{k}: cluster == coop {same_coop}, cluster == multi {same_multi}, bins with a non-identity mapping {nonid}')
        assert same_coop
