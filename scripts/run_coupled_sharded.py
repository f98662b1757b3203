"""Frequency-tied weights + inline permutation alignment with the bins sharded over N ranks
(torchrun): per iteration one all-reduce of the (K, T) weight sums and one all-gather of the
affiliations.  Every rank checks its slice against the same fit done on all bins locally.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_coupled_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from pb_bss_b200.distribution import CACGMMTrainer  # noqa: E402
from pb_bss_b200.parallel import bin_shards  # noqa: E402
from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment  # noqa: E402


def main():
    rank, ws = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if ws > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    F, T, D, K, iters = 65, 120, 4, 2, 6
    y, _ = synth.structured_stft(F, T, D, K, seed=11)
    init = synth.init_affiliation(F, K, T, seed=7)
    lo, hi = bin_shards(F, ws)[rank]
    yd, idv = torch.from_numpy(y).cuda(), torch.from_numpy(init).cuda()
    worst = 0.0
    for axis, inline in (((-3,), False), ((-3, -1), False), ((-3,), True)):
        def aligner():
            return DHTVPermutationAlignment(stft_size=128, segment_start=20, segment_width=20, segment_shift=5,
                                            main_iterations=5, sub_iterations=2) if inline else None
        full = CACGMMTrainer().fit(yd, initialization=idv, iterations=iters, weight_constant_axis=axis,
                                   inline_permutation_aligner=aligner())
        part = CACGMMTrainer().fit(yd[lo:hi].contiguous(), initialization=idv[lo:hi].contiguous(), iterations=iters,
                                   weight_constant_axis=axis, inline_permutation_aligner=aligner(),
                                   total_bins=F, bin_group=None)
        err_w = float((part.weight - full.weight).abs().max())
        err_c = float((part.cacg.covariance - full.cacg.covariance[lo:hi]).abs().max())
        worst = max(worst, err_w, err_c)
        print(f'rank {rank}/{ws} axis {axis} inline {inline}: |dw| {err_w:.2e} |dcov| {err_c:.2e}', flush=True)
    t = torch.tensor([worst], device='cuda')
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = float(t) < 1e-9
        print('sharded coupled fit', 'OK' if ok else 'MISMATCH', float(t), flush=True)
        if not ok:
            sys.exit(1)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
