"""Host-side logic of the bin-sharded multi-GPU path, exercised with two gloo
ranks on CPU tensors (world_size 2, 127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_bin_shards_cover_and_balance():
    from pb_bss_b200.parallel import bin_shards
    for F in (1, 7, 64, 129, 257, 513):
        for ws in (1, 2, 3, 4, 8):
            sh = bin_shards(F, ws)
            assert len(sh) == ws and sh[0][0] == 0 and sh[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(sh[:-1], sh[1:]))
            sizes = [h - l for l, h in sh]
            assert max(sizes) - min(sizes) <= 1
    assert [h - l for l, h in bin_shards(513, 8)] == [65] + [64] * 7


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, F, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from pb_bss_b200.parallel import all_gather_bins, bin_shards, local_bins, world
        assert world() == (rank, ws)
        full = torch.arange(F * 3 * 5, dtype=torch.float64).reshape(F, 3, 5)
        lo, hi = local_bins(F)
        assert (lo, hi) == bin_shards(F, ws)[rank]
        got = all_gather_bins(full[lo:hi].clone(), F)
        ok = bool(torch.equal(got, full))
        # integer payloads (mappings) travel the same way
        m = torch.arange(F * 2, dtype=torch.int64).reshape(F, 2)
        ok = ok and bool(torch.equal(all_gather_bins(m[lo:hi].clone(), F), m))
        # frequency-tied weights: global mean over bins from the ranks' local means
        from pb_bss_b200.parallel import mean_over_all_bins
        aff = torch.rand(F, 3, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
        got = mean_over_all_bins(aff[lo:hi].mean(0), hi - lo, F)
        ok = ok and bool(torch.allclose(got, aff.mean(0), rtol=1e-13, atol=0))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('F', [513, 9])
def test_all_gather_bins_two_gloo_ranks(F):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=10) for _ in range(2))
    assert results == [(0, True), (1, True)]


def test_single_rank_degrades_gracefully():
    from pb_bss_b200.parallel import all_gather_bins, local_bins, world
    assert world() == (0, 1)
    assert local_bins(513) == (0, 513)
    x = torch.randn(7, 2, 3)
    assert all_gather_bins(x, 7) is x
    from pb_bss_b200.parallel import mean_over_all_bins
    x0 = x[0]
    assert mean_over_all_bins(x0, 7, 7) is x0
