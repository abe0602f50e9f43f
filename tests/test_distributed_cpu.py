"""world_size-2 gloo test of the N>1 host logic: path-range sharding + one all-reduce of the double
moment vector + redundant finalisation.  The per-shard moments come from the CPU oracle here (no GPU
in this container); on GPUs the same helpers wrap hw1f_bond_curve_moments / _finish."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL = 3001          # ragged on purpose: 1501 + 1500
SEED = 77


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hw1f_b200 as hw
    from oracle_lib import Oracle
    par = hw.package.parallel
    o = Oracle()
    first, n = par.shard_paths(N_TOTAL, rank, world)
    s, q = o.bond_curve_sums(SEED, n, first_path=first)
    moments = torch.from_numpy(np.concatenate([s, q]))
    par.allreduce_moments(moments)
    P, f = o.curve_finalize(moments[:101].numpy(), N_TOTAL)
    out[rank] = (first, n, moments.numpy().copy(), P, f)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    sys.path.insert(0, ROOT)
    import hw1f_b200 as hw
    par = hw.package.parallel
    for n_total, world in ((1 << 20, 8), (3001, 2), (10, 4), (7, 8)):
        got = [par.shard_paths(n_total, r, world) for r in range(world)]
        assert got[0][0] == 0 and sum(n for _, n in got) == n_total
        for (f0, n0), (f1, _) in zip(got, got[1:]):
            assert f0 + n0 == f1
    with pytest.raises(ValueError):
        par.shard_paths(10, 2, 2)
    with pytest.raises(TypeError):
        par.allreduce_moments(torch.zeros(3))


def test_two_rank_allreduce_matches_single_process(oracle):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    s, q = oracle.bond_curve_sums(SEED, N_TOTAL)
    P, f = oracle.curve_finalize(s, N_TOTAL)
    assert out[0][:2] == (0, 1501) and out[1][:2] == (1501, 1500)
    for r in range(world):
        assert np.allclose(out[r][2], np.concatenate([s, q]), rtol=1e-12)
        assert (out[r][3] == P).all() and (out[r][4] == f).all()     # every rank finalises identically
    assert (out[0][2] == out[1][2]).all()
