"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): member
sharding and the one collective of the design, the all-reduce of the POD
snapshot Gram matrix  G = sum_m X_m^T M X_m  (SURVEY.md 8e)."""
import os
import socket

import numpy as np
import scipy.sparse as sps
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dolfin_navier_scipy_b200 import ensemble as ens


def test_shard_members_partitions_exactly():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = ens.shard_members(n, r, world)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
            assert cover == list(range(n))
            sizes = [np.subtract(*ens.shard_members(n, r, world)[::-1])
                     for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _snapshots(nmembers, nv=40, ns=5):
    rng = np.random.default_rng(7)
    X = rng.standard_normal((ns, nv, nmembers))
    M = sps.diags(rng.uniform(1., 2., nv)).tocsr()
    return X, M


def _worker(rank, world, port, nmembers, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    X, M = _snapshots(nmembers)
    lo, hi = ens.shard_members(nmembers, rank, world)
    G = np.zeros((X.shape[0], X.shape[0]))
    for m in range(lo, hi):
        Xm = X[:, :, m].T                      # (nv, ns)
        G += Xm.T@(M@Xm)
    Gt = ens.allreduce_gram(torch.from_numpy(G))
    out[rank] = Gt.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def test_gram_allreduce_two_ranks_gloo():
    world, nmembers = 2, 5
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, nmembers, out), nprocs=world,
             join=True)
    X, M = _snapshots(nmembers)
    Gref = sum(X[:, :, m]@(M@X[:, :, m].T) for m in range(nmembers))
    for r in range(world):
        assert np.allclose(out[r], Gref, rtol=1e-13, atol=0)
    assert np.array_equal(out[0], out[1])
