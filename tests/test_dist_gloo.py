"""World-size-2 test of the N>1 host logic on CPU (gloo): every global batch is
split over the ranks, each rank contributes un-normalised sums, one all-reduce,
identical result on every rank == the single-process result.  The per-rank
partial sums come from the oracle here (no GPU in this test); on GPUs they come
from enf_negll_grad_partial and the all-reduce is the library's ncclAllReduce
(enf_negll_grad_group), exercised by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import enf_b200 as E
    from chains import build, flat_grads
    from oracle import enf_oracle as O
    D, N, nbatches = 3, 203, 4
    f = build(O, ["cc", "jo", "hh2", "ss"], D, np.random.default_rng(5))
    X = np.random.default_rng(6).standard_normal((D, N))
    ranges = E.batch_ranges(N, nbatches)
    mine = E.dist.shard_batches(ranges, rank, world)
    results = []
    for (s, e), (ls, le) in zip(ranges, mine):
        nl = le - ls
        # this rank's un-normalised sums: N_local * (negll, grads) of its column block
        if nl > 0:
            v, g = O.mvnormal_negll_trafograd(f, X[:, ls:le], zygote_primal=False)
            flat = np.concatenate([[v * nl]] + [a.ravel() * nl for _, a in flat_grads(g, f)])
        else:
            flat = np.zeros(1 + sum(a.size for _, a in flat_grads(O.mvnormal_negll_trafograd(f, X[:, :1])[1], f)))
        tot, n_glob = E.dist.allreduce_sums(flat, nl)
        assert n_glob == e - s
        results.append(tot / n_glob)
    if rank == 0:
        ref = []
        for (s, e) in ranges:
            v, g = O.mvnormal_negll_trafograd(f, X[:, s:e], zygote_primal=False)
            ref.append(np.concatenate([[v]] + [a.ravel() for _, a in flat_grads(g, f)]))
        err = max(np.max(np.abs(a - b) / (np.abs(b) + 1)) for a, b in zip(results, ref))
        out.put(err)
    dist.destroy_process_group()


def test_sharded_batches_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-12


def _rendezvous_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import enf_b200 as E
    payload = bytes(range(128)) if rank == 0 else b""
    got = E.dist.exchange_bytes(payload, rank, world, "127.0.0.1", port, timeout=30.0)
    out.put((rank, got == bytes(range(128))))


def test_torch_free_rendezvous_world3():
    """The NCCL unique id travels from rank 0 to the other ranks over a plain TCP socket (no torch.distributed on the
    library's path): three processes, rank 0 started last so that the others have to retry."""
    ctx = mp.get_context("spawn")
    port, out = _free_port(), None
    out = ctx.Queue()
    procs = [ctx.Process(target=_rendezvous_worker, args=(r, 3, port, out)) for r in (1, 2, 0)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=60) for _ in range(3))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: True, 1: True, 2: True}
