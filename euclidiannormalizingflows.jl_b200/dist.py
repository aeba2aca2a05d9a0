"""Multi-GPU host logic: one process per GPU, samples sharded by columns.

Forward / inverse / ladj need no communication (columns are independent,
parameters are replicated).  The loss+gradient step has exactly one exchange:
an all-reduce (sum, float64) of the raw loss and parameter-gradient sums
(SURVEY §8e).  On GPUs that is an ncclAllReduce issued by the library on its own
stream (enf_negll_grad_group); `torch.distributed` is used only as plumbing to
hand the NCCL unique id to every rank.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _lib as L


def shard_columns(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous column block [start, stop) of rank `rank` among `world` ranks;
    sizes differ by at most one, earlier ranks get the extra column."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batches(ranges: List[Tuple[int, int]], rank: int, world: int) -> List[Tuple[int, int]]:
    """Split every global batch (a contiguous column range of
    src/optimize_whitening.jl:31-32) `world` ways, so batch membership -- hence
    negll_history -- does not depend on the number of GPUs."""
    out = []
    for (s, e) in ranges:
        a, b = shard_columns(e - s, rank, world)
        out.append((s + a, s + b))
    return out


def local_batch_counts(n_global: int, nbatches: int, rank: int, world: int) -> List[int]:
    """Number of columns of every global batch (src/optimize_whitening.jl:31-32 on the GLOBAL sample count) that
    fall to `rank`: what `optimize_whitening(..., group=True, batch_counts=...)` needs when the local sample counts
    of the ranks differ (deriving the batches from the local count would give the ranks different numbers of steps)."""
    from .whitening import batch_ranges
    return [b - a for a, b in shard_batches(batch_ranges(n_global, nbatches), rank, world)]


def allreduce_sums(sums: np.ndarray, n_local: int):
    """Host-side all-reduce of raw float64 sums plus the local sample count over
    the default torch.distributed group (gloo or nccl).  Returns (sums, N_global).
    The GPU path does the same thing with ncclAllReduce inside the library."""
    import torch
    import torch.distributed as dist
    buf = torch.from_numpy(np.concatenate([np.asarray(sums, dtype=np.float64).ravel(), [float(n_local)]]))
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            buf = buf.cuda()
            dist.all_reduce(buf)
            buf = buf.cpu()
        else:
            dist.all_reduce(buf)
    out = buf.numpy()
    return out[:-1].reshape(np.shape(sums)), int(round(out[-1]))


def init_group(ctx) -> Tuple[int, int]:
    """Create the library's NCCL communicator for `ctx` from an initialised
    torch.distributed process group.  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = ctx._lib
    ident = (C.c_ubyte * L.ENF_UNIQUE_ID_BYTES)()
    if rank == 0:
        L.check(lib.enf_group_unique_id(ident))
    t = torch.tensor(list(bytes(ident)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    ident = (C.c_ubyte * L.ENF_UNIQUE_ID_BYTES).from_buffer_copy(raw)
    L.check(lib.enf_group_init(ctx.handle, world, rank, ident), ctx.handle)
    return rank, world
