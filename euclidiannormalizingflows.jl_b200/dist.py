"""Multi-GPU host logic: one process per GPU, samples sharded by columns.

Forward / inverse / ladj need no communication (columns are independent,
parameters are replicated).  The loss+gradient step has exactly one exchange:
an all-reduce (sum, float64) of the raw loss and parameter-gradient sums
(SURVEY §8e).  On GPUs that is one kernel over NVLink peer memory (or an
ncclAllReduce) issued by the library on its own stream (enf_negll_grad_group).
The only thing the ranks must exchange on the host is the 128-byte NCCL unique
id: `init_group` hands it out over a plain TCP socket (MASTER_ADDR / MASTER_PORT,
no PyTorch involved), or over an already initialised torch.distributed group if
the caller has one.
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


def exchange_bytes(payload: bytes, rank: int, world: int, addr: str = "127.0.0.1", port: int = 29400, timeout: float = 120.0) -> bytes:
    """Rank 0 sends `payload` to every other rank over TCP; every rank returns rank 0's payload.  A torch-free
    rendezvous for the NCCL unique id (the Julia shim would use Sockets the same way)."""
    import socket
    import struct
    import time
    if world <= 1:
        return payload
    if rank == 0:
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as srv:
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            srv.settimeout(timeout)
            seen = set()
            while len(seen) < world - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    (r,) = struct.unpack("!i", _recv_exact(conn, 4))
                    conn.sendall(struct.pack("!i", len(payload)) + payload)
                    seen.add(r)
        return payload
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=timeout) as c:
                c.sendall(struct.pack("!i", rank))
                (n,) = struct.unpack("!i", _recv_exact(c, 4))
                return _recv_exact(c, n)
        except (ConnectionRefusedError, ConnectionResetError, socket.timeout):
            if time.monotonic() > deadline:
                raise
            time.sleep(0.05)


def _recv_exact(sock, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        part = sock.recv(n - len(buf))
        if not part:
            raise ConnectionResetError("peer closed the rendezvous socket")
        buf += part
    return buf


def init_group(ctx, rank: int = None, world: int = None, addr: str = None, port: int = None) -> Tuple[int, int]:
    """Create the library's NCCL communicator (and the peer-memory exchange buffers) for `ctx`.  Returns (rank, world).

    With an initialised torch.distributed process group and no explicit rank/world, the unique id travels over that
    group.  Otherwise rank / world come from the arguments or RANK / WORLD_SIZE, and the id travels over a TCP socket
    on MASTER_ADDR : MASTER_PORT + 1 (`exchange_bytes`): no PyTorch on this path."""
    import os
    lib = ctx._lib
    ident = (C.c_ubyte * L.ENF_UNIQUE_ID_BYTES)()
    use_torch = False
    if rank is None and world is None:
        try:
            import torch.distributed as dist
            use_torch = dist.is_initialized()
        except ImportError:
            use_torch = False
    if not use_torch:
        if "ENF_NCCL_LIB" not in os.environ:
            # one libnccl.so.2 per process: prefer the copy of the nvidia-nccl wheel (what PyTorch would load) if it is
            # installed, so that importing torch later in this process still works
            try:
                import importlib.util
                spec = importlib.util.find_spec("nvidia.nccl")
                cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2") if spec else None
                if cand and os.path.exists(cand):
                    os.environ["ENF_NCCL_LIB"] = cand
            except (ImportError, ValueError, IndexError):
                pass
        rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else int(world)
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(port) if port is not None else int(os.environ.get("MASTER_PORT", 29399)) + 1
        if rank == 0:
            L.check(lib.enf_group_unique_id(ident))
        raw = exchange_bytes(bytes(ident), rank, world, addr, port)
        ident = (C.c_ubyte * L.ENF_UNIQUE_ID_BYTES).from_buffer_copy(raw)
        L.check(lib.enf_group_init(ctx.handle, world, rank, ident), ctx.handle)
        return rank, world
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
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
