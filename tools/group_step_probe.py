"""torchrun probe: where does the time of a sharded gradient step go?  (kernel | + exchange | whole group call)"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import enf_b200 as E
from enf_b200 import _lib as L
from chains import build
ctx = E.Context(rank)
if world > 1:
    E.dist.init_group(ctx)
D, nb = 32, 2_500_000
f = build(E, ["cc", "jo", "hh4", "ss"], D, np.random.default_rng(43), np.float32)
X = E.B200Matrix.randn(D, nb * 4, np.float32, ctx=ctx, col0=rank * nb * 4)
ch = E.get_chain(f, D, np.float32, ctx)
lib = ctx._lib
negll = C.c_double(); g = np.empty(ch.nparams, dtype=np.float32)
def bar():
    if world > 1: dist.barrier()
def wall(fn, R=40):
    for _ in range(5): fn()
    ctx.sync(); bar(); t0 = time.perf_counter()
    for _ in range(R): fn()
    ctx.sync(); dt = (time.perf_counter() - t0) / R * 1e6
    if world > 1:
        t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t)
    return dt
xp = lambda i: C.c_void_p(X.cols((i % 4) * nb, (i % 4 + 1) * nb).ptr)
it = [0]
def part(): it[0] += 1; L.check(lib.enf_negll_grad_partial(ch.handle, xp(it[0]), nb, None, None))
def part_ex(): part(); L.check(lib.enf_group_allreduce_sums(ch.handle, nb))
def part_sync(): part(); ctx.sync()
def part_ex_sync(): part_ex(); ctx.sync()
def full(): it[0] += 1; L.check((lib.enf_negll_grad_group if world > 1 else lib.enf_negll_grad)(ch.handle, xp(it[0]), nb, 1, C.byref(negll), g.ctypes.data_as(C.c_void_p)))
def py(): it[0] += 1; E.mvnormal_negll_trafograd(f, X.cols((it[0] % 4) * nb, (it[0] % 4 + 1) * nb), group=world > 1)
res = {"partial (async, stream-bound)": wall(part), "partial + sync per step": wall(part_sync)}
if world > 1:
    res["partial + exchange (async)"] = wall(part_ex); res["partial + exchange + sync per step"] = wall(part_ex_sync)
res["full C call"] = wall(full); res["python wrapper"] = wall(py)
if rank == 0:
    for k, v in res.items(): print(f"{k:40s} {v:8.1f} us", flush=True)
if world > 1: dist.destroy_process_group()
