"""N>1 GPU path: sharded loss+gradient step with the library's ncclAllReduce
(enf_negll_grad_group) equals the single-GPU result on the whole batch.  Needs
>= 2 GPUs (skipped otherwise); launched as two processes from inside the test."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import enf_b200 as E
    from chains import both, flat_grads
    from oracle import enf_oracle as O
    ctx = E.Context(rank)
    # no torch.distributed anywhere in this process: the NCCL unique id travels over a TCP socket (MASTER_PORT + 1)
    assert E.dist.init_group(ctx, rank=rank, world=world) == (rank, world)
    assert "torch" not in sys.modules
    D, N = 32, 10007
    fo, fe = both(["cc", "jo", "hh4", "ss"], D, 3, np.float64)
    X = np.random.default_rng(4).standard_normal((D, N))
    a, b = E.dist.shard_columns(N, rank, world)
    v, g = E.mvnormal_negll_trafograd(fe, E.B200Matrix.from_host(X[:, a:b], ctx), group=True)
    v_ref, g_ref = O.mvnormal_negll_trafograd(fo, X)
    assert abs(v - v_ref) < 1e-12 * (abs(v_ref) + 1), (v, v_ref)
    for (k, x), (_, y) in zip(flat_grads(g, fe), flat_grads(g_ref, fo)):
        y = y.reshape(x.shape)
        assert np.max(np.abs(x - y) / (np.abs(y) + np.sqrt(np.mean(y * y)) + 1e-30)) < 1e-10, k
    # the device-side fit loop over a sharded data set: every rank ends with the parameters of the single-process fit
    from enf_b200.dist import shard_batches
    Xs = np.random.default_rng(5).standard_normal((D, 4000)) * 1.3
    ranges = E.batch_ranges(4000, 5)
    mine = shard_batches(ranges, rank, world)
    Xl = np.concatenate([Xs[:, a:b] for a, b in mine], axis=1)          # this rank's columns of every batch, in batch order
    assert len({b - a for a, b in mine}) == 1                            # equal local batches -> local partition == global batches
    r = E.optimize_whitening(E.B200Matrix.from_host(Xl, ctx), fe, E.ADAGrad(), nbatches=5, nepochs=2, group=True, device_loop=True)
    r_ref = O.optimize_whitening(Xs, fo, O.ADAGrad(), nbatches=5, nepochs=2)
    h, h_ref = np.array(r["negll_history"]), np.array(r_ref["negll_history"])
    assert h.shape == h_ref.shape and np.max(np.abs(h - h_ref) / (np.abs(h_ref) + 1)) < 1e-9, (h, h_ref)
    # second-moment chain (Householder + ScaleShift at D = 128): the all-reduced quantity is [[S, m], [m^T, N]]
    D2, N2 = 128, 6001
    fo2, fe2 = both(["hh8", "ss"], D2, 7, np.float32)
    X2 = (np.random.default_rng(8).standard_normal((D2, N2)) * 1.2).astype(np.float32)
    cut = (N2 // 2) // 4 * 4                                             # 16-byte aligned shard boundary (TMA)
    lo = 0 if rank == 0 else cut
    hi = N2 if rank == world - 1 else cut
    v2, g2 = E.mvnormal_negll_trafograd(fe2, E.B200Matrix.from_host(X2[:, lo:hi], ctx), group=True)
    v2_ref, g2_ref = O.mvnormal_negll_trafograd(fo2, X2.astype(np.float64))
    assert abs(v2 - v2_ref) < 1e-5 * (abs(v2_ref) + 1), (v2, v2_ref)
    for (k, x), (_, y) in zip(flat_grads(g2, fe2), flat_grads(g2_ref, fo2)):
        y = y.reshape(x.shape)
        assert np.max(np.abs(x - y) / (np.abs(y) + np.sqrt(np.mean(y * y)) + 1e-30)) < 4e-5, k
    # ... and the device-side fit loop on it: per-batch moment matrices all-reduced once, then identical steps on every rank
    Xf = (np.random.default_rng(9).standard_normal((D2, 4800)) * 1.4).astype(np.float32)
    ranges = E.batch_ranges(4800, 6)
    mine = shard_batches(ranges, rank, world)
    Xfl = np.concatenate([Xf[:, a:b] for a, b in mine], axis=1)
    rf = E.optimize_whitening(E.B200Matrix.from_host(Xfl, ctx), fe2, E.ADAGrad(), nbatches=6, nepochs=2, group=True, device_loop=True)
    rf_ref = O.optimize_whitening(Xf.astype(np.float64), fo2, O.ADAGrad(), nbatches=6, nepochs=2)
    hf, hf_ref = np.array(rf["negll_history"]), np.array(rf_ref["negll_history"])
    assert hf.shape == hf_ref.shape and np.max(np.abs(hf[:6] - hf_ref[:6]) / (np.abs(hf_ref[:6]) + 1)) < 1e-5, (hf, hf_ref)
    assert np.max(np.abs(hf - hf_ref) / (np.abs(hf_ref) + 1)) < 2e-3, (hf, hf_ref)
    # ranks with DIFFERENT local sample counts: batches must come from the global partition (dist.local_batch_counts), a
    # rank may hold no column of a batch; deriving the batches from the local count is refused on every rank
    from enf_b200.dist import local_batch_counts
    Ng = 1001
    Xu = np.random.default_rng(11).standard_normal((D, Ng)) * 1.1
    cnt = local_batch_counts(Ng, 4, rank, world)
    assert len(cnt) == 5 and sum(local_batch_counts(Ng, 4, r, world)[4] for r in range(world)) == 1
    mine = shard_batches(E.batch_ranges(Ng, 4), rank, world)
    Xul = np.concatenate([Xu[:, a:b] for a, b in mine], axis=1)
    Xud = E.B200Matrix.from_host(Xul, ctx)
    ru_ref = O.optimize_whitening(Xu, fo, O.ADAGrad(), nbatches=4, nepochs=2)
    for dev in (True, False):
        ru = E.optimize_whitening(Xud, fe, E.ADAGrad(), nepochs=2, group=True, device_loop=dev, batch_counts=cnt)
        hu, hu_ref = np.array(ru["negll_history"]), np.array(ru_ref["negll_history"])
        assert hu.shape == hu_ref.shape == (10,) and np.max(np.abs(hu - hu_ref) / (np.abs(hu_ref) + 1)) < 1e-9, (dev, hu, hu_ref)
    Xmis = E.B200Matrix.from_host(Xu[:, :502] if rank == 0 else Xu[:, :498], ctx)      # 4 local batches of 167/... vs 3 of 166
    try:
        E.optimize_whitening(Xmis, fe, E.ADAGrad(), nbatches=3, nepochs=1, group=True, device_loop=True)
        raise SystemExit("mismatching batch counts were not detected")
    except E.EnfError as exc:
        assert "disagree" in str(exc), exc
    ctx.sync()
    print("rank", rank, "ok")
""") % (ROOT, ROOT)


def test_sharded_gradient_step_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
