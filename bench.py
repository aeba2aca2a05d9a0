#!/usr/bin/env python
"""bench.py -- trafo-chain samples/s on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference

One "step" is one fused forward+ladj pass of the C3 chain
(CenterStretch ∘ JohnsonTrafo ∘ HouseholderTrafo(16x4), Float32) over this rank's
shard of synthetic N(0,1) samples that already sit in HBM.  Weak scaling:
1.25e8 samples per GPU (8 GPUs = the 1e9 samples of BASELINE.json configs[2]);
no collective is on this path (columns are independent).

The JSON line also carries
  roofline     achieved HBM GB/s of the fused kernel (algorithmic (2D+1)*4 B/sample)
  e2e          same metric through the public host-matrix API (pinned host
               buffers, H2D + kernel + D2H inside the timed region)
  cpu_baseline the C restatement of the reference's unfused CPU algorithm on a
               bounded sample of the same workload (rank 0, N=1 only)
  parity_at_scale  GPU output of the first 4e6 and the last 1e4 columns of the resident buffer against
               the CPU restatement (N=1)
  secondary    the C5 optimize_whitening gradient step (D=32): kernel / exchange / host time apart, fraction of
               the HBM and instruction-issue bounds, and (N>1) parity of the sharded step with the un-sharded one
  extras       inverse+ladj, forward on skewed samples, C4 on the tensor cores, C4/C2 fits, C1
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = ("C3 chain CenterStretch∘JohnsonTrafo∘HouseholderTrafo(16x4), D=16, Float32, forward+ladj, "
            "%d samples per GPU, synthetic N(0,1) (Philox4x32-10 keyed by global element index)")
D_MAIN, K_HH = 16, 4
N_PER_GPU = 125_000_000          # C3: 1e9 samples over 8 GPUs
N_E2E = 1 << 25                  # samples per e2e step through host buffers (2 GiB in, 2.1 GiB out)
D_GRAD, N_GRAD_BATCH = 32, 2_500_000   # C5: 2.5e8 samples/GPU, nbatches=100
SEED = 42


def tensor_peak_tf32():
    """Dense TF32 tensor peak in TFLOP/s: half of the measured cuBLAS bf16 figure (burst), else half of the nominal 2250."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["bf16_tflops"]) / 2, "measured bf16 / 2 (MEASURED_PEAKS.json)"
    return 1125.0, "nominal bf16 / 2"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


def c3_chain(ns, dtype=np.float32):
    from chains import build
    return build(ns, ["hh%d" % K_HH, "jo", "cs"], D_MAIN, np.random.default_rng(SEED), dtype)


def c5_chain(ns, dtype=np.float32):
    from chains import build
    return build(ns, ["cc", "jo", "hh4", "ss"], D_GRAD, np.random.default_rng(SEED + 1), dtype)


def rel_err(got, ref):
    """max |got - ref| / (|ref| + RMS(ref)): the parity metric of tests/conftest.py."""
    got = np.asarray(got, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    scale = float(np.sqrt(np.mean(ref * ref))) or 1.0
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + scale)))


def chg_n_raw(ch, ctx):
    """number of raw float64 sums of a chain's gradient step (enf_chain_describe)"""
    return int(ch.describe().split("n_raw=")[1].split()[0])


def grads_err(g, g_ref, f):
    from chains import flat_grads
    return max(rel_err(a, b.reshape(a.shape)) for (_, a), (_, b) in zip(flat_grads(g, f), flat_grads(g_ref, f)))


# ------------------------------------------------------------------ CPU restatement
def cpu_reference_rate(n_samples, threads, x_host=None, dtype=np.float32):
    """samples/s of the unfused C restatement (oracle/libenf_ref_cpu.so) for the
    C3 forward+ladj pass on `n_samples` samples.  dtype=float64: the same restatement in Float64 on the
    Float32-rounded parameters and inputs (the exact value both Float32 evaluations approximate)."""
    from oracle import enf_oracle as O
    so = os.path.join(ROOT, "oracle", "libenf_ref_cpu.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    f = c3_chain(O)
    hh, jo, cs = O.flatten(f)
    ps = [np.asfortranarray(hh.V).ravel(order="F"), np.concatenate([jo.gamma, jo.delta, jo.xi, jo.lam]),
          np.concatenate([cs.a, cs.b, cs.c])]
    ps = [np.ascontiguousarray(np.asarray(p, dtype=np.float32), dtype=dtype) for p in ps]
    FP = C.POINTER(C.c_float if dtype == np.float32 else C.c_double)
    parr = (FP * 3)(*[p.ctypes.data_as(FP) for p in ps])
    kinds, Ks = (C.c_int * 3)(5, 2, 0), (C.c_int * 3)(K_HH, 0, 0)
    if x_host is None:
        x_host = np.random.default_rng(SEED).standard_normal((n_samples, D_MAIN), dtype=np.float32)  # memory == D x N col-major
    x_host = np.ascontiguousarray(x_host, dtype=dtype)
    y = np.empty_like(x_host)
    l = np.empty(n_samples, dtype=dtype)
    lib.ref_set_threads(int(threads))
    fn = lib.ref_forward_ladj_f32 if dtype == np.float32 else lib.ref_forward_ladj_f64
    t = time.perf_counter()
    rc = fn(D_MAIN, C.c_int64(n_samples), 3, kinds, Ks, parr, x_host.ctypes.data_as(FP),
            y.ctypes.data_as(FP), l.ctypes.data_as(FP))
    dt = time.perf_counter() - t
    assert rc == 0
    return n_samples / dt, dt, (y, l)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 4_000_000
    cpu_reference_rate(200_000, threads)                      # page in / warm
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_rate(n, threads)
    rates, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        r, _, _ = cpu_reference_rate(n, threads)
        rates.append(r)
    total = time.perf_counter() - t0
    value = n * args.steps / sum(n / r for r in rates)
    line = {
        "impl": "reference", "metric": "trafo-chain fwd+ladj samples/s", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(n / r for r in rates) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.samples_per_gpu, "samples_per_gpu": args.samples_per_gpu, "D": D_MAIN,
                   "householder_K": K_HH, "sample_per_step": n,
                   "note": "CPU arm: every step is a bounded sample of the workload (sample_per_step i.i.d. N(0,1) samples, numpy generator)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{n} samples/step of the same synthetic N(0,1) input; C restatement of the "
                                   "reference's unfused CPU algorithm (Julia is not installed in this image)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": total,
    }
    emit(line)


# ------------------------------------------------------------------ our arm
# The contract is ONE JSON line on stdout.  Native libraries (NCCL's version banner, ...) also write to fd 1, so
# everything but the result line is sent to stderr: fd 1 is pointed at fd 2 and the line goes to the saved fd.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import enf_b200 as E
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = E.Context(local_rank)
    hbm_peak, peak_src = peaks()
    Nl = args.samples_per_gpu
    fe = c3_chain(E)
    X = E.B200Matrix.randn(D_MAIN, Nl, np.float32, seed=SEED, col0=rank * Nl, ctx=ctx)
    Y = X.empty_like()
    Ld = E.B200Matrix(ctx, 1, Nl, np.float32)
    ctx.sync()

    def step():
        E.with_logabsdet_jacobian(fe, X, out=(Y, Ld))

    for _ in range(args.warmup):
        step()
    ctx.sync()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = ctx.launches
    t_wall0 = time.time()
    ctx.record(0)
    for _ in range(args.steps):
        step()
    ctx.record(1)
    ms = ctx.elapsed_ms(0, 1)
    ctx.sync()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = ctx.launches - l0
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = Nl * world / (ms_per_step * 1e-3)
    bytes_per_sample = (2 * D_MAIN + 1) * 4
    my_ms = ctx.elapsed_ms(0, 1) / args.steps                # this rank's kernel time (one launch per step)
    achieved = bytes_per_sample * Nl / (my_ms * 1e-3) / 1e9
    traffic = None                                           # DRAM bytes per launch, from the committed ncu capture
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = float(json.load(f)["dram_bytes_per_sample"]) * Nl

    # ---- parity at benchmark scale: the first 4e6 and the LAST 1e4 columns of the resident buffer (the last ones sit
    # beyond 2^31 elements / 2^33 bytes) against the CPU restatement of the reference's Float32 path on the same inputs
    parity_at_scale, cpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_cpu, n_tail = min(4_000_000, Nl), min(10_000, Nl)
        sample = np.ascontiguousarray(X.cols(0, n_cpu).to_host().T)      # (N, D) C-order == D x N column-major
        cpu_reference_rate(100_000, threads)
        r, dt, (yc, lc) = cpu_reference_rate(n_cpu, threads, sample)
        tail = np.ascontiguousarray(X.cols(Nl - n_tail, Nl).to_host().T)
        _, _, (yt, lt) = cpu_reference_rate(n_tail, threads, tail)
        r1, dt1, _ = cpu_reference_rate(min(1_000_000, n_cpu), 1, sample[:min(1_000_000, n_cpu)])
        _, _, (yc64, lc64) = cpu_reference_rate(n_cpu, threads, sample, np.float64)
        _, _, (yt64, lt64) = cpu_reference_rate(n_tail, threads, tail, np.float64)
        yg, lg = Y.cols(0, n_cpu).to_host().T, Ld.cols(0, n_cpu).to_host()[0]
        ygt, lgt = Y.cols(Nl - n_tail, Nl).to_host().T, Ld.cols(Nl - n_tail, Nl).to_host()[0]
        fin = np.isfinite(yc).all(1) & np.isfinite(lc)                   # the literal Float32 formula overflows on rare tails
        fint = np.isfinite(yt).all(1) & np.isfinite(lt)

        def three_way(g_y, g_l, r_y, r_l, t_y, t_l, ok):
            return {"samples": int(len(g_l)),
                    "cuda_vs_f64": {"y": rel_err(g_y, t_y), "ladj": rel_err(g_l, t_l)},
                    "reference_f32_vs_f64": {"y": rel_err(r_y[ok], t_y[ok]), "ladj": rel_err(r_l[ok], t_l[ok])},
                    "cuda_vs_reference_f32": {"y": rel_err(g_y[ok], r_y[ok]), "ladj": rel_err(g_l[ok], r_l[ok])}}

        parity_at_scale = {
            "first": three_way(yg, lg, yc, lc, yc64, lc64, fin),
            "last": dict(three_way(ygt, lgt, yt, lt, yt64, lt64, fint), first_column=Nl - n_tail),
            "gpu_nonfinite": int((~np.isfinite(yg)).sum() + (~np.isfinite(lg)).sum() + (~np.isfinite(ygt)).sum() + (~np.isfinite(lgt)).sum()),
            "reference_f32_nonfinite_columns": int((~fin).sum() + (~fint).sum()), "tol": 1e-5,
            "against": "C restatement of the reference's CPU path (oracle/libenf_ref_cpu.so) in Float32 (what the reference computes "
                       "for Float32 input) and in Float64 on the same Float32 inputs (the exact value); metric of tests/conftest.py. "
                       "The worst elements of 6.4e7 are inputs of CenterStretch within 0.01 of its centre in a row with e^{ba} = 67, "
                       "where dy/dx = 1/S = 34 amplifies the Float32 rounding of the JohnsonTrafo output it receives: no Float32 "
                       "evaluation of this chain is closer to the exact value than that (the reference's own Float32 distance is "
                       "listed beside the CUDA one).  ok = no non-finite output and the CUDA distance from the exact value is within "
                       "max(tol, 1.25 x the Float32 reference's own distance)"}
        worst = lambda k: max(parity_at_scale[s_][k][q] for s_ in ("first", "last") for q in ("y", "ladj"))
        parity_at_scale["within_tol_of_exact"] = bool(worst("cuda_vs_f64") <= 1e-5)
        parity_at_scale["ok"] = bool(parity_at_scale["gpu_nonfinite"] == 0 and all(
            parity_at_scale[s_]["cuda_vs_f64"][q] <= max(1e-5, 1.25 * parity_at_scale[s_]["reference_f32_vs_f64"][q])
            for s_ in ("first", "last") for q in ("y", "ladj")))
        cpu = {"value": r, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"first {n_cpu} samples of the same synthetic input ({dt:.1f} s); C restatement of the "
                         "reference's unfused CPU algorithm (Julia is not installed in this image)",
               "one_thread": {"value": r1, "unit": "samples/s", "cores": 1,
                              "sample": f"first {min(1_000_000, n_cpu)} samples ({dt1:.1f} s); the reference itself is single-threaded "
                                        "apart from BLAS in src/householder_trafo.jl:4"}}
        del sample, yc, lc, yg, lg, yc64, lc64

    # ---- e2e: public host-matrix API, pinned host buffers, H2D + kernel + D2H timed
    n_e2e = min(N_E2E, Nl)
    xh = ctx.pinned_empty((D_MAIN, n_e2e), np.float32)
    yh = ctx.pinned_empty((D_MAIN, n_e2e), np.float32)
    lh = ctx.pinned_empty((1, n_e2e), np.float32)
    xh[...] = X.cols(0, n_e2e).to_host()
    E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)  # warm (allocates the staging slots)
    e2e_steps = max(3, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_e2e * world * e2e_steps / e2e_s
    # the same pipeline without the kernel (ENF_HOST_COPY_ONLY): what the platform's host<->device path carries with all
    # ranks copying at once -- the ceiling of the end-to-end number
    os.environ["ENF_HOST_COPY_ONLY"] = "1"
    try:
        E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)
        copy_s = max_over_ranks(time.perf_counter() - t0)
    finally:
        del os.environ["ENF_HOST_COPY_ONLY"]
    E.with_logabsdet_jacobian(fe, xh, out=(yh, lh), ctx=ctx)      # leave real results in the host buffers

    # ---- secondary metric (the "fwd+grad" half of BASELINE.json's metric, and the only path with a collective): the C5
    # optimize_whitening gradient step -- fused loss+gradient kernel on this rank's shard of the batch, all-reduce of the
    # raw float64 sums over the group, finish -- with the device time of the kernel and of the exchange measured apart,
    # and (N > 1) a parity check of the sharded step against the un-sharded step on the regenerated global batch
    secondary = None
    try:
        grp = world > 1
        ge = c5_chain(E)
        nb = min(N_GRAD_BATCH, Nl * D_MAIN // D_GRAD // 8)
        Xg = E.B200Matrix(ctx, D_GRAD, nb * 8, np.float32, _ptr=X.ptr, _owner=X)   # the resident samples as D=32 columns
        if grp:
            E.dist.init_group(ctx)
        chg = E.get_chain(ge, D_GRAD, np.float32, ctx)
        lib = ctx._lib
        part = lambda i: E._lib.check(lib.enf_negll_grad_partial(chg.handle, C.c_void_p(Xg.cols(i * nb, (i + 1) * nb).ptr), nb, None, None), ctx.handle)
        nsteps = 40
        for i in range(5):
            E.mvnormal_negll_trafograd(ge, Xg.cols(i * nb, (i + 1) * nb), group=grp)
        ctx.sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(nsteps):
            v_last, g_last = E.mvnormal_negll_trafograd(ge, Xg.cols((i % 8) * nb, (i % 8 + 1) * nb), group=grp)
        g_s = max_over_ranks(time.perf_counter() - t0) / nsteps
        ctx.sync()
        ctx.record(10)
        for i in range(nsteps):
            part(i % 8)
        ctx.record(11)
        k_ms = max_over_ranks(ctx.elapsed_ms(10, 11) / nsteps)
        x_us = None
        if grp:
            barrier()
            ctx.record(12)
            for _ in range(50):
                E._lib.check(lib.enf_group_allreduce_sums(chg.handle, nb), ctx.handle)
            ctx.record(13)
            x_us = max_over_ranks(ctx.elapsed_ms(12, 13) / 50) * 1e3
        prof = {}
        pp = os.path.join(ROOT, "profiles", "r2_grad_c5.json")
        if os.path.exists(pp):
            with open(pp) as f:
                prof = json.load(f)
        elems_per_s = nb * D_GRAD / (k_ms * 1e-3)                         # per GPU, kernel only
        sm_hz = 1e6 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"]) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1.965e9
        issue_peak = 148 * 4 * sm_hz * 32                                 # thread instructions per second at IPC 1 per scheduler
        secondary = {
            "metric": "optimize_whitening gradient-step samples/s (C5 chain CenterContract,JohnsonTrafo,Householder(32x4),ScaleShift; D=32, Float32)",
            "value": nb * world / g_s, "unit": "samples/s", "ms_per_step": g_s * 1e3, "batch_per_gpu": nb, "n_gpus": world,
            "scaling": "weak", "step": "enf_negll_grad" + ("_group: kernel + peer-memory/NCCL all-reduce of %d float64 sums + finish" % (chg_n_raw(chg, ctx) + 1) if grp else ": kernel + finish"),
            "kernel_ms": k_ms, "kernel_samples_per_s_per_gpu": nb / (k_ms * 1e-3), "exchange_us": x_us,
            "host_side_us": g_s * 1e6 - k_ms * 1e3 - (x_us or 0.0),
            "limiter": "gradient kernel (%.0f %% of the step): bound by instruction issue / FP32 pipe, not HBM" % (100 * k_ms / (g_s * 1e3)),
            "roofline": {"hbm_frac": D_GRAD * 4 * nb / (k_ms * 1e-3) / 1e9 / hbm_peak,
                         "thread_instr_per_element": prof.get("thread_instr_per_element"),
                         "issue_frac": (elems_per_s * prof["thread_instr_per_element"] / issue_peak) if prof.get("thread_instr_per_element") else None,
                         "fma_pipe_frac": (elems_per_s * prof["fma_pipe_cycles_per_element"] / issue_peak) if prof.get("fma_pipe_cycles_per_element") else None,
                         "source": "profiles/r2_grad_c5.json (ncu --set full of the same kernel)"},
        }
        # the same steps with the loop on the device (enf_optimize_whitening: gradient kernel + one kernel that reduces the
        # partials, exchanges the sums over NVLink peer memory, finishes and applies the ADAGrad update): no host round
        # trip per step, so no rank skew from the host side either
        try:
            Xdl = Xg.cols(0, 8 * nb)
            E.optimize_whitening(Xdl, ge, E.ADAGrad(), nbatches=8, nepochs=1, device_loop=True, group=grp)
            ctx.sync()
            barrier()
            t0 = time.perf_counter()
            rdl = E.optimize_whitening(Xdl, ge, E.ADAGrad(), nbatches=8, nepochs=5, device_loop=True, group=grp)
            dl_s = max_over_ranks(time.perf_counter() - t0) / 40
            secondary["device_loop"] = {"ms_per_step": dl_s * 1e3, "value": nb * world / dl_s, "unit": "samples/s", "steps": 40,
                                        "negll_first_last": [rdl["negll_history"][0], rdl["negll_history"][-1]],
                                        "step": "optimize_whitening(device_loop=True): 2 launches per step, exchange fused into the update kernel"}
        except Exception as exc:  # noqa: BLE001
            secondary["device_loop"] = {"error": f"{type(exc).__name__}: {exc}"}
        if grp:
            err = None
            if rank == 0:
                i = (nsteps - 1) % 8
                Xall = E.B200Matrix(ctx, D_GRAD, nb * world, np.float32)
                for r in range(world):                                   # rank r's columns of batch i, regenerated from the global element index
                    col0 = (r * Nl * D_MAIN) // D_GRAD + i * nb
                    E._lib.check(lib.enf_fill_normal(ctx.handle, 0, C.c_void_p(Xall.cols(r * nb, (r + 1) * nb).ptr), D_GRAD, nb, col0, SEED), ctx.handle)
                v_ref, g_ref = E.mvnormal_negll_trafograd(ge, Xall)      # un-sharded step on one GPU
                err = max(abs(v_last - v_ref) / (abs(v_ref) + 1), grads_err(g_last, g_ref, ge))
                del Xall
            secondary["group_parity_err"] = err
            secondary["group_parity"] = "sharded step on %d GPUs vs enf_negll_grad on the regenerated global batch (rank 0), max over loss and gradient leaves" % world
        del Xg
    except Exception as exc:  # noqa: BLE001
        secondary = {"error": f"{type(exc).__name__}: {exc}"}
        print(f"[bench] secondary leg failed: {secondary['error']}", file=sys.stderr)

    extras = {}
    if not args.no_extras:
        # the secondary legs must never cost the headline line: a failure is recorded, not raised (the code path is
        # the same on every rank, so the ranks stay in step)
        try:
            # inverse + ladj of the same chain (F2: just another chain)
            fi = E.inverse(fe)
            for _ in range(2):
                E.with_logabsdet_jacobian(fi, Y, out=(X, Ld))
            ctx.record(2)
            for _ in range(3):
                E.with_logabsdet_jacobian(fi, Y, out=(X, Ld))
            ctx.record(3)
            inv_ms = max_over_ranks(ctx.elapsed_ms(2, 3) / 3)
            extras["inverse_ladj"] = {"samples_per_s": Nl * world / (inv_ms * 1e-3), "ms_per_pass": inv_ms,
                                      "hbm_frac": bytes_per_sample * Nl / (inv_ms * 1e-3) / 1e9 / hbm_peak}
            del fi
            # forward + ladj on SKEWED samples: Y = f_true(XW) is how the reference's examples make their skewed data
            # (examples/nf_example_1d.jl:8-15, nf_example_2d.jl:12-18); heavier tails exercise the range-guard path
            for _ in range(2):
                E.with_logabsdet_jacobian(fe, Y, out=(X, Ld))
            ctx.record(14)
            for _ in range(3):
                E.with_logabsdet_jacobian(fe, Y, out=(X, Ld))
            ctx.record(15)
            sk_ms = max_over_ranks(ctx.elapsed_ms(14, 15) / 3)
            chk = X.cols(0, min(Nl, 1_000_000)).to_host()
            extras["fwd_ladj_skewed"] = {"samples_per_s": Nl * world / (sk_ms * 1e-3), "ms_per_pass": sk_ms,
                                         "hbm_frac": bytes_per_sample * Nl / (sk_ms * 1e-3) / 1e9 / hbm_peak,
                                         "input": "f(XW), XW ~ N(0,1): max |x| = %.1f" % float(np.abs(Y.cols(0, min(Nl, 1_000_000)).to_host()).max()),
                                         "nonfinite_in_first_1e6": int((~np.isfinite(chk)).sum())}
            del chk
            # the Gaussian samples again (the legs below reuse the buffer; the sharded ones regenerate it by global index)
            E._lib.check(ctx._lib.enf_fill_normal(ctx.handle, 0, C.c_void_p(X.ptr), D_MAIN, Nl, rank * Nl, SEED), ctx.handle)

            # C4: D=256, 64 reflections + ScaleShift on the tensor cores (tcgen05 3xTF32 GEMM per tile, enf_affine.cu)
            from chains import build
            f4 = build(E, ["hh64", "ss"], 256, np.random.default_rng(SEED + 2), np.float32)
            n4 = min(10_000_000 // 8 if world > 1 else 6_000_000, Nl * D_MAIN // 256)
            X4 = E.B200Matrix(ctx, 256, n4, np.float32, _ptr=X.ptr, _owner=X)
            Y4 = E.B200Matrix(ctx, 256, n4, np.float32, _ptr=Y.ptr, _owner=Y)
            L4 = E.B200Matrix(ctx, 1, n4, np.float32, _ptr=Ld.ptr, _owner=Ld)
            for _ in range(2):
                E.with_logabsdet_jacobian(f4, X4, out=(Y4, L4))
            ctx.record(4)
            for _ in range(3):
                E.with_logabsdet_jacobian(f4, X4, out=(Y4, L4))
            ctx.record(5)
            c4_ms = max_over_ranks(ctx.elapsed_ms(4, 5) / 3)
            # issued TF32 flops per sample, compact WY (enf_wy.cu): T = x W  3 x 2.256.64, V += T U'  3 x 2.64.256
            flops = 3 * 2 * 256 * 64 * 2
            os.environ["ENF_NO_WY"] = "1"                  # the dense fold y = W x + c (enf_affine.cu) on the same chain, for comparison
            try:
                E.with_logabsdet_jacobian(f4, X4, out=(Y4, L4))
                ctx.record(6)
                for _ in range(3):
                    E.with_logabsdet_jacobian(f4, X4, out=(Y4, L4))
                ctx.record(7)
                dense_ms = max_over_ranks(ctx.elapsed_ms(6, 7) / 3)
            finally:
                del os.environ["ENF_NO_WY"]
            extras["c4_d256_k64_tensor"] = {"samples_per_s": n4 * world / (c4_ms * 1e-3), "ms_per_pass": c4_ms, "samples_per_gpu": n4,
                                            "hbm_frac": (2 * 256 + 1) * 4 * n4 / (c4_ms * 1e-3) / 1e9 / hbm_peak,
                                            "tf32_tflops_issued": flops * n4 / (c4_ms * 1e-3) / 1e12,
                                            "tensor_frac": flops * n4 / (c4_ms * 1e-3) / 1e12 / tensor_peak_tf32()[0],
                                            "tensor_peak": {"tflops": tensor_peak_tf32()[0], "source": tensor_peak_tf32()[1]},
                                            "path": "compact WY, two chained tcgen05.mma kind::tf32 GEMMs per 128-sample tile (3xTF32), "
                                                    "y = alpha.(x + U'(W^T x)) + c with the tile held in tensor memory",
                                            "dense_fold_ms_per_pass": dense_ms,
                                            "dense_fold_hbm_frac": (2 * 256 + 1) * 4 * n4 / (dense_ms * 1e-3) / 1e9 / hbm_peak}

            # C4 gradient: loss + dV, da, db of the same chain from tensor-core second moments
            # (enf_moments.cu: S = X X^T with the samples as the contraction dimension) + cluster chain-rule kernel
            ch4 = E.get_chain(f4, 256, np.float32, ctx)
            part = lambda Xm: E._lib.check(ctx._lib.enf_negll_grad_partial(ch4.handle, C.c_void_p(Xm.ptr), Xm.N, None, None), ctx.handle)
            for _ in range(2):
                part(X4)
            ctx.record(8)
            for _ in range(3):
                part(X4)
            ctx.record(9)
            m4_ms = max_over_ranks(ctx.elapsed_ms(8, 9) / 3)
            nb4 = 100_000                                   # C4: N = 1e7, nbatches = 100
            for i in range(3):
                E.mvnormal_negll_trafograd(f4, X4.cols(i * nb4, (i + 1) * nb4), group=world > 1)
            barrier()
            t0 = time.perf_counter()
            for i in range(8):
                E.mvnormal_negll_trafograd(f4, X4.cols(i * nb4, (i + 1) * nb4), group=world > 1)
            g4_s = max_over_ranks(time.perf_counter() - t0) / 8
            extras["grad_c4_d256_k64_moments"] = {
                "moments_samples_per_s": n4 * world / (m4_ms * 1e-3), "moments_ms_per_pass": m4_ms, "samples_per_gpu": n4,
                "hbm_frac": 256 * 4 * n4 / (m4_ms * 1e-3) / 1e9 / hbm_peak,
                "tf32_tflops_issued": 2 * 2 * 256 * 256 * n4 / (m4_ms * 1e-3) / 1e12,
                "tensor_frac": 2 * 2 * 256 * 256 * n4 / (m4_ms * 1e-3) / 1e12 / tensor_peak_tf32()[0],
                "ms_per_step_batch_1e5": g4_s * 1e3, "step_samples_per_s": nb4 * world / g4_s,
                "path": "tcgen05.mma kind::tf32, MN-major operands (TMA 128B/32B-atom swizzle), P = Xh Xh^T + Xh (2Xl)^T, "
                        "float64 chain rule on column-sliced CTAs; step = set_params + moments + chain rule + D2H"
                        + (" + ncclAllReduce of the moments" if world > 1 else "")}

            # the same chain through the whole optimize_whitening loop on the device: one pass for the per-batch moment
            # matrices, then 2 launches per step whose cost does not depend on the number of samples
            nbf = max(1, min(60, n4 // nb4))                 # batches of 1e5 samples per GPU that fit into the C4 buffer
            nfit = nbf * nb4
            fit = lambda ne: E.optimize_whitening(X4.cols(0, nfit), f4, E.ADAGrad(), nbatches=nbf, nepochs=ne, device_loop=True, group=world > 1)
            fit(1)                                          # warm-up: allocations, kernel attributes
            barrier()
            t0 = time.perf_counter()
            fit(1)
            fit1_s = max_over_ranks(time.perf_counter() - t0)
            barrier()
            t0 = time.perf_counter()
            r4 = fit(5)
            fit5_s = max_over_ranks(time.perf_counter() - t0)
            extras["fit_c4_d256_k64_device_loop"] = {
                "samples_per_gpu": nfit, "nbatches": nbf, "total_ms_1_epoch": fit1_s * 1e3, "total_ms_5_epochs": fit5_s * 1e3,
                "us_per_step_after_first_epoch": (fit5_s - fit1_s) / (4 * nbf) * 1e6,
                "negll_first_last": [float(r4["negll_history"][0]), float(r4["negll_history"][-1])]}

            # C2: 1-D JohnsonTrafo + ScaleShiftTrafo whitening fit, 1e7 samples, nbatches=100 (examples/nf_example_1d.jl shape):
            # time per gradient step of the host loop (one fused kernel + host optimizer per step) and of the device loop
            n2 = min(10_000_000, Nl * D_MAIN)
            one = np.ones(1, dtype=np.float32)
            f2 = E.compose(E.JohnsonTrafo(0 * one, 5 * one, 0 * one, 5 * one), E.ScaleShiftTrafo(one.copy(), 0 * one))
            X2 = E.B200Matrix(ctx, 1, n2, np.float32, _ptr=X.ptr, _owner=X)
            rr1 = E.optimize_whitening(X2, f2, E.ADAGrad(), nbatches=100, nepochs=1, device_loop=True, group=world > 1)
            fit_parity = None
            if world > 1 and rank == 0 and n2 % 100 == 0:
                # the same fit un-sharded on one GPU: global batch b = the ranks' batch-b columns, regenerated by global index
                bs = n2 // 100
                Xall2 = E.B200Matrix(ctx, 1, n2 * world, np.float32)
                for b in range(100):
                    for r in range(world):
                        E._lib.check(ctx._lib.enf_fill_normal(ctx.handle, 0, C.c_void_p(Xall2.cols((b * world + r) * bs, (b * world + r + 1) * bs).ptr),
                                                              1, bs, r * Nl * D_MAIN + b * bs, SEED), ctx.handle)
                rr_ref = E.optimize_whitening(Xall2, f2, E.ADAGrad(), nbatches=100, nepochs=1, device_loop=True)
                h, h_ref = np.array(rr1["negll_history"]), np.array(rr_ref["negll_history"])
                fit_parity = float(np.max(np.abs(h - h_ref) / (np.abs(h_ref) + 1))) if h.shape == h_ref.shape else float("inf")
                del Xall2
            ctx.sync(); barrier()
            t0 = time.perf_counter()
            rr = E.optimize_whitening(X2, f2, E.ADAGrad(), nbatches=100, nepochs=20, device_loop=True, group=world > 1)
            dev_s = max_over_ranks(time.perf_counter() - t0) / len(rr["negll_history"])
            t0 = time.perf_counter()
            rr = E.optimize_whitening(X2, f2, E.ADAGrad(), nbatches=100, nepochs=2, group=world > 1)
            host_s = max_over_ranks(time.perf_counter() - t0) / len(rr["negll_history"])
            # ... and in Float64 (C2 is quoted for both types)
            n2d = n2 // 2
            one64 = np.ones(1, dtype=np.float64)
            f2d = E.compose(E.JohnsonTrafo(0 * one64, 5 * one64, 0 * one64, 5 * one64), E.ScaleShiftTrafo(one64.copy(), 0 * one64))
            X2d = E.B200Matrix(ctx, 1, n2d, np.float64, _ptr=Y.ptr, _owner=Y)
            E._lib.check(ctx._lib.enf_fill_normal(ctx.handle, 1, C.c_void_p(X2d.ptr), 1, n2d, rank * n2d, SEED), ctx.handle)
            E.optimize_whitening(X2d, f2d, E.ADAGrad(), nbatches=50, nepochs=1, device_loop=True, group=world > 1)
            ctx.sync(); barrier()
            t0 = time.perf_counter()
            rrd = E.optimize_whitening(X2d, f2d, E.ADAGrad(), nbatches=50, nepochs=20, device_loop=True, group=world > 1)
            dev64_s = max_over_ranks(time.perf_counter() - t0) / len(rrd["negll_history"])
            extras["fit_c2_d1"] = {"batch_per_gpu": n2 // 100, "us_per_step_device_loop": dev_s * 1e6, "us_per_step_host_loop": host_s * 1e6,
                                   "us_per_step_device_loop_f64": dev64_s * 1e6,
                                   "samples_per_s_device_loop": (n2 // 100) * world / dev_s,
                                   "group_parity_err": fit_parity,
                                   "note": "optimize_whitening steps; device loop = enf_optimize_whitening (2 launches/step, CUDA graph per epoch)"}

            # C1: examples/nf_example_2d.jl chain (ScaleShift ∘ Householder([1,0.3]) ∘ CenterStretch), 1e5 Float64 samples:
            # 4 MB working set, L2-resident and launch-latency-bound -> report microseconds per pass
            f1 = E.compose(E.ScaleShiftTrafo(np.array([1.3, 0.4]), np.array([2.5, -1.2])), E.HouseholderTrafo(np.array([1.0, 0.3])),
                           E.CenterStretch(np.array([4.0, 4.1]), np.array([2.0, 2.1]), np.array([3.0, 3.1])))
            f1i = E.inverse(f1)
            X1 = E.B200Matrix.randn(2, 100_000, np.float64, seed=SEED, ctx=ctx)
            Y1, L1, Z1 = X1.empty_like(), E.B200Matrix(ctx, 1, 100_000, np.float64), X1.empty_like()
            for _ in range(5):
                E.with_logabsdet_jacobian(f1, X1, out=(Y1, L1)); E.with_logabsdet_jacobian(f1i, Y1, out=(Z1, L1))
            ctx.record(6)
            for _ in range(50):
                E.with_logabsdet_jacobian(f1, X1, out=(Y1, L1)); E.with_logabsdet_jacobian(f1i, Y1, out=(Z1, L1))
            ctx.record(7)
            c1_us = max_over_ranks(ctx.elapsed_ms(6, 7) / 50) * 1e3
            extras["c1_2d_f64"] = {"us_forward_plus_inverse_with_ladj": c1_us, "samples": 100_000,
                                   "samples_per_s": 2 * 100_000 * world / (c1_us * 1e-6), "note": "L2-resident, launch-latency-bound"}
        except Exception as exc:  # noqa: BLE001
            extras["error"] = f"{type(exc).__name__}: {exc}"
            print(f"[bench] extras leg failed: {extras['error']}", file=sys.stderr)

    if rank == 0:
        line = {
            "metric": "trafo-chain fwd+ladj samples/s", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % Nl,
                       "samples_per_gpu": Nl, "D": D_MAIN, "householder_K": K_HH, "parallelism": "columns sharded, no collective",
                       "l2": "inputs (%.1f GB per pass) >> 126 MB L2, no flush needed" % (bytes_per_sample * Nl / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "traffic_source": "ncu dram__bytes_read+write per sample (profiles/r1_traffic.json) x samples per launch",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": bytes_per_sample},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": n_e2e * D_MAIN * 4,
                    "d2h_bytes_per_step": n_e2e * (D_MAIN + 1) * 4, "samples_per_step": n_e2e, "steps": e2e_steps,
                    "api": "with_logabsdet_jacobian(chain, pinned host matrix) -> enf_forward_ladj_host",
                    "pcie_gbs_per_rank_each_way": [n_e2e * D_MAIN * 4 * e2e_steps / e2e_s / 1e9, n_e2e * (D_MAIN + 1) * 4 * e2e_steps / e2e_s / 1e9],
                    "host_buffers": "pinned, first-touched on the NUMA node of the rank's GPU when the topology is visible (enf_host_alloc)",
                    "copy_only_value": n_e2e * world * e2e_steps / copy_s,
                    "frac_of_copy_only": copy_s / e2e_s,
                    "copy_only": "the same chunked H2D / D2H pipeline with the kernel launch skipped (ENF_HOST_COPY_ONLY=1), all ranks "
                                 "at once: the host<->device ceiling of this box for this traffic pattern"},
            "gpu_launches": launches, "clocks": clocks, "extras": extras,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity_at_scale is not None:
            line["parity_at_scale"] = parity_at_scale
        if secondary is not None:
            line["secondary"] = secondary
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
