#!/usr/bin/env python
"""Kernel-level timing of one chain on resident synthetic samples (CUDA events on
the library stream).  Used for the optimisation loop and for ncu captures:

  python tools/microbench.py --spec hh4,jo,cs --D 16 --N 20000000 --what fwd_ladj
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", default="hh4,jo,cs")
    ap.add_argument("--D", type=int, default=16)
    ap.add_argument("--N", type=int, default=20_000_000)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--what", default="fwd_ladj", choices=["fwd", "fwd_ladj", "negll", "grad"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import ctypes as C
    import enf_b200 as E
    from enf_b200 import _lib as L
    from chains import build
    dt = np.float32 if args.dtype == "f32" else np.float64
    ctx = E.default_context()
    f = build(E, args.spec.split(","), args.D, np.random.default_rng(42), dt)
    X = E.B200Matrix.randn(args.D, args.N, dt, ctx=ctx)
    Y = X.empty_like()
    Ld = E.B200Matrix(ctx, 1, args.N, dt)
    ch = E.get_chain(f, args.D, dt, ctx)
    lib = ctx._lib
    sums, n = C.c_void_p(), C.c_int64()

    def run():
        if args.what == "fwd":
            L.check(lib.enf_forward(ch.handle, C.c_void_p(X.ptr), args.N, C.c_void_p(Y.ptr)))
        elif args.what == "fwd_ladj":
            L.check(lib.enf_forward_ladj(ch.handle, C.c_void_p(X.ptr), args.N, C.c_void_p(Y.ptr), C.c_void_p(Ld.ptr)))
        elif args.what == "negll":
            out = C.c_double()
            L.check(lib.enf_negll(ch.handle, C.c_void_p(X.ptr), args.N, C.byref(out)))
        else:
            L.check(lib.enf_negll_grad_partial(ch.handle, C.c_void_p(X.ptr), args.N, C.byref(sums), C.byref(n)))

    for _ in range(args.warmup):
        run()
    ctx.sync()
    ctx.record(0)
    for _ in range(args.iters):
        run()
    ctx.record(1)
    ms = ctx.elapsed_ms(0, 1) / args.iters
    s = dt().itemsize
    bytes_per_sample = ((2 * args.D + 1) if args.what == "fwd_ladj" else 2 * args.D if args.what == "fwd" else args.D) * s
    gbs = bytes_per_sample * args.N / (ms * 1e-3) / 1e9
    print(f"{args.what} spec={args.spec} D={args.D} N={args.N} {args.dtype}: {ms:.4f} ms  "
          f"{args.N / (ms * 1e-3):.4g} samples/s  {gbs:.1f} GB/s algorithmic ({gbs / 6543.4:.3f} of measured HBM peak)  "
          f"[{ch.describe()}]")


if __name__ == "__main__":
    main()
