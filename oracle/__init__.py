"""CPU oracle for the batched trafo-chain path of bat/EuclidianNormalizingFlows.jl.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`euclidiannormalizingflows.jl_b200/`) may import, link or execute anything in
this directory.  Allowed callers: `tests/`, `__graft_entry__.smoke()`, and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.

Parity status:
  * The reference is pure Julia and Julia is not installed in this image, so
    `oracle/_ref` cannot be built: the reference is UNBUILDABLE here.
  * The scalar kernels are PINNED against the only four known-answer values in
    the reference's own tests (test/test_center_stretch.jl:18-19,
    test/test_johnson_trafo.jl:21-22) and against the property tests in
    test/*.jl re-expressed with finite differences / torch-f64 autograd.
  * PARITY UNPINNED for everything that lives in un-vendored Julia packages:
    ComposedFunction handling (ChangesOfVariables 0.1, InverseFunctions 0.1,
    Functors 0.2.5-0.4), Zygote reverse mode, Optimisers 0.2 (ADAGrad).  Their
    published semantics are restated and the restatement says so at each site.
"""
