/*
 * enf_b200.h -- C ABI of libenf_b200.so: the B200 (sm_100a) implementation of the
 * batched trafo-chain hot path of bat/EuclidianNormalizingFlows.jl.
 *
 * The reference has no FFI of its own (it is pure Julia); the boundary it offers
 * is Julia multiple dispatch on a handful of generic functions.  Each entry point
 * below names the reference method(s) it stands in for (paths relative to the
 * reference repository).  The Julia-side `ccall` shim that binds them is
 * julia/EuclidianNormalizingFlowsB200.jl, described in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative enf_status on failure; the
 *     message is available from enf_last_error().  Nothing aborts or throws
 *     across the ABI.
 *   - sample matrices are D x N, column-major, leading dimension D, no strides
 *     (the memory of a Julia Matrix{T}); sample j is the contiguous column j.
 *   - dtype is all-Float32 or all-Float64 per chain; mixed promotion
 *     (src/center_stretch.jl:5) is the shim's job.
 *   - parameters are passed in struct-field order, every field expanded to a
 *     length-D vector: CenterStretch/CenterContract a,b,c; JohnsonTrafo(Inv)
 *     gamma,delta,xi,lambda; ScaleShiftTrafo a,b; HouseholderTrafo V (D x K,
 *     column-major).  Gradients come back in the same packed layout.
 *   - device buffers are caller-owned handles (enf_alloc / enf_free); host
 *     arrays are borrowed for the duration of the call only.
 *   - every entry point selects its context's device itself and is re-entrant
 *     per context; launches are asynchronous on the context stream, only the
 *     calls documented as blocking wait for the device.
 *   - there is no CPU fallback: without a usable CUDA device enf_init fails.
 */
#ifndef ENF_B200_H
#define ENF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct enf_ctx enf_ctx;
typedef struct enf_chain enf_chain;

typedef enum {
    ENF_OK = 0,
    ENF_ERR_INVALID = -1,   /* bad argument / unsupported shape          */
    ENF_ERR_CUDA = -2,      /* CUDA runtime error                        */
    ENF_ERR_NCCL = -3,      /* NCCL error or libnccl not loadable        */
    ENF_ERR_NOMEM = -4
} enf_status;

typedef enum { ENF_F32 = 0, ENF_F64 = 1 } enf_dtype;

/* One trafo of a chain.  Ops are listed in APPLICATION order (innermost of the
 * Julia `f1 ∘ f2 ∘ ... ∘ fn` first), i.e. the flattening of
 * Base.ComposedFunction that ChangesOfVariables walks. */
typedef enum {
    ENF_CENTER_STRETCH = 0,  /* src/center_stretch.jl:25-45   */
    ENF_CENTER_CONTRACT = 1, /* src/center_stretch.jl:49-69   */
    ENF_JOHNSON = 2,         /* src/johnson_trafo.jl:61-82    */
    ENF_JOHNSON_INV = 3,     /* src/johnson_trafo.jl:86-107   */
    ENF_SCALE_SHIFT = 4,     /* src/scale_shift_trafo.jl:4-30 */
    ENF_HOUSEHOLDER = 5      /* src/householder_trafo.jl:127-160 (V: D x K; K = 1 for a vector V) */
} enf_op_kind;

typedef struct {
    int32_t kind;        /* enf_op_kind                                      */
    int32_t K;           /* number of reflections (HOUSEHOLDER), else 0      */
    const void* params;  /* host pointer, chain dtype, layout as above       */
} enf_op;

/* flags for enf_negll_grad* */
enum {
    /* Report the loss value Zygote's forward pass produces: rrule(similar_fill)
     * returns zeros as its primal (src/abstract_trafo.jl:30-33), so the
     * ScaleShiftTrafo ladj is missing from the value (not from the gradient).
     * This is what optimize_whitening records in negll_history. */
    ENF_NEGLL_ZYGOTE_PRIMAL = 1
};

/* ---- context, errors, memory ------------------------------------------------ */
int enf_init(int device, enf_ctx** out);
int enf_destroy(enf_ctx* ctx);
/* ctx may be NULL: returns the last error of the calling thread. */
const char* enf_last_error(const enf_ctx* ctx);
int enf_device_count(int* n);
int enf_sync(enf_ctx* ctx);                                   /* blocking */

int enf_alloc(enf_ctx* ctx, size_t bytes, void** dptr);
int enf_free(enf_ctx* ctx, void* dptr);
int enf_host_alloc(enf_ctx* ctx, size_t bytes, void** hptr); /* pinned host memory, placed on the GPU's NUMA node */
int enf_host_free(enf_ctx* ctx, void* hptr);
int enf_h2d(enf_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes); /* async on ctx stream */
int enf_d2h(enf_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* blocking */
int enf_memset(enf_ctx* ctx, void* dst_dev, int value, size_t bytes);

/* Fill a D x N device matrix with the benchmark's synthetic N(0,1) samples
 * (counter-based Philox4x32-10 keyed by seed, element index = i + D*(col0 + j);
 * Box-Muller), so every GPU count and the CPU oracle see identical data
 * (SURVEY §8d). */
int enf_fill_normal(enf_ctx* ctx, int dtype, void* x_dev, int D, int64_t N, int64_t col0, uint64_t seed);

/* Element-wise conversion between the two sample types on the device (n elements, async on the context's stream).
 * Replaces the promotion the reference does implicitly, `float(promote_type(eltype(x), eltype(params)...))`
 * (src/center_stretch.jl:5, src/johnson_trafo.jl:30, src/scale_shift_trafo.jl:15): a chain is all-Float32 or
 * all-Float64, so Float32 samples meeting Float64 parameters are widened with this call before the chain runs. */
int enf_convert(enf_ctx* ctx, int dst_dtype, void* dst_dev, int src_dtype, const void* src_dev, int64_t n);

/* ---- chains -------------------------------------------------------------------
 * A chain is the flattened op list of a trafo tree plus device-resident derived
 * constants.  Replaces dispatch on the trafo structs themselves. */
int enf_chain_create(enf_ctx* ctx, int dtype, int D, int n_ops, const enf_op* ops, enf_chain** out);
/* packed_params: all ops' params concatenated in op order (host, chain dtype). */
int enf_chain_set_params(enf_chain* chain, const void* packed_params);
int enf_chain_num_params(const enf_chain* chain, int64_t* n_params);
int enf_chain_destroy(enf_chain* chain);

/* (f::Trafo)(x): src/center_stretch.jl:37,61; src/johnson_trafo.jl:74,99;
 * src/scale_shift_trafo.jl:15-16; src/householder_trafo.jl:156-157 -- and their
 * composition.  y may alias x. */
int enf_forward(enf_chain* chain, const void* x_dev, int64_t N, void* y_dev);

/* ChangesOfVariables.with_logabsdet_jacobian(f, x): src/center_stretch.jl:39,63;
 * src/johnson_trafo.jl:76,101; src/scale_shift_trafo.jl:18; src/householder_trafo.jl:159-160;
 * sum_ladjs src/abstract_trafo.jl:7-9.  ladj_dev: N values (the 1 x N row). */
int enf_forward_ladj(enf_chain* chain, const void* x_dev, int64_t N, void* y_dev, void* ladj_dev);

/* Same through HOST buffers (the call a host-matrix user makes): chunked,
 * H2D / kernel / D2H overlapped on three streams.  Blocking.  Buffers from
 * enf_host_alloc are copied at full PCIe rate; pageable memory also works. */
int enf_forward_ladj_host(enf_chain* chain, const void* x_host, int64_t N, void* y_host, void* ladj_host);

/* mvnormal_negll_trafo(trafo, X): src/optimize_whitening.jl:7-15.  Blocking. */
int enf_negll(enf_chain* chain, const void* x_dev, int64_t N, double* negll_host);

/* mvnormal_negll_trafograd(trafo, X): src/optimize_whitening.jl:18-22 (Zygote
 * pullback + the rrules of src/householder_trafo.jl:22-124 and
 * src/abstract_trafo.jl:17-33).  grads_host: packed like the params, chain
 * dtype.  Blocking.  Float32 chains of only Householder / ScaleShift ops at
 * D = 64, 128, 256 take the second-moment path (tensor-core S = X X^T, chain
 * rule on the moment matrix): x_dev must then be 16-byte aligned. */
int enf_negll_grad(enf_chain* chain, const void* x_dev, int64_t N, int flags,
                   double* negll_host, void* grads_host);

/* Two-phase form used for sharded batches: phase 1 leaves this rank's
 * un-normalised partial sums (float64, `n` of them) in a device buffer owned by
 * the chain; the caller all-reduces them (any transport; enf_negll_grad_group
 * does all three steps with the library's own exchange); phase 2 maps the summed vector to (negll, grads) for the GLOBAL
 * batch size. */
int enf_negll_grad_partial(enf_chain* chain, const void* x_dev, int64_t N_local,
                           double** sums_dev, int64_t* n);
int enf_negll_grad_finish(enf_chain* chain, const double* sums_host, int64_t N_global, int flags,
                          double* negll_host, void* grads_host);

/* ---- the callers next to the path ---------------------------------------------
 * Target log-density of the variational objective, applied element-wise like `my_ll.(z)` of
 * examples/nf_variational_1d.jl:25-27 (no callbacks cross the ABI: a fixed set of densities).
 * ENF_TARGET_GAUSS_MIXTURE: p(z) = sum_k weights[k] N(z | means[k], sigmas[k]), 1 <= K <= 8. */
typedef enum { ENF_TARGET_GAUSS_MIXTURE = 1 } enf_target_kind;
typedef struct {
    int32_t kind;            /* enf_target_kind */
    int32_t K;               /* number of mixture components */
    const double* weights;   /* host pointers, K values each */
    const double* means;
    const double* sigmas;
} enf_target;

/* nELBO(trafo, xi) and nELBO_trafograd(trafo, xi): examples/nf_variational_1d.jl:29-47,
 *   nELBO = -[(sum_ij log p(z_ij) + sum_j ladj_j) / N - (log 2 pi + 1)/2 * D],  (z, ladj) = with_logabsdet_jacobian(trafo, xi)
 * for a D x N matrix xi of standard-normal draws (samples are columns; the example builds a (2 batchsize) x 1 matrix
 * with the roles of rows and columns swapped, :32-34).  One fused kernel: forward + ladj + target log-density + reverse
 * pass.  grads_host: packed like the params, chain dtype; NULL: value only.  flags as for enf_negll_grad.  Blocking. */
int enf_elbo_grad(enf_chain* chain, const enf_target* target, const void* xi_dev, int64_t N, int flags,
                  double* nelbo_host, void* grads_host);

/* Batched operations of the JohnsonSU distribution object (src/johnson_trafo.jl:1-26,109-129), element-wise over N
 * values: out[i] = op(JohnsonSU(gamma, delta, xi, lambda), x[i]); params4 = {gamma, delta, xi, lambda}.  For
 * ENF_JSU_QUANTILE x holds probabilities.  Asynchronous on the context stream; out may alias x. */
typedef enum {
    ENF_JSU_PDF = 0,      /* Distributions.pdf      src/johnson_trafo.jl:120 */
    ENF_JSU_LOGPDF = 1,   /* Distributions.logpdf   :123 */
    ENF_JSU_CDF = 2,      /* Distributions.cdf      :121 */
    ENF_JSU_LOGCDF = 3,   /* Distributions.logcdf   :124 */
    ENF_JSU_CCDF = 4,     /* Distributions.ccdf     :125 */
    ENF_JSU_LOGCCDF = 5,  /* Distributions.logccdf  :126 */
    ENF_JSU_QUANTILE = 6  /* Statistics.quantile    :129 */
} enf_johnsonsu_op;
int enf_johnsonsu(enf_ctx* ctx, int dtype, int op, const double* params4, const void* x_dev, int64_t N, void* out_dev);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------------
 * The only collective on the path is the all-reduce of the loss and the
 * parameter-gradient sums.  libnccl.so.2 is dlopen'ed on first use.
 * enf_group_init also maps a small exchange buffer of every rank into every
 * process (CUDA IPC); sums of up to 64 KB are then all-reduced by one kernel over
 * NVLink peer memory, larger ones (and every one when the mapping is not
 * possible or ENF_NO_P2P is set) by ncclAllReduce. */
#define ENF_UNIQUE_ID_BYTES 128
int enf_group_unique_id(void* id_out /* ENF_UNIQUE_ID_BYTES */);
int enf_group_init(enf_ctx* ctx, int nranks, int rank, const void* id);
int enf_group_destroy(enf_ctx* ctx);
/* negll + grads of the global batch whose columns are sharded over the group:
 * enf_negll_grad_partial -> ncclAllReduce(sum, f64) -> enf_negll_grad_finish.
 * N_global = sum of N_local over ranks (all-reduced together with the sums). */
int enf_negll_grad_group(enf_chain* chain, const void* x_dev, int64_t N_local, int flags,
                         double* negll_host, void* grads_host);
/* The exchange step on its own: sums the raw sums that enf_negll_grad_partial left on the device (and N_local, appended
 * as one more value) over the group, in place, asynchronously on the context stream. */
int enf_group_allreduce_sums(enf_chain* chain, int64_t N_local);

/* ---- optimize_whitening on the device -----------------------------
 * The whole loop of src/optimize_whitening.jl:36-43 -- for epoch, for batch: (negll, grad) ->
 * Optimisers.update -> push!(negll_hist) -- with parameters, optimizer state and loss history kept
 * on the device: two kernel launches per step (captured per epoch into a CUDA graph), no host round trip.  Batches are the contiguous
 * column ranges of src/optimize_whitening.jl:31-32 (batchsize = round(Int, N / nbatches)).
 * Optimizer: ADAGrad with the Optimisers.jl 0.2 rule (eta, epsilon; state starts at epsilon);
 * HouseholderTrafo columns are re-normalised after every update (src/householder_trafo.jl:134-146).
 *   state_inout : packed like the parameters, float64; on entry the accumulated state to continue from
 *                 (ignored when fresh_state != 0), on exit the final state.  May not be NULL.
 *   params_out  : final parameters, packed, chain dtype.  The chain itself is left holding them.
 *   history_out : one loss per step (nepochs * number of batches values; the Zygote-primal value when
 *                 flags has ENF_NEGLL_ZYGOTE_PRIMAL, like negll_history of the reference).
 *   use_group   : != 0: x holds this rank's columns of every batch; sums are all-reduced over the
 *                 context's NCCL group each step (enf_group_init), every rank applies the same update.
 * Blocking. */
int enf_optimize_whitening(enf_chain* chain, const void* x_dev, int64_t N, int64_t nbatches, int64_t nepochs,
                           double eta, double epsilon, int flags, int use_group, int fresh_state,
                           double* state_inout, void* params_out, double* history_out, int64_t* n_steps_out);

/* The same loop over explicitly given batches: batch b is the next local_counts[b] consecutive columns of x.  This is the
 * form for sharded data (use_group != 0): x holds this rank's columns of every global batch in batch order and
 * local_counts[b] how many of them belong to global batch b (zero is allowed: the rank then only takes part in the
 * exchange of that step).  Every rank must pass the same n_batches; the library compares it over the group before
 * anything is launched and fails on every rank (ENF_ERR_INVALID) if they differ.  enf_optimize_whitening(N, nbatches)
 * is this call with the counts of src/optimize_whitening.jl:31-32 derived from the LOCAL N. */
int enf_optimize_whitening_batches(enf_chain* chain, const void* x_dev, int64_t n_batches, const int64_t* local_counts,
                                   int64_t nepochs, double eta, double epsilon, int flags, int use_group, int fresh_state,
                                   double* state_inout, void* params_out, double* history_out, int64_t* n_steps_out);

/* ---- timing on the context stream ---------------------------------------------
 * The library launches on its own stream, which events of other libraries do
 * not see.  ENF_N_EVENTS CUDA events per context: record one, later read the
 * elapsed device time between two recorded events (blocks until `b` is done). */
#define ENF_N_EVENTS 16
int enf_event_record(enf_ctx* ctx, int slot);
int enf_event_elapsed_ms(enf_ctx* ctx, int slot_a, int slot_b, float* ms);

/* ---- introspection used by bench.py / tests ----------------------------------- */
/* Number of kernel launches issued by this context so far. */
int enf_launch_count(const enf_ctx* ctx, int64_t* n);
/* Name of the kernel variant the dispatcher picks for (chain, N, pointers). */
int enf_chain_describe(const enf_chain* chain, char* buf, size_t buflen);
int enf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ENF_B200_H */
