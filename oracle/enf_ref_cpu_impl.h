/* Included twice by enf_ref_cpu.c with REAL = float / double and SUF = f32 / f64.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). */

/* one elementwise trafo the way the reference evaluates it: a broadcast pass
 * that allocates y, a second broadcast pass that allocates a D x N ladj
 * temporary, and a column reduction (src/center_stretch.jl:39-43,
 * src/johnson_trafo.jl:76-80, src/abstract_trafo.jl:9). */
static void NAME(elementwise)(int kind, int D, int64_t N, const REAL* p, const REAL* x, REAL* y, REAL* tmp, REAL* ladj) {
    const REAL *p0 = p, *p1 = p + D, *p2 = p + 2 * D, *p3 = p + 3 * D;
    /* pass 1: y = f.(x, params...) */
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) {
        const REAL* xj = x + j * D;
        REAL* yj = y + j * D;
        for (int i = 0; i < D; ++i) {
            switch (kind) {
                case 0: yj[i] = NAME(center_stretch)(xj[i], p0[i], p1[i], p2[i]); break;
                case 1: yj[i] = NAME(center_contract)(xj[i], p0[i], p1[i], p2[i]); break;
                case 2: yj[i] = NAME(johnsontrafo)(xj[i], p0[i], p1[i], p2[i], p3[i]); break;
                case 3: yj[i] = NAME(johnsontrafo_inv)(xj[i], p0[i], p1[i], p2[i], p3[i]); break;
                default: yj[i] = FMA(xj[i], p0[i], p1[i]); break;   /* muladd.(x, a, b) src/scale_shift_trafo.jl:16 */
            }
        }
    }
    if (kind == 4) { /* src/scale_shift_trafo.jl:22-23: ladj = sum(log.(abs.(a))) filled into N slots */
        REAL l = 0;
        for (int i = 0; i < D; ++i) l += LOG(FABS(p0[i]));
        for (int64_t j = 0; j < N; ++j) ladj[j] = l;
        return;
    }
    /* pass 2: ladjs = ladj_fn.(x or y, params...)  (D x N temporary) */
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) {
        const REAL* xj = x + j * D;
        const REAL* yj = y + j * D;
        REAL* tj = tmp + j * D;
        for (int i = 0; i < D; ++i) {
            switch (kind) {
                case 0: tj[i] = NAME(center_contract_ladj)(yj[i], p0[i], p1[i], p2[i]); break;   /* at y, negated below */
                case 1: tj[i] = NAME(center_contract_ladj)(xj[i], p0[i], p1[i], p2[i]); break;
                case 2: tj[i] = NAME(johnsontrafo_ladj)(xj[i], p0[i], p1[i], p2[i], p3[i]); break;
                default: tj[i] = NAME(johnsontrafo_ladj)(yj[i], p0[i], p1[i], p2[i], p3[i]); break; /* at y, negated below */
            }
        }
    }
    /* pass 3: sum_ladjs: vec(sum(ladjs, dims = 1))' */
    const REAL sgn = (kind == 0 || kind == 3) ? (REAL)-1 : (REAL)1;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) {
        const REAL* tj = tmp + j * D;
        REAL s = 0;
        for (int i = 0; i < D; ++i) s += tj[i];
        ladj[j] = sgn * s;
    }
}

/* householder_trafo!(y, v, x): k = (v'x)/(v'v) (a gemv producing a 1 x N row),
 * then y .= muladd.(-2 .* k, v, x) (src/householder_trafo.jl:4-11). */
static void NAME(householder)(int D, int64_t N, const REAL* v, const REAL* x, REAL* y, REAL* k) {
    REAL vv = 0;
    for (int i = 0; i < D; ++i) vv += v[i] * v[i];
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) {
        const REAL* xj = x + j * D;
        REAL s = 0;
        for (int i = 0; i < D; ++i) s += v[i] * xj[i];
        k[j] = s / vv;
    }
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) {
        const REAL* xj = x + j * D;
        REAL* yj = y + j * D;
        const REAL m2k = (REAL)-2 * k[j];
        for (int i = 0; i < D; ++i) yj[i] = FMA(m2k, v[i], xj[i]);
    }
}

/* with_logabsdet_jacobian of a flattened chain, ChangesOfVariables order: inner
 * first, ladjs added (one more N-pass per trafo).  y and ladj are outputs;
 * returns 0, or -1 when scratch allocation fails. */
int NAME(ref_forward_ladj)(int D, int64_t N, int n_ops, const int* kinds, const int* Ks, const REAL* const* params,
                           const REAL* x, REAL* y, REAL* ladj) {
    const size_t bytes = (size_t)D * (size_t)N * sizeof(REAL);
    REAL* a = (REAL*)malloc(bytes ? bytes : 1);        /* ping-pong activations: every trafo allocates its output */
    REAL* b = (REAL*)malloc(bytes ? bytes : 1);
    REAL* tmp = (REAL*)malloc(bytes ? bytes : 1);      /* the D x N ladj temporary */
    REAL* l1 = (REAL*)malloc((size_t)(N ? N : 1) * sizeof(REAL));
    if (!a || !b || !tmp || !l1) { free(a); free(b); free(tmp); free(l1); return -1; }
    for (int64_t j = 0; j < N; ++j) ladj[j] = 0;
    const REAL* cur = x;
    REAL* nxt = a;
    for (int o = 0; o < n_ops; ++o) {
        if (kinds[o] == 5) { /* chained_householder_trafo!: y .= x, then K in-place reflections (src/householder_trafo.jl:71-78) */
            memcpy(nxt, cur, bytes);
            for (int kk = 0; kk < Ks[o]; ++kk) NAME(householder)(D, N, params[o] + (size_t)kk * D, nxt, nxt, l1);
            /* ladj: similar_zeros -> adds zeros (src/householder_trafo.jl:160) */
#pragma omp parallel for schedule(static)
            for (int64_t j = 0; j < N; ++j) ladj[j] += 0;
        } else {
            NAME(elementwise)(kinds[o], D, N, params[o], cur, nxt, tmp, l1);
#pragma omp parallel for schedule(static)
            for (int64_t j = 0; j < N; ++j) ladj[j] += l1[j];
        }
        cur = nxt;
        nxt = (nxt == a) ? b : a;
    }
    memcpy(y, cur, bytes);
    free(a); free(b); free(tmp); free(l1);
    return 0;
}

/* mvnormal_negll_trafo (src/optimize_whitening.jl:7-15): wlaj, then
 * sum(std_normal_logpdf.(Y)) (allocates D x N) + sum(ladj), / nsamples, negated. */
int NAME(ref_negll)(int D, int64_t N, int n_ops, const int* kinds, const int* Ks, const REAL* const* params,
                    const REAL* x, double* out) {
    REAL* y = (REAL*)malloc((size_t)D * (size_t)(N ? N : 1) * sizeof(REAL));
    REAL* ladj = (REAL*)malloc((size_t)(N ? N : 1) * sizeof(REAL));
    if (!y || !ladj) { free(y); free(ladj); return -1; }
    int rc = NAME(ref_forward_ladj)(D, N, n_ops, kinds, Ks, params, x, y, ladj);
    if (rc == 0) {
        double s = 0, sl = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int64_t e = 0; e < (int64_t)D * N; ++e) s += -((double)y[e] * (double)y[e] + 1.8378770664093454835606594728112) / 2;
#pragma omp parallel for reduction(+ : sl) schedule(static)
        for (int64_t j = 0; j < N; ++j) sl += (double)ladj[j];
        *out = -(s + sl) / (double)N;
    }
    free(y); free(ladj);
    return rc;
}
