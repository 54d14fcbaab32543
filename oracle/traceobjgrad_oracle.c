/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C, FP64, operation-for-operation CPU restatement of Juqbox.jl's Stormer-Verlet
 * objective + discrete-adjoint gradient (`traceobjgrad`, verbose=false).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the reference's
 * own golden files rabi/swap02/cnot2/cnot3/flux/cnot2-leakieq-ref.jld2 (extracted to
 * tests/golden/<case>.json) at the reference's tolerance rtol 1e-10 / atol 1e-14 (test/evalGrad.jl:4-5).
 * The reference itself (Julia) cannot be compiled or run in this image, so there is no
 * oracle/_ref; the goldens are the pin.
 *
 * Reference lines followed (all under /root/reference/src/):
 *   traceobjgrad driver ............ evalobjgrad.jl:504-1038
 *   step! (state, no forcing) ...... StormerVerlet.jl:461-504 (dense), :507-550 (sparse)
 *   step! (adjoint, forcing) ....... StormerVerlet.jl:255-303, :306-356
 *   step_no_forcing! ............... StormerVerlet.jl:365-406, :409-451
 *   neumann! ....................... linear_solvers.jl:81-106
 *   jacobi! ........................ linear_solvers.jl:110-152
 *   KS! / accumulate_matrix! ....... evalobjgrad.jl:2354-2441
 *   bcarrier2 / gradbcarrier2! ..... bsplines.jl:211-304, :321-415
 *   adjoint_grad_calc! ............. evalobjgrad.jl:2567-2619
 *   adjoint_trace_operator! ........ evalobjgrad.jl:2114-2154
 *   init_adjoint!, tracefid* ....... evalobjgrad.jl:2026-2111
 *   penalf2a / penalf2aTrap ........ evalobjgrad.jl:2170-2208
 *   risk-neutral H0 shift .......... ipopt_interface.jl:41-44 (passed in as a diagonal vector)
 *
 * Extensions beyond the golden-pinned core — PARITY UNPINNED BY THE REFERENCE: no reference test or
 * golden exercises them (pFidType is hard-wired to 2 by the constructor, evalobjgrad.jl:164; cnot-lab-ref.jld2 is not in the CI
 * list and its case script cannot build a bcparams with this revision, bsplines.jl:177-181 vs test/cases/cnot-lab-setup.jl:116).
 * They are restated from the cited lines and pinned by central finite differences of the restated objective (tests/):
 *   pFidType 1/3/4 + globalPhase ... objective evalobjgrad.jl:755-763; terminal condition :2026-2059; phase gradient :923-945.
 *       pFidType 1: init_adjoint! has NO branch for it (lambda(T) is left stale, i.e. undefined); the intended terminal condition
 *       (the pFidType-2 formula applied to scomplex0 = exp(i phase) - s, which :825-826 computes for exactly that purpose) is used.
 *   dense wmat_real / wmat_imag .... penalf2a / penalf2aTrap / penalf2imag dense methods :2183-2233; forcing :862,:882-888.
 *   uncoupled controls ............. KS! :2372-2387 (TWO splines p, q per control, ft = 2(p cos(2 pi Rfreq t) - q sin(...)), added to
 *       K if isSymm else to S).  adjoint_grad_calc! :2621-2654 indexes ONE spline per uncoupled control (qu = 2Nc-1+q) and omits
 *       d ft / d(p,q): it is not the gradient of KS!'s model.  unc_grad_literal = 0 (default): exact gradient of KS!'s model
 *       (the same trace combinations, multiplied by d ft/dp = 2cos, d ft/dq = -2sin); unc_grad_literal = 1: the literal lines.
 *
 * Deliberate deviations (none changes a result beyond round-off):
 *   - sparse KS! uses a precomputed position map instead of `A[row,j] +=` lookups (faster than the
 *     reference, so the timed CPU baseline is, if anything, too fast);
 *   - K's sparsity pattern always contains the diagonal (so a noise shift never changes the pattern).
 *
 * Layout: all matrices column-major (Julia), CSC 0-based.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    int n;
    int64_t nnz;
    int64_t *colptr; /* n+1 */
    int64_t *rowval; /* nnz */
    double *nzval;   /* nnz */
} csc_t;

typedef struct {
    /* problem */
    int n, m, Nc, Nfreq, D1, J, objFuncType, sparse;
    int Nunc, NcT;    /* uncoupled controls; NcT = Nc + Nunc = rows of Cfreq and number of (p, q) spline pairs */
    const int *isSymm;
    const double *Rfreq;
    int unc_grad_literal;
    int pFidType;
    double globalPhase;
    const double *wreal, *wimag;   /* dense n x n weights (col-major) or NULL: Diagonal(wdiag), zero imaginary part */
    int solver;       /* 1 = Neumann (J terms), 2 = Jacobi (J = max_iter, tol) — linear_solvers.jl:4-5 */
    double tol;
    double *jac_scaled; /* Jacobi: the S.*=coeff copy */
    csc_t jac_scaled_s;
    int64_t nsteps;
    double T;
    const double *Uinit, *Vtr, *Vti, *wdiag, *Cfreq;
    /* dense operators (col-major n*n), hsym/hanti: Nc consecutive matrices */
    double *H0d;
    const double *Hsymd, *Hantid, *Huncd;
    /* sparse operators */
    csc_t H0s, *Hsyms, *Hantis, *Huncs;
    int64_t **mapK_h0, **mapK_sym, **mapS_anti, **map_unc; /* nz position maps into K / S patterns */
    /* working arrays (Working_Arrays, evalobjgrad.jl:359-442) */
    double *K0d, *S0d, *K05d, *S05d, *K1d, *S1d;
    csc_t K0s, S0s, K05s, S05s, K1s, S1s;
    double *lambdar, *lambdar0, *lambdai, *lambdai0, *lambdar05;
    double *lambdar_n, *lambdar0_n, *lambdai_n, *lambdai0_n, *lambdar05_n;
    double *k1, *k2, *l1, *l2, *rhs, *hr0, *hi0, *hr1, *hi1, *vr, *vi, *vi05, *vr0;
    double *gr, *gi, *gradobjfadj, *tr_adj, *infidelgrad;
    int Npar;
    const double *pcof;
    double dtknot;
} ws_t;

/* ------------------------------------------------------------------ control functions */
/* bsplines.jl:211-304 */
static double bcarrier2(double t, const ws_t *w, int func) {
    int osc = func / 2, q_func = func % 2, D1 = w->D1, Nfreq = w->Nfreq;
    double f = 0.0, dtknot = w->dtknot, width = 3 * dtknot;
    int64_t k = (int64_t)ceil(t / dtknot + 2);
    if (k < 3) k = 3;
    if (k > D1) k = D1;
    for (int freq = 1; freq <= Nfreq; freq++) {
        double fbs1 = 0.0, fbs2 = 0.0;
        int64_t offset1 = 2 * osc * Nfreq * D1 + (freq - 1) * 2 * D1; /* 1-based k is added below */
        int64_t offset2 = offset1 + D1;
        const double *p = w->pcof - 1;
        double tc = dtknot * ((double)k - 1.5);
        double tau = (t - tc) / width;
        double b = 9.0 / 8 + 4.5 * tau + 4.5 * tau * tau;
        fbs1 += p[offset1 + k] * b;
        fbs2 += p[offset2 + k] * b;
        tc = dtknot * ((double)(k - 1) - 1.5);
        tau = (t - tc) / width;
        b = 0.75 - 9 * tau * tau;
        fbs1 += p[offset1 + k - 1] * b;
        fbs2 += p[offset2 + k - 1] * b;
        tc = dtknot * ((double)(k - 2) - 1.5);
        tau = (t - tc) / width;
        b = 9.0 / 8 - 4.5 * tau + 4.5 * tau * tau;
        fbs1 += p[offset1 + k - 2] * b;
        fbs2 += p[offset2 + k - 2] * b;
        double om = w->Cfreq[osc + w->NcT * (freq - 1)];
        if (q_func == 1)
            f += fbs1 * sin(om * t) + fbs2 * cos(om * t);
        else
            f += fbs1 * cos(om * t) - fbs2 * sin(om * t);
    }
    return f;
}

/* bsplines.jl:321-415 */
static void gradbcarrier2(double t, const ws_t *w, int func, double *g) {
    int osc = func / 2, q_func = func % 2, D1 = w->D1, Nfreq = w->Nfreq;
    memset(g, 0, sizeof(double) * w->Npar);
    double dtknot = w->dtknot, width = 3 * dtknot;
    int64_t k = (int64_t)ceil(t / dtknot + 2);
    if (k < 3) k = 3;
    if (k > D1) k = D1;
    g -= 1; /* 1-based */
    for (int freq = 1; freq <= Nfreq; freq++) {
        int64_t offset1 = 2 * osc * Nfreq * D1 + (freq - 1) * 2 * D1;
        int64_t offset2 = offset1 + D1;
        double om = w->Cfreq[osc + w->NcT * (freq - 1)];
        double s = sin(om * t), c = cos(om * t);
        for (int seg = 0; seg < 3; seg++) {
            double tc = dtknot * ((double)(k - seg) - 1.5);
            double tau = (t - tc) / width, bk;
            if (seg == 0) bk = 9.0 / 8 + 4.5 * tau + 4.5 * tau * tau;
            else if (seg == 1) bk = 0.75 - 9 * tau * tau;
            else bk = 9.0 / 8 - 4.5 * tau + 4.5 * tau * tau;
            if (q_func == 1) {
                g[offset1 + k - seg] = bk * s;
                g[offset2 + k - seg] = bk * c;
            } else {
                g[offset1 + k - seg] = bk * c;
                g[offset2 + k - seg] = -bk * s;
            }
        }
    }
}

/* ------------------------------------------------------------------ BLAS-like helpers */
static inline void axpy(int64_t len, double a, const double *x, double *y) {
    for (int64_t i = 0; i < len; i++) y[i] += a * x[i];
}
/* C = alpha*A*B + beta*C, A dense n x n col-major, B,C n x m */
static void gemm_d(int n, int m, const double *A, const double *B, double alpha, double beta, double *C) {
    for (int j = 0; j < m; j++) {
        double *c = C + (int64_t)j * n;
        const double *b = B + (int64_t)j * n;
        if (beta == 0.0) for (int i = 0; i < n; i++) c[i] = 0.0;
        else if (beta != 1.0) for (int i = 0; i < n; i++) c[i] *= beta;
        for (int k = 0; k < n; k++) {
            double ab = alpha * b[k];
            const double *a = A + (int64_t)k * n;
            for (int i = 0; i < n; i++) c[i] += a[i] * ab;
        }
    }
}
/* Julia SparseArrays mul!(C, A, B, alpha, beta) for CSC A */
static void spmm(int m, const csc_t *A, const double *B, double alpha, double beta, double *C) {
    int n = A->n;
    for (int j = 0; j < m; j++) {
        double *c = C + (int64_t)j * n;
        const double *b = B + (int64_t)j * n;
        if (beta == 0.0) for (int i = 0; i < n; i++) c[i] = 0.0;
        else if (beta != 1.0) for (int i = 0; i < n; i++) c[i] *= beta;
        for (int col = 0; col < n; col++) {
            double ab = b[col] * alpha;
            for (int64_t p = A->colptr[col]; p < A->colptr[col + 1]; p++) c[A->rowval[p]] += A->nzval[p] * ab;
        }
    }
}
typedef struct { const double *d; const csc_t *s; } op_t;
static inline void mul(const ws_t *w, double *C, op_t A, const double *B, double alpha, double beta) {
    if (w->sparse) spmm(w->m, A.s, B, alpha, beta, C);
    else gemm_d(w->n, w->m, A.d, B, alpha, beta, C);
}

/* linear_solvers.jl:81-106 — destroys B, uses Tm as scratch */
static void neumann(const ws_t *w, double h, op_t S, double *B, double *Tm, double *X) {
    int64_t len = (int64_t)w->n * w->m;
    memcpy(X, B, sizeof(double) * len);
    memcpy(Tm, B, sizeof(double) * len);
    double coeff = 1.0;
    for (int j = 1; j <= w->J; j++) {
        mul(w, Tm, S, B, 1.0, 0.0);
        coeff *= (0.5 * h);
        axpy(len, coeff, Tm, X);
        memcpy(B, Tm, sizeof(double) * len);
    }
}

/* linear_solvers.jl:110-152 (jacobi!, dense and sparse twins): X = B; repeat T = B - (-h/2 S) X; err = ||T - X||_F;
 * X = T; until err < tol or max_iter sweeps.  The reference scales S in place by -h/2 and restores it on exit; the
 * scaled copy lives in scratch here so that S itself keeps its bits (S*coeff/coeff need not round-trip). B survives. */
static void jacobi(ws_t *w, double h, op_t S, const double *B, double *Tm, double *X) {
    int64_t len = (int64_t)w->n * w->m;
    double coeff = -0.5 * h;
    op_t Sc;
    if (w->sparse) {
        csc_t *d = &w->jac_scaled_s;
        if (!d->nzval || d->nnz < S.s->nnz) { free(d->nzval); d->nzval = (double *)malloc(sizeof(double) * (S.s->nnz + 1)); }
        d->n = S.s->n; d->nnz = S.s->nnz; d->colptr = S.s->colptr; d->rowval = S.s->rowval;
        for (int64_t k = 0; k < S.s->nnz; k++) d->nzval[k] = S.s->nzval[k] * coeff;
        Sc.s = d; Sc.d = NULL;
    } else {
        int64_t nn = (int64_t)w->n * w->n;
        if (!w->jac_scaled) w->jac_scaled = (double *)malloc(sizeof(double) * nn);
        for (int64_t k = 0; k < nn; k++) w->jac_scaled[k] = S.d[k] * coeff;
        Sc.d = w->jac_scaled; Sc.s = NULL;
    }
    memcpy(X, B, sizeof(double) * len);
    for (int j = 1; j <= w->J; j++) {
        mul(w, Tm, Sc, X, 1.0, 0.0);
        double err = 0.0;
        for (int64_t k = 0; k < len; k++) { Tm[k] = B[k] - Tm[k]; double d = Tm[k] - X[k]; err += d * d; }
        memcpy(X, Tm, sizeof(double) * len);
        if (sqrt(err) < w->tol) return;
    }
}

/* linear_solver.solve(h, S, rhs, T, X) of the steppers (StormerVerlet.jl:266,285,470,487) */
static void solve(ws_t *w, double h, op_t S, double *B, double *Tm, double *X) {
    if (w->solver == 2) jacobi(w, h, S, B, Tm, X);
    else neumann(w, h, S, B, Tm, X);
}

/* ------------------------------------------------------------------ KS! */
static void KS(ws_t *w, int level, double t) {
    int n = w->n, Nc = w->Nc;
    if (!w->sparse) {
        double *K = level == 0 ? w->K0d : level == 1 ? w->K05d : w->K1d;
        double *S = level == 0 ? w->S0d : level == 1 ? w->S05d : w->S1d;
        int64_t nn = (int64_t)n * n;
        memcpy(K, w->H0d, sizeof(double) * nn);
        memset(S, 0, sizeof(double) * nn);
        for (int q = 0; q < Nc; q++) {
            double pt = bcarrier2(t, w, 2 * q), qt = bcarrier2(t, w, 2 * q + 1);
            axpy(nn, pt, w->Hsymd + q * nn, K);
            axpy(nn, qt, w->Hantid + q * nn, S);
        }
        for (int q = 0; q < w->Nunc; q++) {                      /* evalobjgrad.jl:2372-2387 */
            int qs = 2 * Nc + 2 * q, qa = qs + 1;
            double pt = bcarrier2(t, w, qs), qt = bcarrier2(t, w, qa);
            double ft = 2 * (pt * cos(2 * M_PI * w->Rfreq[q] * t) - qt * sin(2 * M_PI * w->Rfreq[q] * t));
            axpy(nn, ft, w->Huncd + q * nn, w->isSymm[q] ? K : S);
        }
    } else {
        csc_t *K = level == 0 ? &w->K0s : level == 1 ? &w->K05s : &w->K1s;
        csc_t *S = level == 0 ? &w->S0s : level == 1 ? &w->S05s : &w->S1s;
        memset(K->nzval, 0, sizeof(double) * K->nnz);
        for (int64_t p = 0; p < w->H0s.nnz; p++) K->nzval[w->mapK_h0[0][p]] += 1.0 * w->H0s.nzval[p];
        memset(S->nzval, 0, sizeof(double) * S->nnz);
        for (int q = 0; q < Nc; q++) {
            double pt = bcarrier2(t, w, 2 * q), qt = bcarrier2(t, w, 2 * q + 1);
            for (int64_t p = 0; p < w->Hsyms[q].nnz; p++) K->nzval[w->mapK_sym[q][p]] += pt * w->Hsyms[q].nzval[p];
            for (int64_t p = 0; p < w->Hantis[q].nnz; p++) S->nzval[w->mapS_anti[q][p]] += qt * w->Hantis[q].nzval[p];
        }
        for (int q = 0; q < w->Nunc; q++) {                      /* evalobjgrad.jl:2408-2424 */
            int qs = 2 * Nc + 2 * q, qa = qs + 1;
            double pt = bcarrier2(t, w, qs), qt = bcarrier2(t, w, qa);
            double ft = 2 * (pt * cos(2 * M_PI * w->Rfreq[q] * t) - qt * sin(2 * M_PI * w->Rfreq[q] * t));
            csc_t *A = w->isSymm[q] ? K : S;
            for (int64_t p = 0; p < w->Huncs[q].nnz; p++) A->nzval[w->map_unc[q][p]] += ft * w->Huncs[q].nzval[p];
        }
    }
}
static inline op_t opK(const ws_t *w, int level) {
    op_t o;
    o.d = level == 0 ? w->K0d : level == 1 ? w->K05d : w->K1d;
    o.s = level == 0 ? &w->K0s : level == 1 ? &w->K05s : &w->K1s;
    return o;
}
static inline op_t opS(const ws_t *w, int level) {
    op_t o;
    o.d = level == 0 ? w->S0d : level == 1 ? w->S05d : w->S1d;
    o.s = level == 0 ? &w->S0s : level == 1 ? &w->S05s : &w->S1s;
    return o;
}

/* ------------------------------------------------------------------ steppers */
/* StormerVerlet.jl:461-504 */
static double step_state(ws_t *w, double t, double *u, double *v, double *v05, double h) {
    int64_t len = (int64_t)w->n * w->m;
    op_t K0 = opK(w, 0), S0 = opS(w, 0), K05 = opK(w, 1), S05 = opS(w, 1), K1 = opK(w, 2), S1 = opS(w, 2);
    double *k1 = w->k1, *k2 = w->k2, *l1 = w->l1, *l2 = w->l2, *rhs = w->rhs;
    mul(w, rhs, K05, u, 1.0, 0.0);
    mul(w, rhs, S05, v, 1.0, 1.0);
    solve(w, h, S05, rhs, v05, l1);
    memcpy(v05, v, sizeof(double) * len);
    axpy(len, 0.5 * h, l1, v05);
    mul(w, k1, S0, u, 1.0, 0.0);
    mul(w, k1, K0, v05, -1.0, 1.0);
    mul(w, rhs, S1, u, 1.0, 0.0);
    mul(w, rhs, S1, k1, 0.5 * h, 1.0);
    mul(w, rhs, K1, v05, -1.0, 1.0);
    axpy(len, 0.5 * h, k1, u);
    solve(w, h, S1, rhs, k1, k2);
    axpy(len, 0.5 * h, k2, u);
    mul(w, l2, K05, u, 1.0, 0.0);
    mul(w, l2, S05, v05, 1.0, 1.0);
    for (int64_t i = 0; i < len; i++) v[i] = v[i] + 0.5 * h * (l1[i] + l2[i]);
    return t + h;
}

/* StormerVerlet.jl:255-303 (forcing != NULL) and :365-406 (forcing == NULL) */
static void step_adjoint(ws_t *w, double *mu, double *nu, double *X, double h, const double *uf0, const double *vf0,
                         const double *uf1, const double *vf1) {
    int64_t len = (int64_t)w->n * w->m;
    op_t K0 = opK(w, 0), S0 = opS(w, 0), K05 = opK(w, 1), S05 = opS(w, 1), K1 = opK(w, 2), S1 = opS(w, 2);
    double *k1 = w->k1, *k2 = w->k2, *l1 = w->l1, *l2 = w->l2, *rhs = w->rhs;
    mul(w, rhs, S0, mu, 1.0, 0.0);
    mul(w, rhs, K05, nu, -1.0, 1.0);
    if (uf0) axpy(len, 1.0, uf0, rhs);
    solve(w, h, S0, rhs, k1, k2);
    axpy(len, 0.5 * h, k2, mu);
    memcpy(X, mu, sizeof(double) * len);
    mul(w, l2, K0, X, 1.0, 0.0);
    mul(w, l2, S05, nu, 1.0, 1.0);
    if (vf0) axpy(len, 1.0, vf0, l2);
    mul(w, rhs, S05, nu, 1.0, 0.0);
    mul(w, rhs, S05, l2, 0.5 * h, 1.0);
    mul(w, rhs, K1, X, 1.0, 1.0);
    if (vf1) axpy(len, 1.0, vf1, rhs);
    solve(w, h, S05, rhs, k2, l1);
    for (int64_t i = 0; i < len; i++) nu[i] = nu[i] + (0.5 * h) * (l2[i] + l1[i]);
    mul(w, k1, S1, X, 1.0, 0.0);
    mul(w, k1, K05, nu, -1.0, 1.0);
    if (uf1) axpy(len, 1.0, uf1, k1);
    axpy(len, 0.5 * h, k1, mu);
}

/* ------------------------------------------------------------------ reductions */
static double trace4(const ws_t *w, const double *A, const double *B, const double *C, double sgnC, const double *D) {
    double tr = 0.0;
    int64_t len = (int64_t)w->n * w->m;
    for (int64_t i = 0; i < len; i++) tr += A[i] * B[i] + (sgnC * C[i]) * D[i];
    return tr;
}
/* tracefidcomplex(vr, -vi, vtr, vti), evalobjgrad.jl:2078-2084: ur = vr, ui = -vi */
static void tracefidcomplex(const ws_t *w, const double *vr, const double *vi, double *re, double *im) {
    double N = (double)w->m;
    *re = trace4(w, vr, w->Vtr, vi, -1.0, w->Vti) / N;  /* tr(ur'vtr + ui'vti) */
    *im = trace4(w, vr, w->Vti, vi, +1.0, w->Vtr) / N;  /* tr(ur'vti - ui'vtr) = tr(ur'vti + vi'vtr) */
}
static double tracefidabs2(const ws_t *w, const double *vr, const double *vi) {
    double re, im;
    tracefidcomplex(w, vr, vi, &re, &im);
    return re * re + im * im;
}
/* dense-weight helpers: sum_ij A[i,j] (W B)[i,j] */
static double quadform(const ws_t *w, const double *A, const double *W, const double *B) {
    double f = 0.0;
    int n = w->n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < w->m; j++) {
            double a = A[i + (int64_t)j * n];
            for (int k = 0; k < n; k++) f += a * W[i + (int64_t)k * n] * B[k + (int64_t)j * n];
        }
    return f;
}
/* evalobjgrad.jl:2199-2208 (Diagonal), :2210-2223 (dense) */
static double penalf2aTrap(const ws_t *w, const double *vr) {
    if (w->wreal) return quadform(w, vr, w->wreal, vr);
    double f = 0.0;
    for (int j = 0; j < w->m; j++)
        for (int i = 0; i < w->n; i++) { double x = vr[i + (int64_t)j * w->n]; f += w->wdiag[i] * x * x; }
    return f;
}
/* evalobjgrad.jl:2170-2180 (Diagonal), :2183-2197 (dense) */
static double penalf2a(const ws_t *w, const double *vr, const double *vi) {
    if (w->wreal) return quadform(w, vr, w->wreal, vr) + 2.0 * quadform(w, vi, w->wreal, vi);
    double f = 0.0;
    for (int j = 0; j < w->m; j++)
        for (int i = 0; i < w->n; i++) {
            double x = vr[i + (int64_t)j * w->n], y = vi[i + (int64_t)j * w->n];
            f += (x * x + 2.0 * y * y) * w->wdiag[i];
        }
    return f;
}
/* penalf2imag(vr0, vi05, wmat_imag) = tr(vi05' Wi vr0), evalobjgrad.jl:2226-2233 (0 for a Diagonal wmat_imag) */
static double penalf2imag(const ws_t *w, const double *vr, const double *vi) {
    return w->wimag ? quadform(w, vi, w->wimag, vr) : 0.0;
}
/* C = alpha * W * B + beta * C for the dense weights (mul!(h, wmat, v, tinv, beta), evalobjgrad.jl:862,882-888) */
static void wmul(const ws_t *w, double *C, const double *W, const double *B, double alpha, double beta) {
    gemm_d(w->n, w->m, W, B, alpha, beta, C);
}
/* tr(A' * H * C): evalobjgrad.jl:2114-2131 (dense), :2135-2154 (sparse).  anti: 0 Hsym_q, 1 Hanti_q, 2 Hunc_q */
static double adjoint_trace(const ws_t *w, const double *A, int q, int anti, const double *C) {
    int n = w->n, m = w->m;
    double trace = 0.0;
    if (!w->sparse) {
        const double *B = (anti == 2 ? w->Huncd : anti ? w->Hantid : w->Hsymd) + (int64_t)q * n * n;
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n; i++) {
                double Btmp = 0.0;
                for (int k = 0; k < n; k++) Btmp += B[i + (int64_t)k * n] * C[k + (int64_t)j * n];
                trace += A[i + (int64_t)j * n] * Btmp;
            }
    } else {
        const csc_t *B = anti == 2 ? &w->Huncs[q] : anti ? &w->Hantis[q] : &w->Hsyms[q];
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n; i++) {
                double mat_temp = 0.0;
                for (int64_t p = B->colptr[i]; p < B->colptr[i + 1]; p++) mat_temp += A[B->rowval[p] + (int64_t)j * n] * B->nzval[p];
                trace += mat_temp * C[i + (int64_t)j * n];
            }
    }
    return trace;
}

/* evalobjgrad.jl:2567-2619 */
static void adjoint_grad_calc(ws_t *w, const double *vr0, const double *vi05, const double *vr, const double *lr0,
                              const double *lr05, const double *li, const double *li0, double t0, double dt, double *grad_step) {
    (void)lr0;
    int Npar = w->Npar;
    double *gr = w->gr, *gi = w->gi;
    memset(grad_step, 0, sizeof(double) * Npar);
    for (int q = 0; q < w->Nc; q++) {
        int qs = 2 * q, qa = qs + 1;
        double tt;
        gradbcarrier2(t0, w, qs, gr);
        gradbcarrier2(t0, w, qa, gi);
        tt = adjoint_trace(w, vr0, q, 1, lr05);
        axpy(Npar, -tt, gi, grad_step);
        tt = adjoint_trace(w, vi05, q, 0, lr05);
        axpy(Npar, -tt, gr, grad_step);
        gradbcarrier2(t0 + dt, w, qs, gr);
        gradbcarrier2(t0 + dt, w, qa, gi);
        axpy(Npar, -tt, gr, grad_step);
        tt = adjoint_trace(w, vr, q, 1, lr05);
        axpy(Npar, -tt, gi, grad_step);
        gradbcarrier2(t0 + 0.5 * dt, w, qs, gr);
        gradbcarrier2(t0 + 0.5 * dt, w, qa, gi);
        tt = adjoint_trace(w, vr, q, 0, li);
        axpy(Npar, tt, gr, grad_step);
        tt = adjoint_trace(w, vr0, q, 0, li0);
        axpy(Npar, tt, gr, grad_step);
        tt = adjoint_trace(w, vi05, q, 1, li);
        axpy(Npar, -tt, gi, grad_step);
        tt = adjoint_trace(w, vi05, q, 1, li0);
        axpy(Npar, -tt, gi, grad_step);
    }
    for (int q = 0; q < w->Nunc; q++) {
        /* trace combinations of evalobjgrad.jl:2621-2654: a K-type operator (isSymm) has coefficient c0 at t0 and t0+dt and c1
         * at t0+dt/2; an S-type one has three different ones */
        double c[3];   /* multiplies grad ft at t0, t0+dt, t0+dt/2 */
        if (w->isSymm[q]) {
            double tmp = adjoint_trace(w, vi05, q, 2, lr05);
            c[0] = -tmp; c[1] = -tmp;
            c[2] = adjoint_trace(w, vr, q, 2, li) + adjoint_trace(w, vr0, q, 2, li0);
        } else {
            c[0] = -adjoint_trace(w, vr0, q, 2, lr05);
            c[1] = -adjoint_trace(w, vr, q, 2, lr05);
            c[2] = -(adjoint_trace(w, vi05, q, 2, li) + adjoint_trace(w, vi05, q, 2, li0));
        }
        const double tp[3] = {t0, t0 + dt, t0 + 0.5 * dt};
        if (w->unc_grad_literal) {                     /* the reference's lines as written: one spline, index 2Nc-1+q (1-based q) */
            int qu = 2 * w->Nc - 1 + (q + 1);
            for (int k = 0; k < 3; k++) { gradbcarrier2(tp[k], w, qu, gr); axpy(Npar, c[k], gr, grad_step); }
        } else {                                       /* exact gradient of KS!'s ft = 2 (p cos - q sin) */
            int qs = 2 * w->Nc + 2 * q, qa = qs + 1;
            for (int k = 0; k < 3; k++) {
                double cr = cos(2 * M_PI * w->Rfreq[q] * tp[k]), sr = sin(2 * M_PI * w->Rfreq[q] * tp[k]);
                gradbcarrier2(tp[k], w, qs, gr);
                gradbcarrier2(tp[k], w, qa, gi);
                axpy(Npar, 2.0 * cr * c[k], gr, grad_step);
                axpy(Npar, -2.0 * sr * c[k], gi, grad_step);
            }
        }
    }
}

/* ------------------------------------------------------------------ sparse set-up */
static void csc_alloc(csc_t *A, int n, int64_t nnz) {
    A->n = n; A->nnz = nnz;
    A->colptr = (int64_t *)calloc(n + 1, sizeof(int64_t));
    A->rowval = (int64_t *)calloc(nnz > 0 ? nnz : 1, sizeof(int64_t));
    A->nzval = (double *)calloc(nnz > 0 ? nnz : 1, sizeof(double));
}
static void csc_free(csc_t *A) { free(A->colptr); free(A->rowval); free(A->nzval); }
static void csc_copy_pattern(csc_t *dst, const csc_t *src) {
    csc_alloc(dst, src->n, src->nnz);
    memcpy(dst->colptr, src->colptr, sizeof(int64_t) * (src->n + 1));
    memcpy(dst->rowval, src->rowval, sizeof(int64_t) * src->nnz);
}
/* union pattern of `cnt` CSC matrices (+ optional diagonal), rows sorted within each column */
static void csc_union(csc_t *out, int n, const csc_t **ops, int cnt, int with_diag) {
    char *mark = (char *)calloc((size_t)n * n, 1);
    for (int o = 0; o < cnt; o++)
        for (int c = 0; c < n; c++)
            for (int64_t p = ops[o]->colptr[c]; p < ops[o]->colptr[c + 1]; p++) mark[ops[o]->rowval[p] + (size_t)c * n] = 1;
    if (with_diag) for (int i = 0; i < n; i++) mark[i + (size_t)i * n] = 1;
    int64_t nnz = 0;
    for (size_t i = 0; i < (size_t)n * n; i++) nnz += mark[i];
    csc_alloc(out, n, nnz);
    int64_t p = 0;
    for (int c = 0; c < n; c++) {
        out->colptr[c] = p;
        for (int r = 0; r < n; r++) if (mark[r + (size_t)c * n]) out->rowval[p++] = r;
    }
    out->colptr[n] = p;
    free(mark);
}
static int64_t *csc_map(const csc_t *pat, const csc_t *op) {
    int64_t *map = (int64_t *)malloc(sizeof(int64_t) * (op->nnz > 0 ? op->nnz : 1));
    for (int c = 0; c < op->n; c++)
        for (int64_t p = op->colptr[c]; p < op->colptr[c + 1]; p++) {
            int64_t q = pat->colptr[c];
            while (pat->rowval[q] != op->rowval[p]) q++;
            map[p] = q;
        }
    return map;
}

/* ------------------------------------------------------------------ public entry points */
typedef struct {
    int n, m, Nc, Nfreq, J, objFuncType, sparse;
    int64_t nsteps;
    double T;
    const double *Uinit, *Vtr, *Vti, *wdiag, *Cfreq;
    /* dense: H0 n*n, Hsym Nc*n*n, Hanti Nc*n*n.  sparse: (1+2Nc) CSC operators, order H0, Hsym.., Hanti..,
       colptr concatenated ((1+2Nc)*(n+1)), rowval/nzval concatenated in the same order */
    const double *H0, *Hsym, *Hanti;
    const int64_t *colptr, *rowval;
    const double *nzval;
    int solver;   /* 0/1 Neumann, 2 Jacobi */
    double tol;   /* Jacobi tolerance (already scaled by sqrt(nrhs), linear_solvers.jl:40) */
    /* --- extensions; all-zero = the golden-pinned core --- */
    int pFidType;            /* 0 or 2: pFidType 2; 1, 3, 4 as in evalobjgrad.jl:755-763.  pFidType 3: every pcof vector carries the global
                                phase as an extra LAST entry (stride Npar + 1) and every gradient an extra last entry (:589-596,:923-945) */
    double globalPhase;      /* params.globalPhase (pFidType 1 and 4) */
    const double *wmat_real, *wmat_imag;   /* dense n x n col-major weights, or NULL (Diagonal(wdiag), no imaginary part) */
    int Nunc;                /* uncoupled controls: Cfreq then has Nc + Nunc rows, pcof 2 (Nc + Nunc) Nfreq D1 entries */
    const double *Hunc;      /* dense: Nunc matrices; sparse: the CSC list continues with Hunc.. after Hanti.. */
    const int *isSymm;       /* Nunc */
    const double *Rfreq;     /* Nunc */
    int unc_grad_literal;
} jqo_problem;

static void ws_init(ws_t *w, const jqo_problem *P, int Npar) {
    memset(w, 0, sizeof(*w));
    int n = P->n, m = P->m, Nc = P->Nc;
    w->n = n; w->m = m; w->Nc = Nc; w->Nfreq = P->Nfreq; w->J = P->J; w->objFuncType = P->objFuncType;
    w->sparse = P->sparse; w->nsteps = P->nsteps; w->T = P->T;
    w->solver = P->solver == 2 ? 2 : 1; w->tol = P->tol;
    w->Uinit = P->Uinit; w->Vtr = P->Vtr; w->Vti = P->Vti; w->wdiag = P->wdiag; w->Cfreq = P->Cfreq;
    w->Npar = Npar;
    w->Nunc = P->Nunc; w->NcT = Nc + P->Nunc; w->isSymm = P->isSymm; w->Rfreq = P->Rfreq; w->unc_grad_literal = P->unc_grad_literal;
    w->pFidType = P->pFidType == 0 ? 2 : P->pFidType; w->globalPhase = P->globalPhase;
    w->wreal = P->wmat_real; w->wimag = P->wmat_imag;
    w->D1 = Npar / (2 * w->NcT * P->Nfreq);
    w->dtknot = P->T / (w->D1 - 2);
    int64_t nn = (int64_t)n * n, len = (int64_t)n * m;
    if (!P->sparse) {
        w->H0d = (double *)malloc(sizeof(double) * nn);
        memcpy(w->H0d, P->H0, sizeof(double) * nn);
        w->Hsymd = P->Hsym; w->Hantid = P->Hanti; w->Huncd = P->Hunc;
        double **mats[6] = {&w->K0d, &w->S0d, &w->K05d, &w->S05d, &w->K1d, &w->S1d};
        for (int i = 0; i < 6; i++) *mats[i] = (double *)calloc(nn, sizeof(double));
    } else {
        int nop = 1 + 2 * Nc + P->Nunc;
        csc_t *ops = (csc_t *)calloc(nop, sizeof(csc_t));
        int64_t off = 0;
        for (int o = 0; o < nop; o++) {
            const int64_t *cp = P->colptr + (int64_t)o * (n + 1);
            csc_alloc(&ops[o], n, cp[n]);
            memcpy(ops[o].colptr, cp, sizeof(int64_t) * (n + 1));
            memcpy(ops[o].rowval, P->rowval + off, sizeof(int64_t) * cp[n]);
            memcpy(ops[o].nzval, P->nzval + off, sizeof(double) * cp[n]);
            off += cp[n];
        }
        /* H0 with a full diagonal so that the noise shift has a slot */
        const csc_t *h0p[1] = {&ops[0]};
        csc_union(&w->H0s, n, h0p, 1, 1);
        int64_t *m0 = csc_map(&w->H0s, &ops[0]);
        for (int64_t p = 0; p < ops[0].nnz; p++) w->H0s.nzval[m0[p]] = ops[0].nzval[p];
        free(m0);
        w->Hsyms = ops + 1; w->Hantis = ops + 1 + Nc; w->Huncs = ops + 1 + 2 * Nc;
        const csc_t **kops = (const csc_t **)malloc(sizeof(csc_t *) * (1 + Nc + P->Nunc));
        const csc_t **sops = (const csc_t **)malloc(sizeof(csc_t *) * (Nc + P->Nunc + 1));
        kops[0] = &w->H0s;
        int nk = 1, nsop = 0;
        for (int q = 0; q < Nc; q++) { kops[nk++] = &w->Hsyms[q]; sops[nsop++] = &w->Hantis[q]; }
        for (int q = 0; q < P->Nunc; q++) { if (P->isSymm[q]) kops[nk++] = &w->Huncs[q]; else sops[nsop++] = &w->Huncs[q]; }
        csc_union(&w->K0s, n, kops, nk, 1);
        csc_union(&w->S0s, n, sops, nsop, 0);
        csc_copy_pattern(&w->K05s, &w->K0s); csc_copy_pattern(&w->K1s, &w->K0s);
        csc_copy_pattern(&w->S05s, &w->S0s); csc_copy_pattern(&w->S1s, &w->S0s);
        w->mapK_h0 = (int64_t **)malloc(sizeof(int64_t *));
        w->mapK_h0[0] = csc_map(&w->K0s, &w->H0s);
        w->mapK_sym = (int64_t **)malloc(sizeof(int64_t *) * Nc);
        w->mapS_anti = (int64_t **)malloc(sizeof(int64_t *) * Nc);
        for (int q = 0; q < Nc; q++) { w->mapK_sym[q] = csc_map(&w->K0s, &w->Hsyms[q]); w->mapS_anti[q] = csc_map(&w->S0s, &w->Hantis[q]); }
        w->map_unc = (int64_t **)malloc(sizeof(int64_t *) * (P->Nunc + 1));
        for (int q = 0; q < P->Nunc; q++) w->map_unc[q] = csc_map(P->isSymm[q] ? &w->K0s : &w->S0s, &w->Huncs[q]);
        free(kops); free(sops);
    }
    double **blocks[] = {&w->lambdar, &w->lambdar0, &w->lambdai, &w->lambdai0, &w->lambdar05, &w->lambdar_n, &w->lambdar0_n,
                         &w->lambdai_n, &w->lambdai0_n, &w->lambdar05_n, &w->k1, &w->k2, &w->l1, &w->l2, &w->rhs, &w->hr0,
                         &w->hi0, &w->hr1, &w->hi1, &w->vr, &w->vi, &w->vi05, &w->vr0};
    for (size_t i = 0; i < sizeof(blocks) / sizeof(blocks[0]); i++) *blocks[i] = (double *)calloc(len, sizeof(double));
    double **vecs[] = {&w->gr, &w->gi, &w->gradobjfadj, &w->tr_adj, &w->infidelgrad};
    for (size_t i = 0; i < 5; i++) *vecs[i] = (double *)calloc(Npar, sizeof(double));
}

static void ws_free(ws_t *w) {
    int Nc = w->Nc;
    free(w->jac_scaled); free(w->jac_scaled_s.nzval);
    if (!w->sparse) {
        free(w->H0d); free(w->K0d); free(w->S0d); free(w->K05d); free(w->S05d); free(w->K1d); free(w->S1d);
    } else {
        csc_t *ops = w->Hsyms - 1;
        for (int q = 0; q < Nc; q++) { free(w->mapK_sym[q]); free(w->mapS_anti[q]); }
        for (int q = 0; q < w->Nunc; q++) free(w->map_unc[q]);
        free(w->mapK_h0[0]); free(w->mapK_h0); free(w->mapK_sym); free(w->mapS_anti); free(w->map_unc);
        for (int o = 0; o < 1 + 2 * Nc + w->Nunc; o++) csc_free(&ops[o]);
        free(ops);
        csc_free(&w->H0s); csc_free(&w->K0s); csc_free(&w->S0s); csc_free(&w->K05s); csc_free(&w->S05s); csc_free(&w->K1s); csc_free(&w->S1s);
    }
    double *blocks[] = {w->lambdar, w->lambdar0, w->lambdai, w->lambdai0, w->lambdar05, w->lambdar_n, w->lambdar0_n, w->lambdai_n,
                        w->lambdai0_n, w->lambdar05_n, w->k1, w->k2, w->l1, w->l2, w->rhs, w->hr0, w->hi0, w->hr1, w->hi1, w->vr,
                        w->vi, w->vi05, w->vr0, w->gr, w->gi, w->gradobjfadj, w->tr_adj, w->infidelgrad};
    for (size_t i = 0; i < sizeof(blocks) / sizeof(blocks[0]); i++) free(blocks[i]);
}

/* add (sign * shift) to the diagonal of H0: ipopt_interface.jl:41-44 and :62-64 */
static void shift_h0(ws_t *w, const double *shift, double sign) {
    if (!shift) return;
    int n = w->n;
    if (!w->sparse) { for (int i = 0; i < n; i++) w->H0d[i + (int64_t)i * n] += sign * shift[i]; }
    else
        for (int c = 0; c < n; c++)
            for (int64_t p = w->H0s.colptr[c]; p < w->H0s.colptr[c + 1]; p++)
                if (w->H0s.rowval[p] == c) w->H0s.nzval[p] += sign * shift[c];
}

/* one traceobjgrad(pcof, params, wa, false, evaladjoint): evalobjgrad.jl:504-1038.
 * out[0..3] = objfv, primaryobjf (infidelity), secondaryobjf (leak), traceInfidelity
 * grad = totalgrad; infidelgrad / leakgrad only written when objFuncType != 1 (else infidelgrad aliases totalgrad). */
static void traceobjgrad(ws_t *w, const double *pcof, int evaladjoint, double *out, double *grad, double *infidelgrad_out,
                         double *leakgrad_out) {
    int n = w->n, m = w->m, Npar = w->Npar;
    const int pFid = w->pFidType;
    const double phase = pFid == 3 ? pcof[Npar] : w->globalPhase;   /* evalobjgrad.jl:591-596: the last entry is the global phase */
    int64_t len = (int64_t)n * m, nsteps = w->nsteps;
    double T = w->T, tinv = 1.0 / T;
    w->pcof = pcof;
    double dt = T / nsteps;
    double *vr = w->vr, *vi = w->vi, *vi05 = w->vi05, *vr0 = w->vr0;
    memcpy(vr, w->Uinit, sizeof(double) * len);
    memset(vi, 0, sizeof(double) * len);
    memset(vi05, 0, sizeof(double) * len);
    memset(vr0, 0, sizeof(double) * len);
    double t = 0.0, objfv = 0.0;
    for (int64_t step = 1; step <= nsteps; step++) {
        double forbidden0 = tinv * penalf2aTrap(w, vr);
        memcpy(vr0, vr, sizeof(double) * len);
        KS(w, 0, t);
        KS(w, 1, t + 0.5 * dt);
        KS(w, 2, t + dt);
        t = step_state(w, t, vr, vi, vi05, dt);
        double forbidden = tinv * penalf2a(w, vr, vi05);
        double forbidden_imag1 = tinv * penalf2imag(w, vr0, vi05); /* 0 with a Diagonal wmat_imag, evalobjgrad.jl:2226-2233 */
        objfv = objfv + dt * 0.5 * (forbidden0 + forbidden - 2.0 * forbidden_imag1);
    }
    double primaryobjf, s_re, s_im;
    tracefidcomplex(w, vr, vi, &s_re, &s_im);
    const double cph = cos(phase), sph = sin(phase);
    if (pFid == 1)          /* 1 + |s|^2 - 2 Re(s exp(-i phase)), evalobjgrad.jl:755-757 */
        primaryobjf = 1.0 + tracefidabs2(w, vr, vi) - 2.0 * (s_re * cph + s_im * sph);
    else if (pFid == 2)
        primaryobjf = 1.0 - tracefidabs2(w, vr, vi);
    else {                  /* 1 - tracefidreal(vr, -vi, Re rot, Im rot), rot = exp(i phase) (Vtr + i Vti), :760-762 */
        double tr = 0.0;
        for (int64_t i = 0; i < len; i++) {
            double rr = cph * w->Vtr[i] - sph * w->Vti[i], ri = sph * w->Vtr[i] + cph * w->Vti[i];
            tr += vr[i] * rr + (-vi[i]) * ri;
        }
        primaryobjf = 1.0 - tr / m;
    }
    double secondaryobjf = objfv;
    objfv = primaryobjf + secondaryobjf;
    double traceInfidelity = 1.0 - tracefidabs2(w, vr, vi);
    out[0] = objfv; out[1] = primaryobjf; out[2] = secondaryobjf; out[3] = traceInfidelity;
    if (!evaladjoint) return;

    double *lr = w->lambdar, *lr0 = w->lambdar0, *li = w->lambdai, *li0 = w->lambdai0, *lr05 = w->lambdar05;
    memset(w->gradobjfadj, 0, sizeof(double) * Npar);
    t = T;
    dt = -dt;
    double rs, is;
    tracefidcomplex(w, vr, vi, &rs, &is);
    if (pFid == 1) { rs = cph - rs; is = sph - is; }     /* scomplex0 = exp(i phase) - scomplex0, evalobjgrad.jl:825-826 */
    double phasegrad = 0.0;
    if (pFid == 1 || pFid == 2) {
        for (int64_t i = 0; i < len; i++) { /* init_adjoint!, pFidType == 2 branch (:2029-2042); also the intended one for pFidType 1 */
            double rtmp = (rs * w->Vtr[i] + is * w->Vti[i]) / m;
            lr[i] = rtmp; lr0[i] = rtmp; lr05[i] = rtmp;
            double itmp = (is * w->Vtr[i] - rs * w->Vti[i]) / m;
            li[i] = itmp; li0[i] = itmp;
        }
    } else {
        for (int64_t i = 0; i < len; i++) { /* init_adjoint!, pFidType 3 / 4 (:2043-2057) */
            double rr = cph * w->Vtr[i] - sph * w->Vti[i], ri = sph * w->Vtr[i] + cph * w->Vti[i];
            double rtmp = 0.5 * rr / m, itmp = -0.5 * ri / m;
            lr[i] = rtmp; lr0[i] = rtmp; lr05[i] = rtmp;
            li[i] = itmp; li0[i] = itmp;
            /* primObjGradPhase = -tracefidreal(vfinalr, vfinali, Re(i rot), Im(i rot)), :923-928; vfinali = -vi */
            phasegrad -= (vr[i] * (-ri) + (-vi[i]) * rr) / m;
        }
    }
    if (w->objFuncType != 1) {
        memcpy(w->lambdar_n, lr, sizeof(double) * len); memcpy(w->lambdar0_n, lr0, sizeof(double) * len);
        memcpy(w->lambdai_n, li, sizeof(double) * len); memcpy(w->lambdai0_n, li0, sizeof(double) * len);
        memcpy(w->lambdar05_n, lr05, sizeof(double) * len);
        memset(w->infidelgrad, 0, sizeof(double) * Npar);
    }
    for (int64_t step = nsteps - 1; step >= 0; step--) {
        if (w->wreal) wmul(w, w->hr0, w->wreal, vr, tinv, 0.0);      /* mul!(hr0, wmat_real, vr, tinv, 0.0), :862 */
        else for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) w->hr0[i + (int64_t)j * n] = tinv * w->wdiag[i] * vr[i + (int64_t)j * n];
        double t0 = t;
        memcpy(vr0, vr, sizeof(double) * len);
        KS(w, 0, t);
        KS(w, 1, t + 0.5 * dt);
        KS(w, 2, t + dt);
        t = step_state(w, t, vr, vi, vi05, dt);
        if (w->wreal) {                                              /* evalobjgrad.jl:882-888 */
            wmul(w, w->hi0, w->wreal, vi05, tinv, 0.0);
            wmul(w, w->hr1, w->wreal, vr, tinv, 0.0);
            if (w->wimag) wmul(w, w->hr1, w->wimag, vi05, tinv, 1.0);
            memcpy(w->hi1, w->hi0, sizeof(double) * len);
            if (w->wimag) wmul(w, w->hi1, w->wimag, vr, -tinv, 1.0);
        } else
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n; i++) {
                int64_t e = i + (int64_t)j * n;
                w->hi0[e] = tinv * w->wdiag[i] * vi05[e];
                w->hr1[e] = tinv * w->wdiag[i] * vr[e];
                w->hi1[e] = w->hi0[e]; /* wmat_imag == 0 */
            }
        step_adjoint(w, lr, li, lr05, dt, w->hr0, w->hi0, w->hr1, w->hi1);
        adjoint_grad_calc(w, vr0, vi05, vr, lr0, lr05, li, li0, t0, dt, w->tr_adj);
        axpy(Npar, dt, w->tr_adj, w->gradobjfadj);
        memcpy(li0, li, sizeof(double) * len);
        memcpy(lr0, lr, sizeof(double) * len);
        if (w->objFuncType != 1) {
            step_adjoint(w, w->lambdar_n, w->lambdai_n, w->lambdar05_n, dt, NULL, NULL, NULL, NULL);
            adjoint_grad_calc(w, vr0, vi05, vr, w->lambdar0_n, w->lambdar05_n, w->lambdai_n, w->lambdai0_n, t0, dt, w->tr_adj);
            axpy(Npar, dt, w->tr_adj, w->infidelgrad);
            memcpy(w->lambdai0_n, w->lambdai_n, sizeof(double) * len);
            memcpy(w->lambdar0_n, w->lambdar_n, sizeof(double) * len);
        }
    }
    memcpy(grad, w->gradobjfadj, sizeof(double) * Npar);
    if (pFid == 3) grad[Npar] = phasegrad;                       /* totalgrad[Psize+1] = primObjGradPhase, :931-934 */
    if (w->objFuncType != 1) {
        if (infidelgrad_out) { memcpy(infidelgrad_out, w->infidelgrad, sizeof(double) * Npar); if (pFid == 3) infidelgrad_out[Npar] = phasegrad; }
        if (leakgrad_out) { for (int i = 0; i < Npar; i++) leakgrad_out[i] = w->gradobjfadj[i] - w->infidelgrad[i]; if (pFid == 3) leakgrad_out[Npar] = 0.0; }
    } else if (infidelgrad_out) {
        memcpy(infidelgrad_out, grad, sizeof(double) * (Npar + (pFid == 3))); /* infidelgrad = totalgrad, :951 */
    }
}

/* Batch driver: trajectory (b, s) = candidate pcof[b] under noise sample shift[s]; outputs indexed b*nsamples + s.
 * out: [ntraj][4]; grad/infidelgrad/leakgrad: [ntraj][Npar] (infidelgrad/leakgrad may be NULL).
 * nthreads <= 1 runs serially (the reference is single-threaded); otherwise a pthread pool pulls trajectories
 * from a shared atomic counter (OpenMP's libgomp spec file is not usable with this image's gcc wrapper). */
typedef struct {
    const jqo_problem *P;
    int Npar, nsamples, evaladjoint;
    int64_t ntraj;
    const double *pcof, *shift;
    double *out, *grad, *infidelgrad, *leakgrad;
    atomic_llong *next;
} job_t;

static void *worker(void *arg) {
    job_t *jb = (job_t *)arg;
    const jqo_problem *P = jb->P;
    int Npar = jb->Npar;
    ws_t w;
    ws_init(&w, P, Npar);
    const int ext = P->pFidType == 3 ? 1 : 0;           /* pcof / gradient stride Npar + 1: the global phase rides along */
    double *gtmp = (double *)calloc(Npar + 1, sizeof(double));
    for (;;) {
        int64_t tr = (int64_t)atomic_fetch_add(jb->next, 1);
        if (tr >= jb->ntraj) break;
        int64_t b = tr / jb->nsamples, s = tr % jb->nsamples;
        const double *sh = jb->shift ? jb->shift + s * P->n : NULL;
        shift_h0(&w, sh, +1.0);
        traceobjgrad(&w, jb->pcof + b * (Npar + ext), jb->evaladjoint, jb->out + tr * 4, jb->grad ? jb->grad + tr * (Npar + ext) : gtmp,
                     jb->infidelgrad ? jb->infidelgrad + tr * (Npar + ext) : NULL, jb->leakgrad ? jb->leakgrad + tr * (Npar + ext) : NULL);
        shift_h0(&w, sh, -1.0);
    }
    free(gtmp);
    ws_free(&w);
    return NULL;
}

int jqo_traceobjgrad_batch(const jqo_problem *P, int Npar, int nbatch, const double *pcof, int nsamples, const double *shift,
                           int evaladjoint, int nthreads, double *out, double *grad, double *infidelgrad, double *leakgrad) {
    const int NcT = P->Nc + P->Nunc;
    if (NcT < 1 || Npar % (2 * NcT * P->Nfreq) != 0 || Npar < 3 * 2 * NcT) return -1; /* evalobjgrad.jl:604-606 */
    if (nsamples < 1) nsamples = 1;
    atomic_llong next = 0;
    job_t jb = {P, Npar, nsamples, evaladjoint, (int64_t)nbatch * nsamples, pcof, shift, out, grad, infidelgrad, leakgrad, &next};
    if (nthreads <= 1) { worker(&jb); return 0; }
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int i = 0; i < nthreads; i++) if (pthread_create(&th[started], NULL, worker, &jb) == 0) started++;
    if (started == 0) worker(&jb);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

/* Forward sweep with state history (eval_forward with saveEndOnly=false, src/evalobjgrad.jl:2727-2873, and the
 * verbose=true history of traceobjgrad, :676-680,:748-752): hist_r/hist_i [nsteps/save_every + 1][n*m] = vr, -vi. */
int jqo_forward_history(const jqo_problem *P, int Npar, const double *pcof, const double *shift, int save_every,
                        double *hist_r, double *hist_i, double *out) {
    if (P->Nc + P->Nunc < 1 || Npar % (2 * (P->Nc + P->Nunc) * P->Nfreq) != 0 || Npar < 3 * 2 * (P->Nc + P->Nunc)) return -1;
    if (save_every < 1 || P->nsteps % save_every != 0) return -2;
    ws_t w;
    ws_init(&w, P, Npar);
    shift_h0(&w, shift, +1.0);
    w.pcof = pcof;
    int64_t len = (int64_t)w.n * w.m;
    double dt = w.T / w.nsteps, t = 0.0, tinv = 1.0 / w.T, objfv = 0.0;
    memcpy(w.vr, w.Uinit, sizeof(double) * len);
    memset(w.vi, 0, sizeof(double) * len);
    for (int64_t i = 0; i < len; i++) { hist_r[i] = w.vr[i]; hist_i[i] = -w.vi[i]; }
    for (int64_t step = 1; step <= w.nsteps; step++) {
        double forbidden0 = tinv * penalf2aTrap(&w, w.vr);
        KS(&w, 0, t);
        KS(&w, 1, t + 0.5 * dt);
        KS(&w, 2, t + dt);
        t = step_state(&w, t, w.vr, w.vi, w.vi05, dt);
        objfv += dt * 0.5 * (forbidden0 + tinv * penalf2a(&w, w.vr, w.vi05));
        if (step % save_every == 0) {
            double *hr = hist_r + (step / save_every) * len, *hi = hist_i + (step / save_every) * len;
            for (int64_t i = 0; i < len; i++) { hr[i] = w.vr[i]; hi[i] = -w.vi[i]; }
        }
    }
    out[0] = 1.0 - tracefidabs2(&w, w.vr, w.vi);
    out[1] = objfv;
    ws_free(&w);
    return 0;
}

/* evalctrl (plotstatectrl.jl:246-276): p_q(t), q_q(t) for every coupled control at the given times.
 * p, q: [Nc][ntimes].  Only the fields of the problem that the control functions read are used. */
int jqo_eval_controls(const jqo_problem *P, int Npar, const double *pcof, int ntimes, const double *times, double *p, double *q) {
    const int NcT = P->Nc + P->Nunc;
    if (Npar % (2 * NcT * P->Nfreq) != 0 || Npar / (2 * NcT * P->Nfreq) < 3) return -2;
    ws_t w;
    memset(&w, 0, sizeof(w));
    w.Nc = P->Nc; w.NcT = NcT; w.Nfreq = P->Nfreq; w.Cfreq = P->Cfreq; w.T = P->T; w.Npar = Npar; w.pcof = pcof;
    w.D1 = Npar / (2 * NcT * P->Nfreq);
    w.dtknot = P->T / (w.D1 - 2);
    for (int c = 0; c < NcT; c++)
        for (int i = 0; i < ntimes; i++) {
            p[(int64_t)c * ntimes + i] = bcarrier2(times[i], &w, 2 * c);
            q[(int64_t)c * ntimes + i] = bcarrier2(times[i], &w, 2 * c + 1);
        }
    return 0;
}

int jqo_max_threads(void) {
    long nproc = sysconf(_SC_NPROCESSORS_ONLN);
    return nproc > 0 ? (int)nproc : 1;
}
