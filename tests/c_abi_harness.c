/* C-language harness for the drop-in boundary (include/juqbox_b200.h), built with plain gcc — no CUDA, no Python in the
 * process.  It stands in for the Julia `ccall` side, which cannot run in this image (no julia):
 *
 *   1. compile time: _Static_asserts pin sizeof/offsetof of jq_operator and jq_problem to the layout that the Julia
 *      `struct JqOperator` / `struct JqProblem` of julia/JuqboxB200.jl produce (Julia lays out isbits structs like C:
 *      Int32 x 8, Int64, Float64, 5 Ptr, the 40-byte operator inline, 2 Ptr, Float64, then the 8f-rank-3 tail);
 *   2. `c_abi_harness --layout`: prints those numbers so that tests can compare them with ctypes and with jq_abi_info();
 *   3. `c_abi_harness <libjuqbox_b200.so> <problem.bin>`: loads the library with dlopen, reads a problem the way Julia holds it
 *      (column-major dense matrices), converts the operators to SparseMatrixCSC with 1-BASED indices, shifts them to 0-based
 *      exactly as jq_operator(::SparseMatrixCSC) in the shim does, and calls jq_create / jq_traceobjgrad_batch /
 *      jq_eval_f_grad (twice: evaluation + cache hit) / jq_destroy through the C ABI.  Prints the results with %.17g.
 *
 * Reference signature served: traceobjgrad(pcof0, params, wa, false, true) (src/evalobjgrad.jl:504,1032-1035) and
 * eval_f_par / eval_grad_f_par (src/ipopt_interface.jl:77-148).
 */
#include <dlfcn.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/juqbox_b200.h"

/* Julia: struct JqOperator; format::Int32; nnz::Int64; colptr::Ptr{Int64}; rowval::Ptr{Int64}; nzval::Ptr{Float64}; end */
_Static_assert(sizeof(jq_operator) == 40, "jq_operator must be 40 bytes (Int32 + pad, Int64, 3 Ptr)");
_Static_assert(offsetof(jq_operator, format) == 0 && offsetof(jq_operator, nnz) == 8 && offsetof(jq_operator, colptr) == 16 &&
               offsetof(jq_operator, rowval) == 24 && offsetof(jq_operator, nzval) == 32, "jq_operator field offsets");
/* Julia: struct JqProblem: 8 x Int32, nsteps::Int64, T::Float64, 5 x Ptr, h0::JqOperator, hsym::Ptr, hanti::Ptr, solver_tol::Float64,
 *        global_phase::Float64, wmat_real::Ptr, wmat_imag::Ptr, nuncoupled::Int32, reserved0::Int32, hunc::Ptr, unc_is_symm::Ptr, unc_rfreq::Ptr */
_Static_assert(offsetof(jq_problem, n) == 0 && offsetof(jq_problem, linear_solver) == 28, "eight Int32 first");
_Static_assert(offsetof(jq_problem, nsteps) == 32 && offsetof(jq_problem, T) == 40, "nsteps, T");
_Static_assert(offsetof(jq_problem, uinit) == 48 && offsetof(jq_problem, cfreq) == 80, "five pointers");
_Static_assert(offsetof(jq_problem, h0) == 88 && offsetof(jq_problem, hsym) == 128 && offsetof(jq_problem, hanti) == 136, "operators");
_Static_assert(offsetof(jq_problem, solver_tol) == 144 && offsetof(jq_problem, global_phase) == 152, "solver_tol, global_phase");
_Static_assert(offsetof(jq_problem, wmat_real) == 160 && offsetof(jq_problem, wmat_imag) == 168 && offsetof(jq_problem, nuncoupled) == 176 &&
               offsetof(jq_problem, hunc) == 184 && offsetof(jq_problem, unc_is_symm) == 192 && offsetof(jq_problem, unc_rfreq) == 200, "rank-3 tail");
_Static_assert(sizeof(jq_problem) == 208, "jq_problem must be 208 bytes");

typedef struct { int64_t *colptr, *rowval; double *nzval; int64_t nnz; } csc1;

/* sparse(A) of Julia: column-major scan, exact zeros dropped, indices 1-based */
static csc1 sparse_1based(const double *A, int n) {
    csc1 s;
    s.colptr = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    s.rowval = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n * n + 1));
    s.nzval = (double *)malloc(sizeof(double) * ((size_t)n * n + 1));
    s.nnz = 0;
    for (int c = 0; c < n; c++) {
        s.colptr[c] = s.nnz + 1;
        for (int r = 0; r < n; r++)
            if (A[r + (size_t)c * n] != 0.0) { s.rowval[s.nnz] = r + 1; s.nzval[s.nnz] = A[r + (size_t)c * n]; s.nnz++; }
    }
    s.colptr[n] = s.nnz + 1;
    return s;
}
/* jq_operator(A::SparseMatrixCSC, keep) of julia/JuqboxB200.jl: cp = A.colptr .- 1; rv = A.rowval .- 1 */
static jq_operator shim_operator(const double *A, int n, int sparse) {
    jq_operator op;
    memset(&op, 0, sizeof(op));
    if (!sparse) { op.format = JQ_DENSE; op.nnz = (int64_t)n * n; op.nzval = A; return op; }
    csc1 s = sparse_1based(A, n);
    for (int c = 0; c <= n; c++) s.colptr[c] -= 1;
    for (int64_t k = 0; k < s.nnz; k++) s.rowval[k] -= 1;
    op.format = JQ_CSC; op.nnz = s.nnz; op.colptr = s.colptr; op.rowval = s.rowval; op.nzval = s.nzval;
    return op;
}

static double *rd(FILE *f, size_t cnt) {
    double *p = (double *)malloc(sizeof(double) * (cnt ? cnt : 1));
    if (fread(p, sizeof(double), cnt, f) != cnt) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char **argv) {
    if (argc == 2 && strcmp(argv[1], "--layout") == 0) {
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(jq_problem), sizeof(jq_operator), offsetof(jq_problem, nsteps),
               offsetof(jq_problem, T), offsetof(jq_problem, uinit), offsetof(jq_problem, h0), offsetof(jq_problem, hsym),
               offsetof(jq_problem, solver_tol), offsetof(jq_operator, nnz), offsetof(jq_operator, nzval));
        return 0;
    }
    if (argc < 3) { fprintf(stderr, "usage: %s --layout | <lib.so> <problem.bin>\n", argv[0]); return 2; }
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 3; }
    SYM(jq_create) SYM(jq_destroy) SYM(jq_traceobjgrad_batch) SYM(jq_eval_f_grad) SYM(jq_last_error) SYM(jq_abi_info)
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror("problem file"); return 2; }
    /* header: n m Nc Nfreq J objFuncType sparse npar (int64 each), nsteps, then T, tik0 (double) */
    int64_t hd[9];
    if (fread(hd, sizeof(int64_t), 9, f) != 9) return 2;
    const int n = (int)hd[0], m = (int)hd[1], Nc = (int)hd[2], Nfreq = (int)hd[3], sparse = (int)hd[6], npar = (int)hd[7];
    double *sc = rd(f, 2);
    const double T = sc[0], tik0 = sc[1];
    double *uinit = rd(f, (size_t)n * m), *vtr = rd(f, (size_t)n * m), *vti = rd(f, (size_t)n * m), *wdiag = rd(f, n);
    double *cfreq = rd(f, (size_t)Nc * Nfreq), *H0 = rd(f, (size_t)n * n);
    jq_operator *hs = (jq_operator *)calloc(Nc, sizeof(jq_operator)), *ha = (jq_operator *)calloc(Nc, sizeof(jq_operator));
    for (int q = 0; q < Nc; q++) hs[q] = shim_operator(rd(f, (size_t)n * n), n, sparse);
    for (int q = 0; q < Nc; q++) ha[q] = shim_operator(rd(f, (size_t)n * n), n, sparse);
    double *pcof = rd(f, npar);
    fclose(f);

    jq_problem pb;
    memset(&pb, 0, sizeof(pb));
    pb.n = n; pb.m = m; pb.ncoupled = Nc; pb.nfreq = Nfreq; pb.neumann_terms = (int32_t)hd[4]; pb.obj_func_type = (int32_t)hd[5];
    pb.pfid_type = 2; pb.linear_solver = 1; pb.nsteps = hd[8]; pb.T = T;
    pb.uinit = uinit; pb.vtarget_r = vtr; pb.vtarget_i = vti; pb.wdiag = wdiag; pb.cfreq = cfreq;
    pb.h0 = shim_operator(H0, n, sparse); pb.hsym = hs; pb.hanti = ha;
    if (p_jq_abi_info(1) != (int64_t)sizeof(jq_problem) || p_jq_abi_info(2) != (int64_t)sizeof(jq_operator)) { fprintf(stderr, "ABI size mismatch\n"); return 4; }
    jq_handle *h = NULL;
    if (p_jq_create(&pb, 0, &h) != 0) { fprintf(stderr, "jq_create: %s\n", p_jq_last_error()); return 5; }
    double infid, leak, tinf, *grad = (double *)calloc(npar, sizeof(double));
    /* traceobjgrad(pcof, params, wa, false, true): nbatch = 1, no noise sample */
    if (p_jq_traceobjgrad_batch(h, 1, pcof, npar, 1, NULL, NULL, 1, &infid, &leak, &tinf, grad, NULL, NULL) != 0) {
        fprintf(stderr, "jq_traceobjgrad_batch: %s\n", p_jq_last_error()); return 6;
    }
    printf("traceobjgrad %.17g %.17g %.17g\n", infid, leak, tinf);
    for (int k = 0; k < npar; k++) printf("%.17g%c", grad[k], k + 1 < npar ? ' ' : '\n');
    /* eval_f_par + eval_grad_f_par as one fused call, then the cache hit of the second callback */
    double fval, *gf = (double *)calloc(npar, sizeof(double));
    int32_t evaluated = -1;
    for (int rep = 0; rep < 2; rep++) {
        if (p_jq_eval_f_grad(h, pcof, npar, 1, NULL, NULL, tik0, NULL, &fval, gf, &infid, &leak, NULL, &evaluated) != 0) {
            fprintf(stderr, "jq_eval_f_grad: %s\n", p_jq_last_error()); return 7;
        }
        printf("eval_f_grad %d %.17g\n", (int)evaluated, fval);
        for (int k = 0; k < npar; k++) printf("%.17g%c", gf[k], k + 1 < npar ? ' ' : '\n');
    }
    /* error path: wrong pcof length must come back as JQ_ERR_PCOF_LENGTH with a message, like error() at evalobjgrad.jl:604-606 */
    int rc = p_jq_traceobjgrad_batch(h, 1, pcof, npar - 1, 1, NULL, NULL, 1, &infid, &leak, &tinf, grad, NULL, NULL);
    printf("badlen %d %s\n", rc, p_jq_last_error());
    p_jq_destroy(h);
    return 0;
}
