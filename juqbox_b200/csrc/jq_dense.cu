// Dense-operator trajectory kernel on the FP64 tensor-core path (kernel id 6).
//
// For problems whose control Hamiltonians are genuinely dense / unstructured (nothing the register-resident planners recognise) the
// operator products ARE real contractions: Y = K(t) X with K n x n dense and X = [U_1 ... U_S] the n x (m S) block of the states of S
// noise samples that share one pcof vector — they see the same K(t), S(t) up to the diagonal shift of Hconst (src/ipopt_interface.jl:
// 41-44), which is added in the tile epilogue.  Every product of the steppers, of the Neumann series and of the gradient traces is
// one `gemm` on FP64 MMA (mma.sync.aligned.m8n8k4.f64: DMMA in SASS); K(t), S(t) are assembled per step in shared memory exactly like the
// reference's KS! (src/evalobjgrad.jl:2354-2389).  One CTA of 8 warps owns one (candidate, tile of ST samples); all n x (m ST)
// blocks live in shared memory (column-major, leading dimension padded so that the MMA fragment loads are bank-conflict free).
//
// Same algorithm and operation order as the generic kernel (jq_generic.cu), which remains the cross-check:
// forward loop src/evalobjgrad.jl:698-753, infidelity :755-766, terminal condition :810-844 / :2026-2059, backward loop :859-921,
// steppers src/StormerVerlet.jl:255-303,:461-504, Neumann src/linear_solvers.jl:94-106, gradient :2567-2619.
// Scope: objFuncType 1, Neumann solver, diagonal guard-level weights, any pFidType, coupled and uncoupled controls, n <= 64.
#include "jq_common.h"

#define DN_THREADS 256
#define DN_MAXST 8

namespace {

struct DenseGeom {
    int n8, n4, ldk, ldx;      // rows padded to 8, contraction length padded to 4, leading dimensions of matrices / blocks
    int ST, N8;                // samples per tile, columns m * ST padded to 8
    int nitems, tiles_per_cand;
    int ops_smem;              // 1: Hconst, Hsym_q, Hanti_q copied to shared memory; 0: read from their padded global copy (large n)
    int nthr;                  // threads per CTA: one warp per 8 x 8 output tile of a product, 2 ... 8 warps
};

struct DCtx {
    const DevProblem *P;
    int n, m, Nc, Nfreq, D1, Npar, J, ns, ncols;   // ns: samples in this tile, ncols = m * ns
    DenseGeom G;
    double *H0, *Hs, *Ha;      // (1 + 2 Nc) dense operators, row-major n8 x ldk
    double *KS;                // K(t), S(t) at the three time levels of a step: K0, S0, K05, S05, K1, S1 (sz doubles each)
    int sz;
    __device__ __forceinline__ double *K(int l) const { return KS + (size_t)(2 * l) * sz; }
    __device__ __forceinline__ double *S(int l) const { return KS + (size_t)(2 * l + 1) * sz; }
    double *vr, *vi, *vi05, *vr0, *lr, *li, *lr05, *li0, *rhs, *scr, *k1, *k2, *l1, *l2;
    double *pcof, *gsm, *ctrl, *shift, *red;      // shift: block-shaped (n x ncols, ldx): diagonal of Hconst's noise shift of the column's sample
    double dtknot, tinv;
};

// src/bsplines.jl:211-304 (0-based indices); pcof in shared memory
__device__ double dn_bcarrier2(const DCtx &c, double t, int func) {
    const int osc = func >> 1, qf = func & 1;
    const double width = 3.0 * c.dtknot;
    long long k = (long long)ceil(t / c.dtknot + 2.0);
    k = k < 3 ? 3 : (k > c.D1 ? c.D1 : k);
    double f = 0.0;
    for (int fr = 0; fr < c.Nfreq; ++fr) {
        const int off1 = 2 * osc * c.Nfreq * c.D1 + fr * 2 * c.D1 - 1, off2 = off1 + c.D1;
        double fbs1 = 0.0, fbs2 = 0.0;
        double tau = (t - c.dtknot * ((double)k - 1.5)) / width;
        double b = 9.0 / 8 + 4.5 * tau + 4.5 * tau * tau;
        fbs1 += c.pcof[off1 + k] * b; fbs2 += c.pcof[off2 + k] * b;
        tau = (t - c.dtknot * ((double)(k - 1) - 1.5)) / width;
        b = 0.75 - 9.0 * tau * tau;
        fbs1 += c.pcof[off1 + k - 1] * b; fbs2 += c.pcof[off2 + k - 1] * b;
        tau = (t - c.dtknot * ((double)(k - 2) - 1.5)) / width;
        b = 9.0 / 8 - 4.5 * tau + 4.5 * tau * tau;
        fbs1 += c.pcof[off1 + k - 2] * b; fbs2 += c.pcof[off2 + k - 2] * b;
        double sn, cs;
        sincos(c.P->cfreq[osc + c.Nc * fr] * t, &sn, &cs);
        f += qf ? fbs1 * sn + fbs2 * cs : fbs1 * cs - fbs2 * sn;
    }
    return f;
}

// control values at the three time levels (uncoupled controls as in jq_generic.cu / KS! :2372-2387), then K(level), S(level)
__device__ __forceinline__ void dn_assemble(const DCtx &c, double t, double dt) {
    const int nf = 2 * c.Nc;
    for (int idx = threadIdx.x; idx < 3 * nf; idx += blockDim.x) {
        const int level = idx / nf, func = idx % nf;
        const double tt = level == 0 ? t : (level == 1 ? t + 0.5 * dt : t + dt);
        const int kind = c.P->ctrl_kind[func >> 1];
        double v;
        if (kind == 0) v = dn_bcarrier2(c, tt, func);
        else {
            v = 0.0;
            if ((func & 1) == (kind == 2)) {
                double sr, cr;
                sincos(2.0 * M_PI * c.P->ctrl_rfreq[func >> 1] * tt, &sr, &cr);
                v = 2.0 * (dn_bcarrier2(c, tt, func & ~1) * cr - dn_bcarrier2(c, tt, func | 1) * sr);
            }
        }
        c.ctrl[idx] = v;
    }
    __syncthreads();
    const int sz = c.G.n8 * c.G.ldk;
    for (int r = threadIdx.x; r < sz; r += blockDim.x) {           // every operator entry is read once for the three levels
        double kk[3], ss[3];
        const double h0 = c.H0[r];
        for (int l = 0; l < 3; ++l) { kk[l] = h0; ss[l] = 0.0; }
        for (int q = 0; q < c.Nc; ++q) {
            const double hs = c.Hs[q * sz + r], ha = c.Ha[q * sz + r];
            for (int l = 0; l < 3; ++l) { kk[l] = fma(c.ctrl[l * nf + 2 * q], hs, kk[l]); ss[l] = fma(c.ctrl[l * nf + 2 * q + 1], ha, ss[l]); }
        }
        for (int l = 0; l < 3; ++l) { c.K(l)[r] = kk[l]; c.S(l)[r] = ss[l]; }
    }
    __syncthreads();
}

// One operator product of a sum: a * A X, A an n8 x n4 row-major matrix (ldk); shift: A is a K(t), i.e. + a * diag(shift_s) X for the
// columns of sample s (the per-sample diagonal of Hconst).  A == nullptr: term absent.
struct DTerm { const double *A; const double *X; double a; bool shift; };

// value(row, col) = sum of up to three products (one accumulator pair, n4 / 4 DMMA per product and 8 x 8 tile, one tile per warp and
// turn), handed to ep(idx, row, col, value) for every live element; ep stores (and may update other blocks at the same position: a tile
// is always owned by the same warp, so consecutive products into one block need no barrier in between).  Ends with a CTA barrier.
template <class EP>
__device__ __forceinline__ void dn_gemm(const DCtx &c, DTerm t0, DTerm t1, DTerm t2, EP ep) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int RT = c.G.n8 >> 3, CT = c.G.N8 >> 3, ldk = c.G.ldk, ldx = c.G.ldx;
    for (int t = warp; t < RT * CT; t += (blockDim.x >> 5)) {
        const int rt = t % RT, ct = t / RT;
        const int row = 8 * rt + (lane >> 2), col = 8 * ct + 2 * (lane & 3);
        const int aoff = (8 * rt + (lane >> 2)) * ldk + (lane & 3), boff = (8 * ct + (lane >> 2)) * ldx + (lane & 3);
        double v0 = 0.0, v1 = 0.0;
        const DTerm terms[3] = {t0, t1, t2};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (!terms[q].A) continue;
            double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;       // two accumulator chains: a DMMA waits 26 cycles for its predecessor
            const double *ap = terms[q].A + aoff, *bp = terms[q].X + boff;
            int kk = 0;
            for (; kk + 4 < c.G.n4; kk += 8) {
                const double a = ap[kk], b = bp[kk], a2 = ap[kk + 4], b2 = bp[kk + 4];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a2), "d"(b2));
            }
            if (kk < c.G.n4) {
                const double a = ap[kk], b = bp[kk];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            }
            c0 += d0; c1 += d1;
            if (terms[q].shift) {            // padding rows / columns of the shift block and of X are zero
                c0 = fma(c.shift[col * ldx + row], terms[q].X[col * ldx + row], c0);
                c1 = fma(c.shift[(col + 1) * ldx + row], terms[q].X[(col + 1) * ldx + row], c1);
            }
            v0 = fma(terms[q].a, c0, v0);
            v1 = fma(terms[q].a, c1, v1);
        }
        if (row < c.n) {
            if (col < c.ncols) ep(col * ldx + row, row, col, v0);
            if (col + 1 < c.ncols) ep((col + 1) * ldx + row, row, col + 1, v1);
        }
    }
    __syncthreads();
}
#define DT_NONE DTerm{nullptr, nullptr, 0.0, false}

// elementwise loop over the live n x ncols part of the blocks (no divisions: a warp per column, lanes over the rows): (i, col, idx)
#define DN_FOR for (int col = threadIdx.x >> 5; col < c.ncols; col += (blockDim.x >> 5)) \
                   for (int i = threadIdx.x & 31, idx = col * c.G.ldx + i; i < c.n; i += 32, idx += 32)

// X = sum_{j<=J} (h/2)^j S^j B ; B destroyed, T scratch (src/linear_solvers.jl:94-106).  X must already hold B (the caller's product
// epilogue writes both); the update X += coeff T rides on the epilogue of the product that makes T.
__device__ __forceinline__ void dn_neumann(const DCtx &c, const double *Smat, double h, double *B, double *T, double *X) {
    double coeff = 1.0;
    for (int it = 0; it < c.J; ++it) {
        coeff *= 0.5 * h;
        dn_gemm(c, DTerm{Smat, B, 1.0, false}, DT_NONE, DT_NONE, [&](int idx, int, int, double v) { T[idx] = v; X[idx] += coeff * v; });
        double *sw = B; B = T; T = sw;
    }
}

// src/StormerVerlet.jl:461-504
__device__ __forceinline__ void dn_state_step(const DCtx &c, double h) {
    double *u = c.vr, *v = c.vi, *v05 = c.vi05;
    // rhs = K05 u + S05 v ; l1 = neumann(S05, rhs) ; v05 = v + h/2 l1
    dn_gemm(c, DTerm{c.K(1), u, 1.0, true}, DTerm{c.S(1), v, 1.0, false}, DT_NONE, [&](int idx, int, int, double val) { c.rhs[idx] = val; c.l1[idx] = val; });
    dn_neumann(c, c.S(1), h, c.rhs, c.scr, c.l1);
    DN_FOR v05[idx] = v[idx] + 0.5 * h * c.l1[idx];
    __syncthreads();
    // k1 = S0 u - K0 v05
    dn_gemm(c, DTerm{c.S(0), u, 1.0, false}, DTerm{c.K(0), v05, -1.0, true}, DT_NONE, [&](int idx, int, int, double val) { c.k1[idx] = val; });
    // rhs = S1 u + h/2 S1 k1 - K1 v05 ; k2 = neumann(S1, rhs) ; u += h/2 k1 (after the products have read u: next barrier)
    dn_gemm(c, DTerm{c.S(2), u, 1.0, false}, DTerm{c.S(2), c.k1, 0.5 * h, false}, DTerm{c.K(2), v05, -1.0, true},
            [&](int idx, int, int, double val) { c.rhs[idx] = val; c.k2[idx] = val; });
    DN_FOR u[idx] += 0.5 * h * c.k1[idx];
    __syncthreads();
    dn_neumann(c, c.S(2), h, c.rhs, c.scr, c.k2);
    DN_FOR u[idx] += 0.5 * h * c.k2[idx];
    __syncthreads();
    // l2 = K05 u + S05 v05 ; v += h/2 (l1 + l2)
    dn_gemm(c, DTerm{c.K(1), u, 1.0, true}, DTerm{c.S(1), v05, 1.0, false}, DT_NONE, [&](int idx, int, int, double val) { v[idx] += 0.5 * h * (c.l1[idx] + val); });
}

// src/StormerVerlet.jl:255-303; forcing with the diagonal weights: hr0 = W vr0 / T, hi0 = hi1 = W vi05 / T, hr1 = W vr / T
__device__ __forceinline__ void dn_adjoint_step(const DCtx &c, double *mu, double *nu, double *X, double h) {
    const double *w = c.P->wdiag;
    // rhs = S0 mu - K05 nu + hr0 ; k2 = neumann(S0, rhs) ; mu += h/2 k2 ; X = mu
    dn_gemm(c, DTerm{c.S(0), mu, 1.0, false}, DTerm{c.K(1), nu, -1.0, true}, DT_NONE, [&](int idx, int row, int, double val) {
        val += c.tinv * w[row] * c.vr0[idx];
        c.rhs[idx] = val; c.k2[idx] = val;
    });
    dn_neumann(c, c.S(0), h, c.rhs, c.scr, c.k2);
    DN_FOR { mu[idx] += 0.5 * h * c.k2[idx]; X[idx] = mu[idx]; }
    __syncthreads();
    // l2 = K0 X + S05 nu + hi0
    dn_gemm(c, DTerm{c.K(0), X, 1.0, true}, DTerm{c.S(1), nu, 1.0, false}, DT_NONE,
            [&](int idx, int row, int, double val) { c.l2[idx] = val + c.tinv * w[row] * c.vi05[idx]; });
    // rhs = S05 nu + h/2 S05 l2 + K1 X + hi1 ; l1 = neumann(S05, rhs) ; nu += h/2 (l2 + l1)
    dn_gemm(c, DTerm{c.S(1), nu, 1.0, false}, DTerm{c.S(1), c.l2, 0.5 * h, false}, DTerm{c.K(2), X, 1.0, true}, [&](int idx, int row, int, double val) {
        val += c.tinv * w[row] * c.vi05[idx];
        c.rhs[idx] = val; c.l1[idx] = val;
    });
    dn_neumann(c, c.S(1), h, c.rhs, c.scr, c.l1);
    DN_FOR nu[idx] += 0.5 * h * (c.l2[idx] + c.l1[idx]);
    __syncthreads();
    // kappa1 = S1 X - K05 nu + hr1 ; mu += h/2 kappa1
    dn_gemm(c, DTerm{c.S(2), X, 1.0, false}, DTerm{c.K(1), nu, -1.0, true}, DT_NONE,
            [&](int idx, int row, int, double val) { mu[idx] += 0.5 * h * (val + c.tinv * w[row] * c.vr[idx]); });
}

// per-sample sums of `cnt` values: warp w takes samples w, w + 8, ...; f(s, e_in_sample, idx, acc)
template <int CNT, class F>
__device__ __forceinline__ void dn_sample_sums(const DCtx &c, double *out /* [ST][CNT] in shared memory */, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s < c.ns; s += (blockDim.x >> 5)) {
        double acc[CNT];
        for (int k = 0; k < CNT; ++k) acc[k] = 0.0;
        for (int j = 0; j < c.m; ++j)
            for (int i = lane; i < c.n; i += 32) f(s, i, j, (s * c.m + j) * c.G.ldx + i, acc);
        for (int k = 0; k < CNT; ++k) {
            double x = acc[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) out[s * CNT + k] = x;
        }
    }
    __syncthreads();
}

// One step's contribution to the gradient of every sample of the tile (src/evalobjgrad.jl:2567-2619).
__device__ __forceinline__ void dn_grad_step(const DCtx &c, double t0, double dt) {
    const int sz = c.G.n8 * c.G.ldk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < c.Nc; ++q) {
        double *aX = c.rhs, *sX = c.scr, *sLi = c.k1, *sLi0 = c.k2, *aLi = c.l1;
        auto put = [&](double *dst) { return [dst](int idx, int, int, double val) { dst[idx] = val; }; };
        dn_gemm(c, DTerm{c.Ha + q * sz, c.lr05, 1.0, false}, DT_NONE, DT_NONE, put(aX));
        dn_gemm(c, DTerm{c.Hs + q * sz, c.lr05, 1.0, false}, DT_NONE, DT_NONE, put(sX));
        dn_gemm(c, DTerm{c.Hs + q * sz, c.li, 1.0, false}, DT_NONE, DT_NONE, put(sLi));
        dn_gemm(c, DTerm{c.Hs + q * sz, c.li0, 1.0, false}, DT_NONE, DT_NONE, put(sLi0));
        dn_gemm(c, DTerm{c.Ha + q * sz, c.li, 1.0, false}, DTerm{c.Ha + q * sz, c.li0, 1.0, false}, DT_NONE, put(aLi));      // Ha (li + li0)
        dn_sample_sums<5>(c, c.red, [&](int, int, int, int idx, double *T) {
            T[0] += c.vr0[idx] * aX[idx];                               // tr(vr0, Ha, lr05)
            T[1] += c.vi05[idx] * sX[idx];                              // tr(vi05, Hs, lr05)
            T[2] += c.vr[idx] * aX[idx];                                // tr(vr, Ha, lr05)
            T[3] += c.vr[idx] * sLi[idx] + c.vr0[idx] * sLi0[idx];      // tr(vr,Hs,li) + tr(vr0,Hs,li0)
            T[4] += c.vi05[idx] * aLi[idx];                             // tr(vi05,Ha,li) + tr(vi05,Ha,li0)
        });
        const int kind = c.P->ctrl_kind[q];
        for (int s = warp; s < c.ns; s += (blockDim.x >> 5)) {
            const double *T = c.red + s * 5;
            double *g = c.gsm + s * c.Npar;
            if (lane < 2 * c.Nfreq) {                                   // role (frequency, alpha), as in jq_generic.cu
                const int fr = lane >> 1, alpha = lane & 1;
                const int base = 2 * q * c.Nfreq * c.D1 + fr * 2 * c.D1 + alpha * c.D1 - 1;
                const double om = c.P->cfreq[q + c.Nc * fr];
                for (int tp = 0; tp < 3; ++tp) {
                    const double tt = tp == 0 ? t0 : (tp == 1 ? t0 + dt : t0 + 0.5 * dt);
                    double Pc = tp == 2 ? T[3] : -T[1];
                    double Qc = tp == 0 ? -T[0] : (tp == 1 ? -T[2] : -T[4]);
                    if (kind != 0) {
                        const double C0 = kind == 1 ? Pc : Qc;
                        double sr, cr;
                        sincos(2.0 * M_PI * c.P->ctrl_rfreq[q] * tt, &sr, &cr);
                        Pc = 2.0 * cr * C0;
                        Qc = -2.0 * sr * C0;
                    }
                    double sn, cs;
                    sincos(om * tt, &sn, &cs);
                    const double Xv = alpha == 0 ? Pc * cs + Qc * sn : Qc * cs - Pc * sn;
                    long long k = (long long)ceil(tt / c.dtknot + 2.0);
                    k = k < 3 ? 3 : (k > c.D1 ? c.D1 : k);
                    const double width = 3.0 * c.dtknot;
                    double tau = (tt - c.dtknot * ((double)k - 1.5)) / width;
                    g[base + k] += Xv * (9.0 / 8 + 4.5 * tau + 4.5 * tau * tau);
                    tau = (tt - c.dtknot * ((double)(k - 1) - 1.5)) / width;
                    g[base + k - 1] += Xv * (0.75 - 9.0 * tau * tau);
                    tau = (tt - c.dtknot * ((double)(k - 2) - 1.5)) / width;
                    g[base + k - 2] += Xv * (9.0 / 8 - 4.5 * tau + 4.5 * tau * tau);
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(DN_THREADS) jq_dense_kernel(DevProblem P, LaunchArgs A, DenseGeom G) {
    extern __shared__ __align__(16) double sm[];
    DCtx c;
    c.P = &P; c.G = G;
    c.n = P.n; c.m = P.m; c.Nc = P.Nc; c.Nfreq = P.Nfreq; c.J = P.J; c.Npar = A.Npar; c.D1 = A.D1;
    c.dtknot = P.T / (A.D1 - 2);
    c.tinv = 1.0 / P.T;
    const int sz = G.n8 * G.ldk, bsz = G.N8 * G.ldx;
    double *p = sm;
    if (G.ops_smem) { c.H0 = p; p += (1 + 2 * c.Nc) * sz; }
    else c.H0 = const_cast<double *>(P.dense_ops);          // same padded layout, through L1 / L2
    c.Hs = c.H0 + sz;
    c.Ha = c.Hs + c.Nc * sz;
    c.KS = p; p += 6 * sz; c.sz = sz;
    c.vr = p; c.vi = p + bsz; c.vi05 = p + 2 * bsz; c.vr0 = p + 3 * bsz; c.lr = p + 4 * bsz; c.li = p + 5 * bsz; c.lr05 = p + 6 * bsz;
    c.li0 = p + 7 * bsz; c.rhs = p + 8 * bsz; c.scr = p + 9 * bsz; c.k1 = p + 10 * bsz; c.k2 = p + 11 * bsz; c.l1 = p + 12 * bsz;
    c.l2 = p + 13 * bsz; p += 14 * bsz;
    c.pcof = p; p += c.Npar;
    c.gsm = p; p += G.ST * c.Npar;
    c.ctrl = p; p += 6 * c.Nc;
    c.shift = p; p += bsz;
    c.red = p;                                          // [ST][5]
    // dense operators: padded global copy (row-major n8 x ldk per operator) -> shared memory, once per CTA
    if (G.ops_smem)
        for (int e = threadIdx.x; e < (1 + 2 * c.Nc) * sz; e += blockDim.x) c.H0[e] = P.dense_ops[e];
    for (int e = threadIdx.x; e < 14 * bsz; e += blockDim.x) c.vr[e] = 0.0;     // padding rows / columns stay zero for good
    __syncthreads();

    for (int item = blockIdx.x; item < G.nitems; item += gridDim.x) {
        const int b = item / G.tiles_per_cand, s0 = (item % G.tiles_per_cand) * G.ST;
        c.ns = A.nsamples - s0 < G.ST ? A.nsamples - s0 : G.ST;
        c.ncols = c.m * c.ns;
        for (int k = threadIdx.x; k < c.Npar; k += blockDim.x) c.pcof[k] = A.pcof[(size_t)b * A.pstride + k];
        for (int k = threadIdx.x; k < G.ST * c.Npar; k += blockDim.x) c.gsm[k] = 0.0;
        for (int e = threadIdx.x; e < 14 * bsz; e += blockDim.x) c.vr[e] = 0.0;
        for (int e = threadIdx.x; e < bsz; e += blockDim.x) c.shift[e] = 0.0;
        __syncthreads();
        DN_FOR {
            c.vr[idx] = P.uinit[i + (size_t)c.n * (col % c.m)];
            if (A.shift) c.shift[idx] = A.shift[(size_t)(s0 + col / c.m) * c.n + i];
        }
        __syncthreads();
        const double phase = P.pFidType == 3 ? A.pcof[(size_t)b * A.pstride + c.Npar] : P.globalPhase;

        // ---------------- forward sweep ----------------
        double dt = P.T / (double)P.nsteps, t = 0.0;
        // lane 0 of the warp that owns sample s keeps its running penalty
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double pens[DN_MAXST / 2];                      // at least 2 warps per CTA: at most DN_MAXST / 2 samples per warp
        for (int k = 0; k < DN_MAXST / 2; ++k) pens[k] = 0.0;
        for (long long step = 0; step < P.nsteps; ++step) {
            dn_assemble(c, t, dt);
            dn_sample_sums<1>(c, c.red, [&](int, int i, int, int idx, double *a) { a[0] += P.wdiag[i] * c.vr[idx] * c.vr[idx]; });      // penalf2aTrap
            if (lane == 0) for (int s = warp, k = 0; s < c.ns; s += (blockDim.x >> 5), ++k) pens[k] += c.red[s];
            dn_state_step(c, dt);
            t = t + dt;
            dn_sample_sums<1>(c, c.red, [&](int, int i, int, int idx, double *a) {                                                    // penalf2a
                a[0] += P.wdiag[i] * (c.vr[idx] * c.vr[idx] + 2.0 * c.vi05[idx] * c.vi05[idx]);
            });
            if (lane == 0) for (int s = warp, k = 0; s < c.ns; s += (blockDim.x >> 5), ++k) pens[k] += c.red[s];
            __syncthreads();
        }
        // infidelity and terminal condition per sample
        dn_sample_sums<2>(c, c.red, [&](int, int i, int j, int idx, double *a) {
            const double tr_ = P.vtr[i + (size_t)c.n * j], ti_ = P.vti[i + (size_t)c.n * j];
            a[0] += c.vr[idx] * tr_ - c.vi[idx] * ti_;
            a[1] += c.vr[idx] * ti_ + c.vi[idx] * tr_;
        });
        double sph, cph;
        sincos(phase, &sph, &cph);
        const int pfid = P.pFidType;
        if (lane == 0)
            for (int s = warp, k = 0; s < c.ns; s += (blockDim.x >> 5), ++k) {
                const double re = c.red[2 * s] / c.m, im = c.red[2 * s + 1] / c.m, abs2 = re * re + im * im;
                const double infid = pfid == 1 ? 1.0 + abs2 - 2.0 * (re * cph + im * sph) : pfid == 2 ? 1.0 - abs2 : 1.0 - (re * cph - im * sph);
                const size_t traj = (size_t)b * A.nsamples + s0 + s;
                double *o = A.scal + traj * 4;
                o[0] = infid; o[1] = 0.5 * dt * c.tinv * pens[k]; o[2] = 1.0 - abs2; o[3] = 0.0;
                if (pfid == 3 && A.evaladjoint) A.grad[traj * A.gstride + c.Npar] = re * sph + im * cph;
            }
        if (!A.evaladjoint) { __syncthreads(); continue; }

        // ---------------- backward sweep ----------------
        DN_FOR {
            const int s = col / c.m, j = col % c.m;
            const double re = c.red[2 * s] / c.m, im = c.red[2 * s + 1] / c.m;
            const double rs_ = pfid == 1 ? cph - re : re, is_ = pfid == 1 ? sph - im : im;
            const double tr_ = P.vtr[i + (size_t)c.n * j], ti_ = P.vti[i + (size_t)c.n * j];
            double lr, li;
            if (pfid <= 2) { lr = (rs_ * tr_ + is_ * ti_) / c.m; li = (is_ * tr_ - rs_ * ti_) / c.m; }
            else { lr = 0.5 * (cph * tr_ - sph * ti_) / c.m; li = -0.5 * (sph * tr_ + cph * ti_) / c.m; }
            c.lr[idx] = lr; c.lr05[idx] = lr; c.li[idx] = li; c.li0[idx] = li;
        }
        t = P.T;
        dt = -dt;
        __syncthreads();
        for (long long step = P.nsteps - 1; step >= 0; --step) {
            const double t0 = t;
            DN_FOR c.vr0[idx] = c.vr[idx];
            dn_assemble(c, t, dt);
            dn_state_step(c, dt);
            t = t + dt;
            dn_adjoint_step(c, c.lr, c.li, c.lr05, dt);
            dn_grad_step(c, t0, dt);
            DN_FOR c.li0[idx] = c.li[idx];
            __syncthreads();
        }
        for (int k = threadIdx.x; k < c.ns * c.Npar; k += blockDim.x) {
            const int s = k / c.Npar, kk = k % c.Npar;
            A.grad[((size_t)b * A.nsamples + s0 + s) * A.gstride + kk] = dt * c.gsm[s * c.Npar + kk];
        }
        __syncthreads();
    }
}

int pad_ld(int x) {            // smallest ld >= x with ld % 16 in {4, 12}: 8 rows x 4 consecutive doubles hit 32 distinct banks
    int ld = x;
    while (ld % 16 != 4 && ld % 16 != 12) ++ld;
    return ld;
}
size_t dense_bytes(const DevProblem &P, const LaunchArgs &A, const DenseGeom &G) {
    const size_t d = (size_t)(6 + (G.ops_smem ? 1 + 2 * P.Nc : 0)) * G.n8 * G.ldk + (size_t)14 * G.N8 * G.ldx + A.Npar + (size_t)G.ST * A.Npar +
                     6 * P.Nc + (size_t)G.N8 * G.ldx + (size_t)G.ST * 5 + 8;
    return d * sizeof(double);
}

DenseGeom dense_geometry(const DevProblem &P, const LaunchArgs &A, size_t *bytes) {
    DenseGeom G{};
    G.n8 = (P.n + 7) & ~7;
    G.n4 = (P.n + 3) & ~3;
    G.ldk = pad_ld(G.n4);
    G.ldx = pad_ld(G.n8);
    const size_t limit = 227 * 1024;
    // the largest sample tile that fits, operators in shared memory if they fit too (else from their global copy)
    for (int ST = A.nsamples < DN_MAXST ? A.nsamples : DN_MAXST;; --ST) {
        G.ST = ST;
        G.N8 = (P.m * ST + 7) & ~7;
        G.ops_smem = 1;
        *bytes = dense_bytes(P, A, G);
        if (*bytes <= limit) break;
        if (ST == 1) {
            G.ops_smem = 0;
            *bytes = dense_bytes(P, A, G);
            break;
        }
    }
    const int tiles = (G.n8 / 8) * (G.N8 / 8);
    G.nthr = 32 * (tiles < 2 ? 2 : tiles > 8 ? 8 : tiles);
    G.tiles_per_cand = (A.nsamples + G.ST - 1) / G.ST;
    G.nitems = (A.ntraj / A.nsamples) * G.tiles_per_cand;
    return G;
}

}  // namespace

void jq_dense_padding(int n, int *n8, int *ldk) { *n8 = (n + 7) & ~7; *ldk = pad_ld((n + 3) & ~3); }

bool jq_dense_supported(const DevProblem &P, char *why, size_t len) {
    const char *r = nullptr;
    if (!P.dense_ops) r = "no dense operator copy (n > 64)";
    else if (P.objFuncType != 1) r = "objFuncType 2/3 run on the generic kernel";
    else if (P.solver != 1) r = "the Jacobi solver runs on the generic kernel";
    else if (P.wreal) r = "dense forbidden-state weights run on the generic kernel";
    if (r) { snprintf(why, len, "%s", r); return false; }
    why[0] = 0;
    return true;
}

cudaError_t jq_dense_launch(const DevProblem &P, const LaunchArgs &A, cudaStream_t st, int *nctas, int *regs, size_t *smem, int *traj_per_cta) {
    if (A.hist_r) return cudaErrorNotSupported;
    size_t bytes = 0;
    const DenseGeom G = dense_geometry(P, A, &bytes);
    if (bytes > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(jq_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, jq_dense_kernel);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jq_dense_kernel, G.nthr, bytes);
    if (occ < 1) occ = 1;
    const int grid = G.nitems < sms * occ ? G.nitems : sms * occ;
    jq_dense_kernel<<<grid, G.nthr, bytes, st>>>(P, A, G);
    if (nctas) *nctas = grid;
    if (regs) *regs = fa.numRegs;
    if (smem) *smem = bytes;
    if (traj_per_cta) *traj_per_cta = G.ST;
    return cudaGetLastError();
}
