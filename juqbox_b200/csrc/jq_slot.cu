// Warp-slot trajectory kernel: the state columns of a trajectory live in REGISTERS for the whole forward and
// backward time loops; shared memory is only the exchange medium of the sparse operator products.
//
// Mapping.  A "slot" is NL = 2^k lanes of a warp.  Lane l of a slot owns rows r = l, l+NL, .. (R rows) and C of the m
// columns of one trajectory, i.e. R*C elements of every n x m block as registers.  32/NL slots share a warp, a CTA
// of 4 warps holds TPC trajectories (each m/C slots).  Columns never couple inside the time loops, so a slot only
// needs __syncwarp between the store and the neighbour loads of a product; slots meet (through shared memory and
// __syncthreads) once per CH-step chunk, at the infidelity between the sweeps and at the final gradient sum.
//
// Operators.  Every row keeps, in registers, its H0 diagonal and, per control q, up to WQ entries
// (neighbour position, Hsym_q value, Hanti_q value) — the ladder-operator Hamiltonians of every named config have 2.
// One "pass" over a block x stores the lane's elements into the slot's exchange buffer and accumulates
//   A_q = Hsym_q x   and/or   D_q = Hanti_q x        (one neighbour load feeds both),
// from which every time level's product is a register-only combination:
//   K(t) x = h0 .* x + sum_q p_q(t) A_q ,   S(t) x = sum_q q_q(t) D_q .
// The A_q, D_q of the adjoint passes are also exactly what the gradient traces need
// (tr(A' H C) = sum A .* (H C)), so the gradient costs no extra products.
//
// Controls.  Time points are shared by all trajectories of a CTA: every CH steps all threads cooperatively fill a
// shared-memory table with knot index, the three quadratic B-spline values, cos/sin of every carrier at the 2CH+1
// time points of the chunk, and from it p_q, q_q for every resident trajectory.
//
// Reference lines: see jq_generic.cu (same algorithm, same order of the time levels); the only algebraic
// regrouping is S1*u + (h/2) S1*k1 = S1*(u + (h/2) k1)  (src/StormerVerlet.jl:483-484).
#include "jq_common.h"

#include <cstdio>
#include <vector>

#define SLOT_WARPS 4
#define SLOT_THREADS (SLOT_WARPS * 32)
#define SLOT_CH 16   // steps per control-table chunk

struct SlotParams {
    DevProblem P;
    LaunchArgs A;
    int NL, NLR, SPT, TPC, nslots;          // lanes per slot, NL*R, slots per trajectory, trajectories / slots per CTA
    const int *plan_pos;                    // [NLR][NC][WQ] neighbour row position (own position for padding)
    const double *plan_hs, *plan_ha;        // [NLR][NC][WQ]
    const double *plan_d0, *plan_w;         // [NLR]
    int o_exch, o_pcof, o_gsm, o_times, o_tabb, o_tabph, o_tabpq, o_red, o_tabk;   // offsets in doubles
};

struct SlotPlan {
    int R, C, NC, WQ, NL, NLR, SPT, TPC, nslots;
    int *d_pos = nullptr;
    double *d_hs = nullptr, *d_ha = nullptr, *d_d0 = nullptr, *d_w = nullptr;
};

namespace {

// ---------------------------------------------------------------------------------------------------------------
// exchange-buffer access: C columns of one row, laid out in planes so that consecutive lanes hit consecutive banks
template <int C> struct Xch;
template <> struct Xch<1> {
    static __device__ __forceinline__ void st(double *b, int pos, int, const double *x) { b[pos] = x[0]; }
    static __device__ __forceinline__ void ld(const double *b, int pos, int, double *x) { x[0] = b[pos]; }
};
template <> struct Xch<2> {
    static __device__ __forceinline__ void st(double *b, int pos, int, const double *x) { reinterpret_cast<double2 *>(b)[pos] = make_double2(x[0], x[1]); }
    static __device__ __forceinline__ void ld(const double *b, int pos, int, double *x) { double2 v = reinterpret_cast<const double2 *>(b)[pos]; x[0] = v.x; x[1] = v.y; }
};
template <> struct Xch<3> {
    static __device__ __forceinline__ void st(double *b, int pos, int nlr, const double *x) { b[pos] = x[0]; b[nlr + pos] = x[1]; b[2 * nlr + pos] = x[2]; }
    static __device__ __forceinline__ void ld(const double *b, int pos, int nlr, double *x) { x[0] = b[pos]; x[1] = b[nlr + pos]; x[2] = b[2 * nlr + pos]; }
};
template <> struct Xch<4> {
    static __device__ __forceinline__ void st(double *b, int pos, int nlr, const double *x) {
        reinterpret_cast<double2 *>(b)[pos] = make_double2(x[0], x[1]);
        reinterpret_cast<double2 *>(b)[nlr + pos] = make_double2(x[2], x[3]);
    }
    static __device__ __forceinline__ void ld(const double *b, int pos, int nlr, double *x) {
        double2 v = reinterpret_cast<const double2 *>(b)[pos], w = reinterpret_cast<const double2 *>(b)[nlr + pos];
        x[0] = v.x; x[1] = v.y; x[2] = w.x; x[3] = w.y;
    }
};

template <int R, int C, int NC, int WQ>
struct Lane {
    // operator rows (constant for the whole kernel)
    int pos[R][NC][WQ];
    double hs[R][NC][WQ], ha[R][NC][WQ];
    double d0[R], w[R];          // H0 diagonal (+ noise shift), guard weight / T
    int own[R];
    // control values at the three time levels of the current step: [level][q]
    double p[3][NC], q[3][NC];
    double *buf;                 // this slot's exchange buffers (2 x NLR*C doubles)
    int nlr, parity;
};

#define FOR_RC for (int k = 0; k < R; ++k) for (int c = 0; c < C; ++c)
#define UNROLL _Pragma("unroll")

// One pass: A_q = Hsym_q x and/or D_q = Hanti_q x for the lane's elements.
template <int R, int C, int NC, int WQ, bool WA, bool WD>
__device__ __forceinline__ void pass(Lane<R, C, NC, WQ> &L, const double (&x)[R][C], double (&A)[R][NC][C], double (&D)[R][NC][C]) {
    double *b = L.buf + L.parity * (L.nlr * C);
    L.parity ^= 1;
    UNROLL for (int k = 0; k < R; ++k) Xch<C>::st(b, L.own[k], L.nlr, x[k]);
    __syncwarp();
    UNROLL for (int k = 0; k < R; ++k)
        UNROLL for (int qq = 0; qq < NC; ++qq) {
            UNROLL for (int c = 0; c < C; ++c) { if (WA) A[k][qq][c] = 0.0; if (WD) D[k][qq][c] = 0.0; }
            UNROLL for (int e = 0; e < WQ; ++e) {
                double xv[C];
                Xch<C>::ld(b, L.pos[k][qq][e], L.nlr, xv);
                UNROLL for (int c = 0; c < C; ++c) {
                    if (WA) A[k][qq][c] = fma(L.hs[k][qq][e], xv[c], A[k][qq][c]);
                    if (WD) D[k][qq][c] = fma(L.ha[k][qq][e], xv[c], D[k][qq][c]);
                }
            }
        }
}

// X = sum_{j<=J} (h/2)^j S_level^j B   (src/linear_solvers.jl:94-106); B is consumed.
template <int R, int C, int NC, int WQ>
__device__ __forceinline__ void neumann(Lane<R, C, NC, WQ> &L, int J, double h, int level, double (&B)[R][C], double (&X)[R][C]) {
    double dummy[R][NC][C], D[R][NC][C];
    UNROLL FOR_RC X[k][c] = B[k][c];
    double coeff = 1.0;
    for (int it = 0; it < J; ++it) {
        pass<R, C, NC, WQ, false, true>(L, B, dummy, D);
        coeff *= 0.5 * h;
        UNROLL FOR_RC {
            double t = 0.0;
            UNROLL for (int qq = 0; qq < NC; ++qq) t = fma(L.q[level][qq], D[k][qq][c], t);
            B[k][c] = t;
            X[k][c] = fma(coeff, t, X[k][c]);
        }
    }
}

// src/StormerVerlet.jl:461-504.  u, v updated in place; v05 returned.
template <int R, int C, int NC, int WQ>
__device__ __forceinline__ void state_step(Lane<R, C, NC, WQ> &L, int J, double h, double (&u)[R][C], double (&v)[R][C], double (&v05)[R][C]) {
    double A[R][NC][C], D[R][NC][C], rhs[R][C], l1[R][C], s0u[R][C];
    pass<R, C, NC, WQ, true, true>(L, u, A, D);
    UNROLL FOR_RC {
        double r = L.d0[k] * u[k][c], s = 0.0;
        UNROLL for (int qq = 0; qq < NC; ++qq) { r = fma(L.p[1][qq], A[k][qq][c], r); s = fma(L.q[0][qq], D[k][qq][c], s); }
        rhs[k][c] = r;         // K05 u
        s0u[k][c] = s;         // S0 u
    }
    pass<R, C, NC, WQ, false, true>(L, v, A, D);
    UNROLL FOR_RC UNROLL for (int qq = 0; qq < NC; ++qq) rhs[k][c] = fma(L.q[1][qq], D[k][qq][c], rhs[k][c]);   // + S05 v
    neumann<R, C, NC, WQ>(L, J, h, 1, rhs, l1);
    UNROLL FOR_RC v05[k][c] = fma(0.5 * h, l1[k][c], v[k][c]);
    pass<R, C, NC, WQ, true, true>(L, v05, A, D);
    double k1v[R][C], s05v[R][C];
    UNROLL FOR_RC {
        double k0 = L.d0[k] * v05[k][c], k1 = k0, s = 0.0;
        UNROLL for (int qq = 0; qq < NC; ++qq) {
            k0 = fma(L.p[0][qq], A[k][qq][c], k0);
            k1 = fma(L.p[2][qq], A[k][qq][c], k1);
            s = fma(L.q[1][qq], D[k][qq][c], s);
        }
        k1v[k][c] = k1;                                         // K1 v05
        s05v[k][c] = s;                                         // S05 v05
        u[k][c] = fma(0.5 * h, s0u[k][c] - k0, u[k][c]);        // u + (h/2) kappa1,  kappa1 = S0 u - K0 v05
    }
    pass<R, C, NC, WQ, false, true>(L, u, A, D);
    UNROLL FOR_RC {
        double s = -k1v[k][c];
        UNROLL for (int qq = 0; qq < NC; ++qq) s = fma(L.q[2][qq], D[k][qq][c], s);
        rhs[k][c] = s;                                          // S1 (u + (h/2) kappa1) - K1 v05
    }
    double k2[R][C];
    neumann<R, C, NC, WQ>(L, J, h, 2, rhs, k2);
    UNROLL FOR_RC u[k][c] = fma(0.5 * h, k2[k][c], u[k][c]);
    pass<R, C, NC, WQ, true, false>(L, u, A, D);
    UNROLL FOR_RC {
        double l2 = fma(L.d0[k], u[k][c], s05v[k][c]);
        UNROLL for (int qq = 0; qq < NC; ++qq) l2 = fma(L.p[1][qq], A[k][qq][c], l2);
        v[k][c] = fma(0.5 * h, l1[k][c] + l2, v[k][c]);
    }
}

// src/StormerVerlet.jl:255-303 with the diagonal-W forcing of src/evalobjgrad.jl:862,882-888, fused with the five
// traces per control of adjoint_grad_calc! (src/evalobjgrad.jl:2578-2618):
//   T[q][0] = tr(vr0,Ha,lr05)  T[q][1] = tr(vi05,Hs,lr05)  T[q][2] = tr(vr,Ha,lr05)
//   T[q][3] = tr(vr,Hs,li)+tr(vr0,Hs,li0)                   T[q][4] = tr(vi05,Ha,li)+tr(vi05,Ha,li0)
template <int R, int C, int NC, int WQ>
__device__ __forceinline__ void adjoint_step(Lane<R, C, NC, WQ> &L, int J, double h, double (&mu)[R][C], double (&nu)[R][C],
                                             const double (&vr0)[R][C], const double (&vi05)[R][C], const double (&vr)[R][C],
                                             double (&T)[NC][5]) {
    double A[R][NC][C], D[R][NC][C], rhs[R][C], s05n[R][C];
    pass<R, C, NC, WQ, false, true>(L, mu, A, D);
    UNROLL FOR_RC {
        double s = L.w[k] * vr0[k][c];                          // hr0
        UNROLL for (int qq = 0; qq < NC; ++qq) s = fma(L.q[0][qq], D[k][qq][c], s);
        rhs[k][c] = s;                                          // S0 mu + hr0
    }
    pass<R, C, NC, WQ, true, true>(L, nu, A, D);
    UNROLL FOR_RC {
        double kk = L.d0[k] * nu[k][c], s = 0.0;
        UNROLL for (int qq = 0; qq < NC; ++qq) {
            kk = fma(L.p[1][qq], A[k][qq][c], kk);
            s = fma(L.q[1][qq], D[k][qq][c], s);
            T[qq][3] = fma(vr0[k][c], A[k][qq][c], T[qq][3]);
            T[qq][4] = fma(vi05[k][c], D[k][qq][c], T[qq][4]);
        }
        rhs[k][c] -= kk;                                        // - K05 nu
        s05n[k][c] = s;                                         // S05 nu
    }
    double k2[R][C];
    neumann<R, C, NC, WQ>(L, J, h, 0, rhs, k2);
    UNROLL FOR_RC mu[k][c] = fma(0.5 * h, k2[k][c], mu[k][c]);  // X = mu
    pass<R, C, NC, WQ, true, true>(L, mu, A, D);
    double l2[R][C], k1x[R][C], s1x[R][C];
    UNROLL FOR_RC {
        double k0 = L.d0[k] * mu[k][c], k1 = k0, s = 0.0;
        UNROLL for (int qq = 0; qq < NC; ++qq) {
            k0 = fma(L.p[0][qq], A[k][qq][c], k0);
            k1 = fma(L.p[2][qq], A[k][qq][c], k1);
            s = fma(L.q[2][qq], D[k][qq][c], s);
            T[qq][0] = fma(vr0[k][c], D[k][qq][c], T[qq][0]);
            T[qq][1] = fma(vi05[k][c], A[k][qq][c], T[qq][1]);
            T[qq][2] = fma(vr[k][c], D[k][qq][c], T[qq][2]);
        }
        const double hi0 = L.w[k] * vi05[k][c];
        l2[k][c] = k0 + s05n[k][c] + hi0;                       // K0 X + S05 nu + hi0
        k1x[k][c] = k1 + hi0;                                   // K1 X + hi1
        s1x[k][c] = s;                                          // S1 X
    }
    pass<R, C, NC, WQ, false, true>(L, l2, A, D);
    UNROLL FOR_RC {
        double s = 0.0;
        UNROLL for (int qq = 0; qq < NC; ++qq) s = fma(L.q[1][qq], D[k][qq][c], s);
        rhs[k][c] = s05n[k][c] + 0.5 * h * s + k1x[k][c];        // S05 nu + (h/2) S05 l2 + K1 X + hi1
    }
    double l1[R][C];
    neumann<R, C, NC, WQ>(L, J, h, 1, rhs, l1);
    UNROLL FOR_RC nu[k][c] = fma(0.5 * h, l2[k][c] + l1[k][c], nu[k][c]);
    pass<R, C, NC, WQ, true, true>(L, nu, A, D);
    UNROLL FOR_RC {
        double kk = L.d0[k] * nu[k][c];
        UNROLL for (int qq = 0; qq < NC; ++qq) {
            kk = fma(L.p[1][qq], A[k][qq][c], kk);
            T[qq][3] = fma(vr[k][c], A[k][qq][c], T[qq][3]);
            T[qq][4] = fma(vi05[k][c], D[k][qq][c], T[qq][4]);
        }
        mu[k][c] = fma(0.5 * h, s1x[k][c] - kk + L.w[k] * vr[k][c], mu[k][c]);   // kappa1 = S1 X - K05 nu + hr1
    }
}

__device__ __forceinline__ double slot_sum(double x, int NL) {
    for (int o = 1; o < NL; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Fill the control table for `nst` steps starting at time t (all threads of the CTA).
template <int NC>
__device__ void fill_table(const SlotParams &S, double *sm, double t, double dt, int nst, double dtknot) {
    double *times = sm + S.o_times, *tabb = sm + S.o_tabb, *tabph = sm + S.o_tabph, *tabpq = sm + S.o_tabpq;
    int *tabk = reinterpret_cast<int *>(sm + S.o_tabk);
    const int npts = 2 * nst + 1, Nfreq = S.P.Nfreq, D1 = S.A.D1;
    __syncthreads();                       // the previous chunk's table is no longer in use
    if (threadIdx.x == 0) {
        double tt = t;
        times[0] = tt;
        for (int i = 0; i < nst; ++i) {    // same recurrence as the reference: t + 0.5 dt, then t = t + dt
            times[2 * i + 1] = tt + 0.5 * dt;
            tt = tt + dt;
            times[2 * i + 2] = tt;
        }
    }
    __syncthreads();
    const double width = 3.0 * dtknot;
    for (int idx = threadIdx.x; idx < npts * (NC * Nfreq + 1); idx += SLOT_THREADS) {
        const int i = idx / (NC * Nfreq + 1), j = idx % (NC * Nfreq + 1);
        const double tt = times[i];
        if (j == NC * Nfreq) {             // src/bsplines.jl:224-253
            long long k = (long long)ceil(tt / dtknot + 2.0);
            k = k < 3 ? 3 : (k > D1 ? D1 : k);
            tabk[i] = (int)k;
            double tau = (tt - dtknot * ((double)k - 1.5)) / width;
            tabb[3 * i + 0] = 9.0 / 8 + 4.5 * tau + 4.5 * tau * tau;
            tau = (tt - dtknot * ((double)(k - 1) - 1.5)) / width;
            tabb[3 * i + 1] = 0.75 - 9.0 * tau * tau;
            tau = (tt - dtknot * ((double)(k - 2) - 1.5)) / width;
            tabb[3 * i + 2] = 9.0 / 8 - 4.5 * tau + 4.5 * tau * tau;
        } else {
            const int qq = j / Nfreq, fr = j % Nfreq;
            double sn, cs;
            sincos(S.P.cfreq[qq + NC * fr] * tt, &sn, &cs);
            tabph[2 * (i * NC * Nfreq + j)] = cs;
            tabph[2 * (i * NC * Nfreq + j) + 1] = sn;
        }
    }
    __syncthreads();
    const double *pcof = sm + S.o_pcof;
    for (int idx = threadIdx.x; idx < npts * S.TPC * NC; idx += SLOT_THREADS) {
        const int i = idx / (S.TPC * NC), rem = idx % (S.TPC * NC), tr = rem / NC, qq = rem % NC;
        const int k = tabk[i];
        const double b0 = tabb[3 * i], b1 = tabb[3 * i + 1], b2 = tabb[3 * i + 2];
        const double *pc = pcof + tr * S.A.Npar;
        double pv = 0.0, qv = 0.0;
        for (int fr = 0; fr < Nfreq; ++fr) {   // src/bsplines.jl:229-261
            const int off1 = 2 * qq * Nfreq * D1 + fr * 2 * D1 - 1, off2 = off1 + D1;
            const double fbs1 = pc[off1 + k] * b0 + pc[off1 + k - 1] * b1 + pc[off1 + k - 2] * b2;
            const double fbs2 = pc[off2 + k] * b0 + pc[off2 + k - 1] * b1 + pc[off2 + k - 2] * b2;
            const double cs = tabph[2 * (i * NC * Nfreq + qq * Nfreq + fr)], sn = tabph[2 * (i * NC * Nfreq + qq * Nfreq + fr) + 1];
            pv += fbs1 * cs - fbs2 * sn;
            qv += fbs1 * sn + fbs2 * cs;
        }
        tabpq[(i * S.TPC + tr) * 2 * NC + 2 * qq] = pv;
        tabpq[(i * S.TPC + tr) * 2 * NC + 2 * qq + 1] = qv;
    }
    __syncthreads();
}

template <int R, int C, int NC, int WQ>
__global__ void __launch_bounds__(SLOT_THREADS) jq_slot_kernel(const __grid_constant__ SlotParams S) {
    extern __shared__ double sm[];
    const DevProblem &P = S.P;
    const LaunchArgs &A = S.A;
    const int NL = S.NL, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane & (NL - 1);                          // lane within slot
    const int slot = warp * (32 / NL) + (lane / NL);        // slot within CTA
    const int tloc = slot / S.SPT, cg = slot % S.SPT;       // trajectory within CTA, column group
    const int traj = blockIdx.x * S.TPC + tloc;
    const bool live = tloc < S.TPC && traj < A.ntraj;       // dead slots compute on zeros and write nothing
    const int b = live ? traj / A.nsamples : 0, s = live ? traj % A.nsamples : 0;
    const int n = P.n, m = P.m, c0 = cg * C, Npar = A.Npar, D1 = A.D1, Nfreq = P.Nfreq, J = P.J;
    const double tinv = 1.0 / P.T, dtknot = P.T / (D1 - 2);
    const int tl = tloc < S.TPC ? tloc : 0;                 // table row used by this slot

    Lane<R, C, NC, WQ> L;
    L.nlr = S.NLR;
    L.parity = 0;
    L.buf = sm + S.o_exch + slot * (2 * S.NLR * C);
    UNROLL for (int k = 0; k < R; ++k) {
        const int r = k * NL + l;
        L.own[k] = r;
        L.d0[k] = S.plan_d0[r] + ((live && A.shift && r < n) ? A.shift[(size_t)s * n + r] : 0.0);
        L.w[k] = S.plan_w[r] * tinv;
        UNROLL for (int qq = 0; qq < NC; ++qq)
            UNROLL for (int e = 0; e < WQ; ++e) {
                const int ix = (r * NC + qq) * WQ + e;
                L.pos[k][qq][e] = S.plan_pos[ix];
                L.hs[k][qq][e] = S.plan_hs[ix];
                L.ha[k][qq][e] = S.plan_ha[ix];
            }
    }
    // stage this CTA's pcof vectors and zero the per-slot gradient accumulators
    for (int idx = threadIdx.x; idx < S.TPC * Npar; idx += SLOT_THREADS) {
        const int tr = idx / Npar, k = idx % Npar, tg = blockIdx.x * S.TPC + tr;
        sm[S.o_pcof + idx] = tg < A.ntraj ? A.pcof[(size_t)(tg / A.nsamples) * Npar + k] : 0.0;
    }
    for (int idx = threadIdx.x; idx < S.nslots * Npar; idx += SLOT_THREADS) sm[S.o_gsm + idx] = 0.0;

    double vr[R][C], vi[R][C], vi05[R][C];
    UNROLL FOR_RC {
        const int r = L.own[k];
        vr[k][c] = (live && r < n) ? P.uinit[r + (size_t)n * (c0 + c)] : 0.0;
        vi[k][c] = 0.0;
        vi05[k][c] = 0.0;
    }
    const double *tabpq = sm + S.o_tabpq;

    // ------------------------------------------------------------ forward sweep (src/evalobjgrad.jl:698-753)
    double dt = P.T / (double)P.nsteps, t = 0.0, pen = 0.0;
    for (long long s0 = 0; s0 < P.nsteps; s0 += SLOT_CH) {
        const int nst = (int)((P.nsteps - s0) < SLOT_CH ? (P.nsteps - s0) : SLOT_CH);
        fill_table<NC>(S, sm, t, dt, nst, dtknot);
        UNROLL for (int qq = 0; qq < NC; ++qq) { L.p[2][qq] = tabpq[(0 * S.TPC + tl) * 2 * NC + 2 * qq]; L.q[2][qq] = tabpq[(0 * S.TPC + tl) * 2 * NC + 2 * qq + 1]; }
        for (int ls = 0; ls < nst; ++ls) {
            UNROLL for (int qq = 0; qq < NC; ++qq) {
                L.p[0][qq] = L.p[2][qq]; L.q[0][qq] = L.q[2][qq];
                const double *r1 = tabpq + ((2 * ls + 1) * S.TPC + tl) * 2 * NC, *r2 = tabpq + ((2 * ls + 2) * S.TPC + tl) * 2 * NC;
                L.p[1][qq] = r1[2 * qq]; L.q[1][qq] = r1[2 * qq + 1];
                L.p[2][qq] = r2[2 * qq]; L.q[2][qq] = r2[2 * qq + 1];
            }
            UNROLL FOR_RC pen = fma(L.w[k], vr[k][c] * vr[k][c], pen);                                  // penalf2aTrap
            state_step<R, C, NC, WQ>(L, J, dt, vr, vi, vi05);
            UNROLL FOR_RC pen = fma(L.w[k], vr[k][c] * vr[k][c] + 2.0 * vi05[k][c] * vi05[k][c], pen);   // penalf2a
            t = t + dt;
        }
    }
    // infidelity (pFidType 2) and leak: slot partials -> shared -> per-trajectory sums in slot order
    double *red = sm + S.o_red;
    {
        double re = 0.0, im = 0.0;
        UNROLL FOR_RC {
            const int r = L.own[k];
            const double tr_ = (live && r < n) ? P.vtr[r + (size_t)n * (c0 + c)] : 0.0, ti_ = (live && r < n) ? P.vti[r + (size_t)n * (c0 + c)] : 0.0;
            re += vr[k][c] * tr_ - vi[k][c] * ti_;
            im += vr[k][c] * ti_ + vi[k][c] * tr_;
        }
        re = slot_sum(re, NL); im = slot_sum(im, NL); pen = slot_sum(pen, NL);
        __syncthreads();
        if (l == 0) { red[slot * 4] = re; red[slot * 4 + 1] = im; red[slot * 4 + 2] = pen; }
        __syncthreads();
    }
    double rs = 0.0, is = 0.0, pens = 0.0;
    for (int j = 0; j < S.SPT; ++j) {
        const int sl = (tloc < S.TPC ? tloc : 0) * S.SPT + j;
        rs += red[sl * 4]; is += red[sl * 4 + 1]; pens += red[sl * 4 + 2];
    }
    rs /= m; is /= m;
    const double infid = 1.0 - (rs * rs + is * is);
    if (live && cg == 0 && l == 0) {
        double *o = A.scal + (size_t)traj * 4;
        o[0] = infid; o[1] = 0.5 * dt * pens; o[2] = infid; o[3] = 0.0;   // w already carries 1/T
    }
    if (!A.evaladjoint) return;

    // ------------------------------------------------------------ backward sweep (src/evalobjgrad.jl:810-921)
    double lr[R][C], li[R][C], vr0[R][C];
    UNROLL FOR_RC {
        const int r = L.own[k];
        const double tr_ = (live && r < n) ? P.vtr[r + (size_t)n * (c0 + c)] : 0.0, ti_ = (live && r < n) ? P.vti[r + (size_t)n * (c0 + c)] : 0.0;
        lr[k][c] = (rs * tr_ + is * ti_) / m;     // init_adjoint!, src/evalobjgrad.jl:2029-2042
        li[k][c] = (is * tr_ - rs * ti_) / m;
    }
    // gradient scatter: lane u < NU of the slot owns (control, frequency, alpha) and a 3-knot register window
    const int NU = NC * Nfreq * 2;
    const bool upd = l < NU;
    const int uq = upd ? l / (2 * Nfreq) : 0, uf = upd ? (l >> 1) % Nfreq : 0, ua = l & 1;
    const int gbase = 2 * uq * Nfreq * D1 + uf * 2 * D1 + ua * D1 - 1;
    double *gsm = sm + S.o_gsm + slot * Npar;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    int kw = D1;
    const double *tabb = sm + S.o_tabb, *tabph = sm + S.o_tabph;
    const int *tabk = reinterpret_cast<const int *>(sm + S.o_tabk);

    t = P.T;
    dt = -dt;
    for (long long s0 = 0; s0 < P.nsteps; s0 += SLOT_CH) {
        const int nst = (int)((P.nsteps - s0) < SLOT_CH ? (P.nsteps - s0) : SLOT_CH);
        fill_table<NC>(S, sm, t, dt, nst, dtknot);
        UNROLL for (int qq = 0; qq < NC; ++qq) { L.p[2][qq] = tabpq[(0 * S.TPC + tl) * 2 * NC + 2 * qq]; L.q[2][qq] = tabpq[(0 * S.TPC + tl) * 2 * NC + 2 * qq + 1]; }
        for (int ls = 0; ls < nst; ++ls) {
            UNROLL for (int qq = 0; qq < NC; ++qq) {
                L.p[0][qq] = L.p[2][qq]; L.q[0][qq] = L.q[2][qq];
                const double *r1 = tabpq + ((2 * ls + 1) * S.TPC + tl) * 2 * NC, *r2 = tabpq + ((2 * ls + 2) * S.TPC + tl) * 2 * NC;
                L.p[1][qq] = r1[2 * qq]; L.q[1][qq] = r1[2 * qq + 1];
                L.p[2][qq] = r2[2 * qq]; L.q[2][qq] = r2[2 * qq + 1];
            }
            UNROLL FOR_RC vr0[k][c] = vr[k][c];
            state_step<R, C, NC, WQ>(L, J, dt, vr, vi, vi05);
            double T[NC][5];
            UNROLL for (int qq = 0; qq < NC; ++qq) UNROLL for (int j = 0; j < 5; ++j) T[qq][j] = 0.0;
            adjoint_step<R, C, NC, WQ>(L, J, dt, lr, li, vr0, vi05, vr, T);
            UNROLL for (int qq = 0; qq < NC; ++qq) UNROLL for (int j = 0; j < 5; ++j) T[qq][j] = slot_sum(T[qq][j], NL);
            if (upd) {
                double Tq[5];
                UNROLL for (int j = 0; j < 5; ++j) {
                    Tq[j] = T[0][j];
                    UNROLL for (int qq = 1; qq < NC; ++qq) Tq[j] = (uq == qq) ? T[qq][j] : Tq[j];
                }
                // time points in decreasing order: t0 (row 2ls), t0 + dt/2 (2ls+1), t0 + dt (2ls+2)
                UNROLL for (int tp = 0; tp < 3; ++tp) {
                    const int i = 2 * ls + tp;
                    const double Pc = tp == 1 ? Tq[3] : -Tq[1];
                    const double Qc = tp == 0 ? -Tq[0] : (tp == 1 ? -Tq[4] : -Tq[2]);
                    const double cs = tabph[2 * (i * NC * Nfreq + uq * Nfreq + uf)], sn = tabph[2 * (i * NC * Nfreq + uq * Nfreq + uf) + 1];
                    const double X = ua == 0 ? Pc * cs + Qc * sn : Qc * cs - Pc * sn;
                    const int k = tabk[i];
                    while (kw > k) { gsm[gbase + kw] += acc0; acc0 = acc1; acc1 = acc2; acc2 = 0.0; --kw; }
                    acc0 = fma(tabb[3 * i], X, acc0);
                    acc1 = fma(tabb[3 * i + 1], X, acc1);
                    acc2 = fma(tabb[3 * i + 2], X, acc2);
                }
            }
            t = t + dt;
        }
    }
    if (upd) { gsm[gbase + kw] += acc0; gsm[gbase + kw - 1] += acc1; gsm[gbase + kw - 2] += acc2; }
    __syncthreads();
    // total gradient of each resident trajectory = dt * sum of its slots' partial gradients, in slot order
    for (int idx = threadIdx.x; idx < S.TPC * Npar; idx += SLOT_THREADS) {
        const int tr = idx / Npar, k = idx % Npar, tg = blockIdx.x * S.TPC + tr;
        if (tg >= A.ntraj) continue;
        double g = 0.0;
        for (int j = 0; j < S.SPT; ++j) g += sm[S.o_gsm + (tr * S.SPT + j) * Npar + k];
        A.grad[(size_t)tg * Npar + k] = dt * g;
    }
}

typedef void (*slot_kernel_t)(const SlotParams);
struct Inst { int R, C, NC, WQ; slot_kernel_t fn; };
#define INST(R, C, NC, WQ) {R, C, NC, WQ, jq_slot_kernel<R, C, NC, WQ>}
const Inst kInst[] = {
    INST(1, 2, 1, 2), INST(1, 3, 1, 2), INST(1, 4, 1, 2), INST(1, 4, 2, 2), INST(1, 2, 2, 2),
    INST(2, 2, 3, 2), INST(3, 1, 3, 2), INST(1, 1, 1, 2), INST(1, 1, 2, 2),
};

const Inst *find_inst(int R, int C, int NC, int WQ) {
    for (const Inst &i : kInst)
        if (i.R == R && i.C == C && i.NC == NC && i.WQ == WQ) return &i;
    return nullptr;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
SlotPlan *jq_slot_plan_create(const DevProblem &P, const HostOps &H, char *err, size_t errlen) {
    const int n = H.n, m = H.m, Nc = H.Nc;
    auto no = [&](const char *why) { snprintf(err, errlen, "%s", why); return (SlotPlan *)nullptr; };
    if (P.objFuncType != 1) return no("objFuncType != 1 uses the generic kernel (second adjoint set)");
    if (n > 128) return no("n > 128");
    // H0 must be diagonal for this kernel
    std::vector<double> d0(n, 0.0);
    for (int r = 0; r < n; ++r)
        for (int p = H.rowptr[r]; p < H.rowptr[r + 1]; ++p) {
            if (H.col[p] == r) d0[r] = H.val[p];
            else if (H.val[p] != 0.0) return no("Hconst has off-diagonal entries");
        }
    // union pattern of Hsym_q / Hanti_q per row
    int WQ = 0;
    std::vector<std::vector<std::vector<int>>> cols(n, std::vector<std::vector<int>>(Nc));
    for (int q = 0; q < Nc; ++q)
        for (int r = 0; r < n; ++r) {
            std::vector<int> &cc = cols[r][q];
            for (int o : {1 + q, 1 + Nc + q}) {
                const int *rp = H.rowptr + o * (n + 1);
                for (int p = rp[r]; p < rp[r + 1]; ++p) {
                    bool have = false;
                    for (int x : cc) have |= (x == H.col[p]);
                    if (!have) cc.push_back(H.col[p]);
                }
            }
            WQ = (int)cc.size() > WQ ? (int)cc.size() : WQ;
        }
    if (WQ > 2) return no("more than 2 entries per row and control (not a ladder-type control Hamiltonian)");
    WQ = 2;
    // lanes per slot: power of two covering n with R <= 4 rows per lane and enough lanes for the gradient scatter
    int NL = 2;
    while (NL < 32 && NL < n) NL <<= 1;
    while (NL < 32 && NL < Nc * H.Nfreq * 2) NL <<= 1;
    if (NL < Nc * H.Nfreq * 2) return no("more (control, frequency) pairs than lanes");
    const int R = (n + NL - 1) / NL;
    int C = 1;
    for (int c = 1; c <= 4; ++c) if (m % c == 0 && R * c <= 4) C = c;
    const Inst *inst = find_inst(R, C, Nc, WQ);
    if (!inst && C > 1) { for (int c = C - 1; c >= 1 && !inst; --c) if (m % c == 0) { inst = find_inst(R, c, Nc, WQ); if (inst) C = c; } }
    if (!inst) return no("no template instantiation for this (rows per lane, columns per lane, controls)");

    SlotPlan *pl = new SlotPlan();
    pl->R = R; pl->C = C; pl->NC = Nc; pl->WQ = WQ; pl->NL = NL; pl->NLR = NL * R;
    pl->SPT = m / C;
    pl->nslots = SLOT_WARPS * (32 / NL);
    pl->TPC = pl->nslots / pl->SPT;
    if (pl->TPC < 1) { delete pl; return no("a trajectory does not fit in one CTA"); }
    const int NLR = pl->NLR;
    std::vector<int> pos((size_t)NLR * Nc * WQ);
    std::vector<double> hs((size_t)NLR * Nc * WQ, 0.0), ha((size_t)NLR * Nc * WQ, 0.0), d0p(NLR, 0.0), wp(NLR, 0.0);
    std::vector<double> wd(n);
    cudaMemcpy(wd.data(), P.wdiag, sizeof(double) * n, cudaMemcpyDeviceToHost);
    auto value_at = [&](int o, int r, int c) {
        const int *rp = H.rowptr + o * (n + 1);
        double v = 0.0;
        for (int p = rp[r]; p < rp[r + 1]; ++p) if (H.col[p] == c) v += H.val[p];
        return v;
    };
    for (int r = 0; r < NLR; ++r) {
        if (r < n) { d0p[r] = d0[r]; wp[r] = wd[r]; }
        for (int q = 0; q < Nc; ++q)
            for (int e = 0; e < WQ; ++e) {
                const size_t ix = ((size_t)r * Nc + q) * WQ + e;
                pos[ix] = r;                                     // padding: own position, zero coefficients
                if (r < n && e < (int)cols[r][q].size()) {
                    const int c = cols[r][q][e];
                    pos[ix] = c;
                    hs[ix] = value_at(1 + q, r, c);
                    ha[ix] = value_at(1 + Nc + q, r, c);
                }
            }
    }
    bool ok = cudaMalloc(&pl->d_pos, pos.size() * sizeof(int)) == cudaSuccess && cudaMalloc(&pl->d_hs, hs.size() * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&pl->d_ha, ha.size() * sizeof(double)) == cudaSuccess && cudaMalloc(&pl->d_d0, NLR * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&pl->d_w, NLR * sizeof(double)) == cudaSuccess;
    if (!ok) { jq_slot_plan_destroy(pl); return no("cudaMalloc failed for the slot plan"); }
    cudaMemcpy(pl->d_pos, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->d_hs, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->d_ha, ha.data(), ha.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->d_d0, d0p.data(), NLR * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->d_w, wp.data(), NLR * sizeof(double), cudaMemcpyHostToDevice);
    err[0] = 0;
    return pl;
}

void jq_slot_plan_destroy(SlotPlan *pl) {
    if (!pl) return;
    for (void *p : {(void *)pl->d_pos, (void *)pl->d_hs, (void *)pl->d_ha, (void *)pl->d_d0, (void *)pl->d_w}) if (p) cudaFree(p);
    delete pl;
}

cudaError_t jq_slot_launch(SlotPlan *pl, const DevProblem &P, const LaunchArgs &A, cudaStream_t st, int *nctas, int *regs,
                           size_t *smem, int *traj_per_cta) {
    const Inst *inst = find_inst(pl->R, pl->C, pl->NC, pl->WQ);
    if (!inst) return cudaErrorNotSupported;
    SlotParams S{};
    S.P = P; S.A = A;
    S.NL = pl->NL; S.NLR = pl->NLR; S.SPT = pl->SPT; S.TPC = pl->TPC; S.nslots = pl->nslots;
    S.plan_pos = pl->d_pos; S.plan_hs = pl->d_hs; S.plan_ha = pl->d_ha; S.plan_d0 = pl->d_d0; S.plan_w = pl->d_w;
    const int npts = 2 * SLOT_CH + 1, NC = pl->NC;
    int o = 0;
    auto take = [&](int cnt) { int at = o; o += (cnt + 1) & ~1; return at; };   // keep 16-byte alignment
    S.o_exch = take(pl->nslots * 2 * pl->NLR * pl->C);
    S.o_pcof = take(pl->TPC * A.Npar);
    S.o_gsm = take(pl->nslots * A.Npar);
    S.o_times = take(npts);
    S.o_tabb = take(3 * npts);
    S.o_tabph = take(2 * npts * NC * P.Nfreq);
    S.o_tabpq = take(npts * pl->TPC * 2 * NC);
    S.o_red = take(pl->nslots * 4);
    S.o_tabk = take((npts + 1) / 2);
    const size_t bytes = (size_t)o * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(inst->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, inst->fn);
    if (e != cudaSuccess) return e;
    const int grid = (A.ntraj + pl->TPC - 1) / pl->TPC;
    inst->fn<<<grid, SLOT_THREADS, bytes, st>>>(S);
    if (nctas) *nctas = grid;
    if (regs) *regs = fa.numRegs;
    if (smem) *smem = bytes;
    if (traj_per_cta) *traj_per_cta = pl->TPC;
    return cudaGetLastError();
}
