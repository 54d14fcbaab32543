"""Optimizer glue: the mirror of setup_ipopt_problem / run_optimizer (src/ipopt_interface.jl:267-437).

The reference drives Ipopt (L-BFGS Hessian approximation, box bounds on the coefficients, optionally the leakage as an
inequality constraint).  Ipopt is not available in this image; this mirror keeps the same callback layer —
eval_f_par / eval_grad_f_par / eval_g_par / eval_jac_g_par with the last-evaluation cache, Tikhonov terms, convergence
history and early-stop thresholds of intermediate_par (:212-240) — and hands it to scipy's L-BFGS-B (objFuncType 1/2)
or SLSQP (objFuncType 3, leak <= leak_ubound).  Every objective/gradient evaluation is one batched GPU call.
Optimizer trajectories are not expected to match Ipopt's (unpinned, out of scope); the callbacks are.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .api import Working_Arrays, eval_f_par, eval_g_par, eval_grad_f_par, eval_jac_g_par
from .params import objparams


class _StopOptimization(Exception):
    pass


@dataclass
class IpoptProblemMirror:
    params: objparams
    wa: Working_Arrays
    nCoeff: int
    minCoeff: np.ndarray
    maxCoeff: np.ndarray
    maxIter: int = 50
    lbfgsMax: int = 10
    ipTol: float = 1e-5
    nodes: np.ndarray = field(default_factory=lambda: np.array([0.0]))
    weights: np.ndarray = field(default_factory=lambda: np.array([1.0]))
    x: Optional[np.ndarray] = None
    status: str = ""


def setup_ipopt_problem(params, wa, nCoeff, minCoeff, maxCoeff, maxIter=50, lbfgsMax=10, startFromScratch=True, ipTol=1e-5,
                        acceptTol=1e-5, acceptIter=15, nodes=(0.0,), weights=(1.0,)):
    """Same signature as the reference (src/ipopt_interface.jl:267-273); returns a problem object for run_optimizer."""
    rng = np.random.default_rng(0)
    params.last_pcof = 1e9 * rng.random(nCoeff)            # :277-281: invalidate the evaluation cache
    params.last_infidelity_grad = 1e9 * rng.random(nCoeff)
    if params.objFuncType != 1:
        params.last_leak_grad = 1e9 * rng.random(nCoeff)
    return IpoptProblemMirror(params, wa, int(nCoeff), np.asarray(minCoeff, float), np.asarray(maxCoeff, float), int(maxIter),
                              int(lbfgsMax), float(ipTol), np.atleast_1d(np.asarray(nodes, float)),
                              np.atleast_1d(np.asarray(weights, float)))


def run_optimizer(prob: IpoptProblemMirror, pcof0, baseName: str = ""):
    """Mirror of run_optimizer (src/ipopt_interface.jl:413-437): returns the optimised coefficient vector."""
    from scipy.optimize import minimize
    p, wa = prob.params, prob.wa
    x0 = np.clip(np.asarray(pcof0, float).copy(), prob.minCoeff, prob.maxCoeff)
    # L-BFGS-B / SLSQP have no barrier: work in box-scaled variables y = x / scale so that the first steps stay inside
    scale = np.maximum(np.abs(prob.minCoeff), np.abs(prob.maxCoeff))
    scale[scale == 0.0] = 1.0

    def f(y):
        return float(eval_f_par(y * scale, p, wa, prob.nodes, prob.weights))

    def g(y):
        out = np.zeros(prob.nCoeff)
        eval_grad_f_par(y * scale, out, p, wa, prob.nodes, prob.weights)
        return out * scale

    def intermediate(yk, *_):
        # intermediate_par (:212-240): history + early stop on the objective / trace-infidelity thresholds
        obj = f(yk)
        if p.saveConvHist:
            p.objHist.append(obj)
            p.primaryHist.append(p.lastTraceInfidelity)
            p.secondaryHist.append(p.lastLeakIntegral)
        if obj < p.objThreshold or p.lastTraceInfidelity < p.traceInfidelityThreshold:
            raise _StopOptimization()

    bounds = list(zip(prob.minCoeff / scale, prob.maxCoeff / scale))
    x0 = x0 / scale
    try:
        if p.objFuncType == 3:
            def gfun(y):
                buf = np.zeros(1)
                return p.leak_ubound - eval_g_par(y * scale, buf, p, wa, prob.nodes, prob.weights)

            def gjac(y):
                if np.linalg.norm(y * scale - p.last_pcof) > 1e-15:
                    f(y)
                jac = np.zeros(prob.nCoeff)
                eval_jac_g_par(y * scale, np.zeros(0, np.int32), np.zeros(0, np.int32), jac, p, wa, prob.nodes, prob.weights)
                return -jac * scale
            res = minimize(f, x0, jac=g, bounds=bounds, method="SLSQP", callback=intermediate,
                           constraints=[{"type": "ineq", "fun": gfun, "jac": gjac}],
                           options={"maxiter": prob.maxIter, "ftol": prob.ipTol * 1e-3})
        else:
            res = minimize(f, x0, jac=g, bounds=bounds, method="L-BFGS-B", callback=intermediate,
                           options={"maxiter": prob.maxIter, "maxcor": prob.lbfgsMax, "gtol": prob.ipTol, "ftol": 1e-15})
        prob.x, prob.status = res.x * scale, str(res.message)
    except _StopOptimization:
        prob.x, prob.status = p.last_pcof.copy(), "stopped by objective / trace-infidelity threshold"
    if baseName:
        np.savetxt(baseName + ".dat", prob.x, fmt="%.13e")      # the reference's .dat format (one %.13e per line)
    return prob.x


def run_optimizer_multistart(prob: IpoptProblemMirror, pcof0s, maxIter: Optional[int] = None, lbfgsMax: Optional[int] = None,
                             gtol: float = 1e-7, max_backtracks: int = 12, ls_batch: int = 4):
    """Many independent optimisations advanced in lock step (north star: "batched candidate pcof vectors for multi-start
    or line search").  Start vectors pcof0s [B, nCoeff]; every objective/gradient request of all B members is ONE
    jq_traceobjgrad_batch call, so B starts cost about as much wall time as one while the GPU has idle SMs.

    Each member runs the same deterministic algorithm: projected L-BFGS (two-loop recursion over the last `lbfgsMax`
    pairs, box-scaled variables as in run_optimizer) with Armijo backtracking; the objective is infidelity + leak +
    Tikhonov (eval_f_par's objective, src/ipopt_interface.jl:77-100), with the risk-neutral nodes/weights of `prob`.
    Returns (pcofs [B, nCoeff], objective [B], history [iters + 1, B]).  A member's iterates do not depend on which other
    members share the batch.  The Armijo backtracking is batched too: `ls_batch` trial steps (1, 1/2, 1/4, ...) of every member
    that still searches go into ONE objective-only launch and the first acceptable one is taken, so an iteration usually costs two
    launches (trial steps, then objective + gradient at the accepted point) instead of up to `max_backtracks` + 1.
    """
    from .api import tikhonov_grad, tikhonov_pen
    from .configs import noise_shift
    p, wa = prob.params, prob.wa
    if p.objFuncType != 1:
        raise ValueError("run_optimizer_multistart handles objFuncType == 1 (no inequality constraint)")
    X = np.clip(np.atleast_2d(np.asarray(pcof0s, float)).copy(), prob.minCoeff, prob.maxCoeff)
    B, n = X.shape
    maxIter = prob.maxIter if maxIter is None else maxIter
    mem = prob.lbfgsMax if lbfgsMax is None else lbfgsMax
    scale = np.maximum(np.abs(prob.minCoeff), np.abs(prob.maxCoeff))
    scale[scale == 0] = 1.0
    lo, hi = prob.minCoeff / scale, prob.maxCoeff / scale
    plain = len(prob.nodes) == 1 and prob.nodes[0] == 0.0 and prob.weights[0] == 1.0
    shifts = None if plain else noise_shift(p.Ntot, prob.nodes)
    w = None if plain else prob.weights

    def fg(Y, need_grad=True):
        Xc = Y * scale
        r = wa.evaluate(Xc, shifts, w, evaladjoint=need_grad)
        f = (r["infid"] + r["leak"]).reshape(B) + np.array([tikhonov_pen(x, p) for x in Xc])
        if not need_grad:
            return f, None
        g = r["grad"].reshape(B, n) + np.array([tikhonov_grad(x, p) for x in Xc])
        return f, g * scale

    def f_only(Yrows):
        """Objective of arbitrary many points (trial steps): one objective-only launch."""
        Xc = Yrows * scale
        r = wa.evaluate(Xc, shifts, w, evaladjoint=False)
        return (r["infid"] + r["leak"]).reshape(len(Xc)) + np.array([tikhonov_pen(x, p) for x in Xc])

    Y = X / scale
    f, g = fg(Y)
    hist = [f.copy()]
    S, Yd = [], []                                   # per-iteration [B, n] curvature pairs
    valid = []                                       # [B] masks: pair usable for that member
    active = np.ones(B, bool)
    for _ in range(maxIter):
        # projected gradient: components pushing out of the box at an active bound are dropped
        pg = g.copy()
        pg[(Y <= lo) & (g > 0)] = 0.0
        pg[(Y >= hi) & (g < 0)] = 0.0
        active &= np.abs(pg).max(axis=1) > gtol
        if not active.any():
            break
        # two-loop recursion, all members at once
        q = pg.copy()
        alphas = []
        for s, yv, ok in zip(reversed(S), reversed(Yd), reversed(valid)):
            rho = np.where(ok, 1.0 / np.where(ok, (s * yv).sum(1), 1.0), 0.0)
            a = rho * (s * q).sum(1)
            q -= a[:, None] * yv
            alphas.append((a, rho))
        if S:
            s, yv, ok = S[-1], Yd[-1], valid[-1]
            gamma = np.where(ok, (s * yv).sum(1) / np.where(ok, (yv * yv).sum(1), 1.0), 1.0)
            q *= gamma[:, None]
        for (a, rho), s, yv in zip(reversed(alphas), S, Yd):
            b = rho * (yv * q).sum(1)
            q += (a - b)[:, None] * s
        d = -q
        d[pg == 0.0] = 0.0                            # variables held at an active bound do not move
        bad = (d * pg).sum(1) >= 0                    # not a descent direction: steepest descent
        d[bad] = -pg[bad]
        if not S:
            d /= np.maximum(1.0, np.abs(d).max(axis=1))[:, None] * 4.0       # first step: a quarter of the box at most

        def line_search(d, todo):
            """Projected Armijo backtracking for the members in `todo`; returns the mask of members that found no step.  Trial
            steps are evaluated `ls_batch` at a time in one launch; the smallest power of 1/2 that satisfies Armijo wins, exactly
            as in the one-trial-per-launch loop."""
            todo = todo.copy()
            k0 = 0
            while k0 < max_backtracks and todo.any():
                ks = np.arange(k0, min(k0 + ls_batch, max_backtracks))
                idx = np.nonzero(todo)[0]
                # [member, trial, n] trial points of the searching members; the other members ride along unchanged (fg needs B rows)
                Yt = np.clip(Y[idx, None, :] + (0.5 ** ks)[None, :, None] * d[idx, None, :], lo, hi)
                ft = f_only(Yt.reshape(-1, n)).reshape(len(idx), len(ks))
                ok_ = (ft <= f[idx, None] + 1e-4 * (g[idx, None, :] * (Yt - Y[idx, None, :])).sum(2)) & \
                      (np.abs(Yt - Y[idx, None, :]).max(axis=2) > 0)
                first = np.where(ok_.any(axis=1), ok_.argmax(axis=1), -1)
                hit = first >= 0
                Ynew[idx[hit]] = Yt[hit, first[hit]]
                fnew[idx[hit]] = ft[hit, first[hit]]
                todo[idx[hit]] = False
                k0 += len(ks)
            return todo

        Ynew, fnew = Y.copy(), f.copy()
        failed = line_search(d, active)
        if failed.any():                              # retry those members once along the projected steepest descent
            dsd = np.where(failed[:, None], -pg / np.maximum(1.0, np.abs(pg).max(axis=1))[:, None], 0.0)
            failed = line_search(dsd, failed)
        active &= ~failed                             # members with no acceptable step stop here
        fchk, gnew = fg(Ynew)
        s, yv = Ynew - Y, gnew - g
        ok = (s * yv).sum(1) > 1e-12 * np.sqrt((s * s).sum(1) * (yv * yv).sum(1))
        S.append(s); Yd.append(yv); valid.append(ok)
        if len(S) > mem:
            S.pop(0); Yd.pop(0); valid.pop(0)
        Y, f, g = Ynew, fchk, gnew
        hist.append(f.copy())
    return Y * scale, f, np.array(hist)
