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
