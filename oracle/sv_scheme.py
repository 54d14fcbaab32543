"""ORACLE (test infrastructure): the Stormer-Verlet scheme of the reference in its textbook form — exact linear solves and
function forcing — restating Juqbox.step (src/StormerVerlet.jl:36-91) and the analytic 2x2 problems of
test/test-stormer-verlet.jl:7-135.  Pinned by test/reference_solutions/err-mat-ref.jld2 at 1e-13 (:137-182).
The production steppers use the same stage equations with the truncated Neumann solve (src/StormerVerlet.jl:461-504)."""
import math

import numpy as np


def step(K, S, t, u, v, h, uforce, vforce):
    """One step; returns (t+h, u, v, v05).  src/StormerVerlet.jl:36-48 (forcing functions) + :65-91."""
    In = np.eye(len(u))
    uforce0, vforce05, uforce1 = uforce(t), vforce(t + 0.5 * h), uforce(t + h)
    K0, S0, K05, S05, K1, S1 = K(t), S(t), K(t + 0.5 * h), S(t + 0.5 * h), K(t + h), S(t + h)
    rhs = K05 @ u + S05 @ v + vforce05
    l1 = np.linalg.solve(In - 0.5 * h * S05, rhs)
    v05 = v + 0.5 * h * l1
    kappa1 = S0 @ u - K0 @ v05 + uforce0
    rhs = S1 @ (u + 0.5 * h * kappa1) - K1 @ v05 + uforce1
    kappa2 = np.linalg.solve(In - 0.5 * h * S1, rhs)
    u = u + 0.5 * h * (kappa1 + kappa2)
    l2 = K05 @ u + S05 @ v05 + vforce05
    v = v + 0.5 * h * (l1 + l2)
    return t + h, u, v, v05


def timesteptest(cfl, testcase):
    """test/test-stormer-verlet.jl:7-135 -> (cg_err, ce_err) at the final time."""
    if testcase in (1, 2):
        K0, S0 = np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros((2, 2))
    else:
        K0, S0 = np.zeros((2, 2)), np.array([[0.0, 1.0], [-1.0, 0.0]])
    T, omega = 5 * math.pi, 2 * math.pi
    maxeig = np.max(np.abs(np.linalg.eigvals(K0 + S0)))
    dt = cfl / maxeig
    nsteps = int(math.ceil(T / dt))
    dt = T / nsteps
    zero = lambda t: np.zeros(2)
    phi1 = lambda t: 0.25 * (t - math.sin(omega * t) / omega)
    phidot = lambda t: 0.5 * math.sin(0.5 * omega * t) ** 2
    quad = lambda t: 4 / T ** 2 * t * (T - t)
    if testcase == 1:
        timefunc, uforce, vforce = (lambda t: 0.25 * (1.0 - math.cos(omega * t))), zero, zero
    elif testcase == 0:
        timefunc, uforce, vforce = (lambda t: 0.25 * (1 - math.sin(omega * t))), zero, zero
    elif testcase == 2:
        timefunc = quad
        uforce = lambda t: np.array([(quad(t) - phidot(t)) * math.sin(phi1(t)), 0.0])
        vforce = lambda t: np.array([0.0, -(quad(t) - phidot(t)) * math.cos(phi1(t))])
    else:
        timefunc = quad
        uforce = lambda t: np.array([-phidot(t) * math.sin(phi1(t)), quad(t) * math.cos(phi1(t))])
        vforce = lambda t: np.array([-quad(t) * math.sin(phi1(t)), phidot(t) * math.cos(phi1(t))])
    K = lambda t: timefunc(t) * K0
    S = lambda t: timefunc(t) * S0
    u, v, t = np.array([1.0, 0.0]), np.zeros(2), 0.0
    for _ in range(nsteps):
        t, u, v, _ = step(K, S, t, u, v, dt, uforce, vforce)
    if testcase in (1, 2, 3):
        phi = 0.25 * (t - 1.0 / omega * math.sin(omega * t))
        cg, ce = math.cos(phi), -1j * math.sin(phi)
    else:
        phi = 0.25 * (t + 1 / omega * (math.cos(omega * t) - 1.0))
        cg, ce = math.cos(phi), -math.sin(phi)
    cg_err = math.sqrt((u[0] - np.real(cg)) ** 2 + (v[0] + np.imag(cg)) ** 2)
    ce_err = math.sqrt((u[1] - np.real(ce)) ** 2 + (v[1] + np.imag(ce)) ** 2)
    return cg_err, ce_err
