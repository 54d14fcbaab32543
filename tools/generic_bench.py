#!/usr/bin/env python3
"""Throughput of the generic kernel (kernel id 1) on shapes only it serves, and of the 45 x 12 shape (8-warp fibre kernel).

    python tools/generic_bench.py        (needs a GPU)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq                                    # noqa: E402
from juqbox_b200 import configs                             # noqa: E402
from juqbox_b200.params import objparams                    # noqa: E402
from oracle import oracle_traceobjgrad                      # noqa: E402


def dense_random(n=7, m=3, Nc=2, Nfreq=2, D1=5, nsteps=300, sparse=False):
    rng = np.random.default_rng(12)
    sym = lambda a: (a + a.T) / 2
    H0 = sym(rng.standard_normal((n, n))) * 0.3
    Hs = [sym(rng.standard_normal((n, n))) for _ in range(Nc)]
    Ha = [(lambda a: (a - a.T) / 2)(rng.standard_normal((n, n))) for _ in range(Nc)]
    Vt = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))[0][:, :m]
    p = objparams([m], [n - m], 3.0, nsteps, Uinit=np.eye(n, m), Utarget=Vt, Cfreq=rng.standard_normal((Nc, Nfreq)), Rfreq=[1.0, 2.0],
                  Hconst=H0, Hsym_ops=Hs, Hanti_ops=Ha, use_sparse=sparse)
    return p, 2 * Nc * Nfreq * D1


def run(name, p, npar, B, kernel=0, check=True):
    rng = np.random.default_rng(1)
    pc = rng.uniform(-0.02, 0.02, (B, npar))
    wa = jq.Working_Arrays(p, npar)
    wa.set_kernel(kernel)
    r = wa.evaluate(pc)
    r = wa.evaluate(pc)
    ms = wa.last_kernel_ms
    err = float("nan")
    if check:
        o = oracle_traceobjgrad(p, pc[:2])
        err = max(np.linalg.norm(r["grad"][b, 0] - o["grad"][b, 0]) / np.linalg.norm(o["grad"][b, 0]) for b in range(2))
    print(f"| {name} | {p.Ntot} x {p.N} | {p.nsteps} | {wa.last_kernel} | {B} | {ms:.2f} | {B / (ms * 1e-3):.4g} | {err:.1e} | ctas {int(wa.query(4))} smem {int(wa.query(6))} regs {int(wa.query(5))} |", flush=True)
    wa.close()


if __name__ == "__main__":
    print("| shape | n x m | nsteps | kernel | candidates | kernel ms | evals/s | rel. grad err vs oracle | launch |")
    print("|---|---|---|---|---|---|---|---|---|")
    def run_samples(name, p, npar, B, S, kernel):
        """B candidates x S noise samples (shared pcof per candidate): what the dense kernel batches into one contraction."""
        rng = np.random.default_rng(1)
        pc = rng.uniform(-0.02, 0.02, (B, npar))
        sh = rng.uniform(-0.05, 0.05, (S, p.Ntot))
        wa = jq.Working_Arrays(p, npar)
        wa.set_kernel(kernel)
        wa.evaluate(pc, sh)
        wa.evaluate(pc, sh)
        ms = wa.last_kernel_ms
        print(f"| {name} | {p.Ntot} x {p.N} | {p.nsteps} | {wa.last_kernel} | {B} x {S} samples | {ms:.2f} | {B * S / (ms * 1e-3):.4g} | - | ctas {int(wa.query(4))} smem {int(wa.query(6))} regs {int(wa.query(5))} |", flush=True)
        wa.close()


    for k in (1, 6):
        p, npar = dense_random()
        run("dense random 7 x 3 (tests)", p, npar, 4096, k)
        p, npar = dense_random(n=24, m=6, nsteps=400)
        run("dense random 24 x 6", p, npar, 1024, k)
        run_samples("dense random 24 x 6", p, npar, 128, 8, k)
        p, npar = dense_random(n=45, m=12, nsteps=200)
        run("dense random 45 x 12", p, npar, 256, k)
        p, npar = dense_random(n=64, m=8, nsteps=200)
        try:
            run("dense random 64 x 8", p, npar, 256, k)
        except Exception as e:       # the dense kernel's blocks + K(t), S(t) do not fit in shared memory: automatic mode uses the generic kernel
            print(f"| dense random 64 x 8 | 64 x 8 | 200 | {k} | 256 | - | - | - | {str(e).split(':')[-1].strip()} |")
    cfg = configs.qudit_system([3, 2, 2], [2, 1, 1])
    run("qudits 5 x 3 x 3 (45 x 12), auto", cfg.params, cfg.nCoeff, 512, 0)
    run("qudits 5 x 3 x 3 (45 x 12), generic", cfg.params, cfg.nCoeff, 128, 1)
    cfg = configs.qudit_system([2, 2, 2], [1, 1, 1], exchange=0.02)
    run("qudits 3 x 3 x 3 with remote exchange (27 x 8), generic", cfg.params, cfg.nCoeff, 512, 0)
    cfg = configs.example("cnot2")
    run("cnot2 example on the generic kernel", cfg.params, cfg.nCoeff, 512, 1)
