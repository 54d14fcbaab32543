"""Latency of ONE traceobjgrad evaluation (nbatch = 1: what each Ipopt iteration of the reference calls) per kernel.

    python tools/single_eval_latency.py        (needs a GPU)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq                                    # noqa: E402
from juqbox_b200 import configs                             # noqa: E402
from oracle import oracle_traceobjgrad                      # noqa: E402

print("| config | kernel | kernel ms | host call ms (numpy in/out) | CPU oracle, 1 thread ms |")
print("|---|---|---|---|---|")
for name in configs.EXAMPLES:
    cfg = configs.example(name)
    pc = configs.synthetic_pcof(cfg, 1)
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
    w = cfg.weights if name == "risk_neutral" else None
    t0 = time.perf_counter()
    oracle_traceobjgrad(cfg.params, pc, shifts, nthreads=1)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    for k in (3, 4, 5, 7):
        wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
        try:
            wa.set_kernel(k)
        except Exception:
            wa.close()
            continue
        wa.evaluate(pc, shifts, w)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            wa.evaluate(pc, shifts, w)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"| {name} | {k} | {wa.last_kernel_ms:.3f} | {min(ts):.3f} | {cpu_ms:.1f} |", flush=True)
        wa.close()
