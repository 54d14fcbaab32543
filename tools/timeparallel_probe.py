#!/usr/bin/env python3
"""Time-parallel evaluation (kernel 7) of one pcof vector against the latency kernel: agreement and time, per number of segments.

    python tools/timeparallel_probe.py cnot2 8,16,32,64 [scale]        (needs a GPU)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq                                    # noqa: E402
from juqbox_b200 import configs                             # noqa: E402

name = sys.argv[1]
segs = [int(k) for k in sys.argv[2].split(",")]
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 1
cfg = configs.example(name)
pc = configs.synthetic_pcof(cfg, nb) * scale
shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
wa.set_kernel(0)
ref = wa.evaluate(pc, shifts)
for _ in range(2):
    ref = wa.evaluate(pc, shifts)
print(f"{name}: kernel {wa.last_kernel} {wa.last_kernel_ms:.3f} ms  infid {ref['infid'].ravel()[0]:.15e} leak {ref['leak'].ravel()[0]:.6e}", flush=True)
gn = np.linalg.norm(ref["grad"])
wa.set_kernel(7)
for ns in segs:
    wa.set_time_segments(ns)
    for adj in (True, False):
        r = wa.evaluate(pc, shifts, None, adj)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            r = wa.evaluate(pc, shifts, None, adj)
            ts.append((time.perf_counter() - t0) * 1e3)
        ei = np.abs(r["infid"] - ref["infid"]).max()
        el = np.abs(r["leak"] - ref["leak"]).max()
        eg = np.linalg.norm(r["grad"] - ref["grad"]) / gn if adj else 0.0
        print(f"  nseg {int(wa.query(7)):4d} adj {int(adj)}: kernels {wa.last_kernel_ms:.3f} ms, host call {min(ts):.3f} ms, launches {int(wa.query(2))}, CTAs {int(wa.query(4))}, "
              f"regs {int(wa.query(5))}; |d infid| {ei:.2e} |d leak| {el:.2e} rel d grad {eg:.2e}", flush=True)
wa.close()
