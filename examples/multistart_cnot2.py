#!/usr/bin/env python3
"""Multi-start optimisation of the two-qubit CNOT of examples/cnot2-setup.jl: B random initial coefficient vectors
(the example's own `0.01*maxpar*rand`, scaled up) optimised in lock step; every objective / gradient request of all B
members is one batched GPU call.  Needs a GPU.

    python examples/multistart_cnot2.py [B=256] [iterations=40]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import juqbox_b200 as jq
from juqbox_b200 import configs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = configs.example("cnot2")
p = cfg.params
p.quiet = True
p.tik0 = 0.01
wa = jq.Working_Arrays(p, cfg.nCoeff)
minC, maxC = jq.assign_thresholds(p, cfg.D1, cfg.maxpar)
rng = np.random.default_rng(2456)
starts = rng.uniform(-1, 1, (B, cfg.nCoeff)) * 0.2 * np.minimum(np.abs(minC), np.abs(maxC))
prob = jq.setup_ipopt_problem(p, wa, cfg.nCoeff, minC, maxC, maxIter=iters, lbfgsMax=10)
t0 = time.perf_counter()
pcofs, f, hist = jq.run_optimizer_multistart(prob, starts)
dt = time.perf_counter() - t0
res = wa.evaluate(pcofs, evaladjoint=False)
infid = res["infid"].ravel()
order = np.argsort(f)
print(f"{B} starts x {hist.shape[0] - 1} L-BFGS iterations in {dt:.2f}s ({wa.last_kernel_ms:.1f} ms per batched objective evaluation)")
print(f"objective: start median {np.median(hist[0]):.3e} -> final best {f[order[0]]:.3e}, median {np.median(f):.3e}, worst {f[order[-1]]:.3e}")
print(f"gate infidelity of the best five starts: {', '.join('%.2e' % infid[i] for i in order[:5])}")
wa.close()
