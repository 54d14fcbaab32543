#!/usr/bin/env python3
"""End-to-end use of the drop-in API: optimise the SWAP 0-2 gate of the risk-neutral example, first with the nominal
Hamiltonian (nquad = 1), then risk-neutral with 9 Gauss-Legendre noise samples per evaluation
(examples/Risk_Neutral/run_all.jl:90-100,133-140), and sweep the objective over the Hamiltonian perturbation like
ep_plot (:6-32).  Needs a GPU."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import juqbox_b200 as jq
from juqbox_b200 import configs

cfg = configs.example("risk_neutral")
p = cfg.params
p.quiet = True
wa = jq.Working_Arrays(p, cfg.nCoeff)
pcof0 = configs.synthetic_pcof(cfg, 1)[0]
minC, maxC = jq.assign_thresholds_freq([cfg.maxpar[0]] * p.Nfreq, p.Ncoupled, p.Nfreq, cfg.D1)
maxiter = int(sys.argv[1]) if len(sys.argv) > 1 else 60
ep_vals = np.linspace(-2 * np.pi * 3e-2, 2 * np.pi * 3e-2, 1001)

for label, nodes, weights in (("nominal", [0.0], [1.0]), ("risk-neutral", cfg.nodes, cfg.weights)):
    prob = jq.setup_ipopt_problem(p, wa, cfg.nCoeff, minC, maxC, maxIter=maxiter, lbfgsMax=5, nodes=nodes, weights=weights)
    t0 = time.perf_counter()
    pcof = jq.run_optimizer(prob, pcof0)
    dt = time.perf_counter() - t0
    sweep = jq.traceobjgrad_batch(pcof, p, wa, nodes=ep_vals, evaladjoint=False)      # 1001 trajectories, one launch
    print(f"{label:13s}: {len(p.objHist)} iterations in {dt:.1f}s, objective {p.objHist[-1]:.3e}, infidelity {p.lastTraceInfidelity:.3e}; "
          f"sweep over eps: max objective {sweep['objf'].max():.3e}, at eps=0 {sweep['objf'][0, 500]:.3e}  [{prob.status}]")
    p.objHist.clear(); p.primaryHist.clear(); p.secondaryHist.clear()
wa.close()
