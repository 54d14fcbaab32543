"""One pcof vector through the time-parallel evaluation (kernel 7), three times: with JQ_SEG_TIMING=1 the library prints the CUDA-event
time of every stage (propagator launch, joins, sweeps, sums); also the target of ncu launch lists.

    JQ_SEG_TIMING=1 python tools/timeparallel_stages.py cnot2 0        (0 = automatic number of segments; needs a GPU)
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq
from juqbox_b200 import configs
name = sys.argv[1]; ns = int(sys.argv[2])
cfg = configs.example(name)
pc = configs.synthetic_pcof(cfg, 1)
shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
wa.set_kernel(7); wa.set_time_segments(ns)
for _ in range(3):
    wa.evaluate(pc, shifts)
wa.close()
