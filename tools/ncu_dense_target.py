"""One launch of the dense (FP64 MMA) kernel for ncu: 24 x 6 dense random operators, 128 candidates x 8 noise samples."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import juqbox_b200 as jq
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from generic_bench import dense_random
p, npar = dense_random(n=24, m=6, nsteps=400)
rng = np.random.default_rng(1)
wa = jq.Working_Arrays(p, npar)
r = wa.evaluate(rng.uniform(-0.02, 0.02, (128, npar)), rng.uniform(-0.05, 0.05, (8, p.Ntot)))
print("kernel", wa.last_kernel, "ms", wa.last_kernel_ms)
